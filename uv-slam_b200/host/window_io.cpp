// window_io.cpp — see window_io.h.  Layout: "UVSWIN01" | estimate_extrinsic, estimate_td (int32 LE) | for every array
// in the fixed order of window.py (_F64 then _I32): ndim, shape[ndim] (int32 LE), data (float64 / int32 LE).
#include "window_io.h"

#include <cstdint>
#include <cstdio>
#include <vector>

namespace uvs_host {
namespace {

struct Writer {
  std::FILE *f;
  bool ok = true;
  void raw(const void *p, size_t n) { if (ok && n && std::fwrite(p, 1, n, f) != n) ok = false; }
  void i32(int32_t v) { raw(&v, 4); }
  template <typename T>
  void array(const T *p, std::initializer_list<int> shape) {
    i32((int32_t)shape.size());
    size_t cnt = 1;
    for (int s : shape) { i32(s); cnt *= (size_t)s; }
    if (cnt && !p) { ok = false; return; }
    raw(p, cnt * sizeof(T));
  }
};

}  // namespace

// the format is little-endian; the writer emits host byte order
static_assert(
#if defined(__BYTE_ORDER__) && defined(__ORDER_LITTLE_ENDIAN__)
    __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__,
#else
    true,
#endif
    "uvs_window v1 is little-endian: add a byte swap to Writer::raw for this target");

int save_window(const UvsWindow &w, const char *path) {
  if (!path) return UVS_ERR_INVALID_ARG;
  // the same size / consistency checks as uvs_upload_windows, before anything is written
  if (w.n_relo < 0 || (w.n_relo > 0 && (!w.relo_pose || !w.relo_point || !w.relo_pts_j))) return UVS_ERR_INVALID_ARG;
  if (w.n_frames < 1 || w.n_points < 0 || w.n_lines < 0 || w.n_proj < 0 || w.n_line_obs < 0 || w.n_vp_obs < 0 || w.n_imu < 0 ||
      w.prior_n < 0 || w.prior_n_blocks < 0)
    return UVS_ERR_INVALID_ARG;
  if (w.prior_n > 0) {
    if (!w.prior_block_kind || !w.prior_block_id || !w.prior_J || !w.prior_r || !w.prior_x0 || w.prior_n_blocks == 0) return UVS_ERR_INVALID_ARG;
    int cols = 0;
    for (int b = 0; b < w.prior_n_blocks; b++) {
      const int k = w.prior_block_kind[b];
      cols += (k == UVS_BLOCK_POSE || k == UVS_BLOCK_EXPOSE) ? 6 : (k == UVS_BLOCK_SPEEDBIAS ? 9 : 1);
    }
    if (cols != w.prior_n) return UVS_ERR_INVALID_ARG;   // local sizes of the kept blocks must add up to prior_n
  }
  // global sizes of the kept blocks of the prior: pose / extrinsic 7, speed-bias 9, td 1 (marginalization_factor.cpp:203-215)
  int x0_len = 0;
  if (w.prior_n > 0) {
    if (!w.prior_block_kind) return UVS_ERR_INVALID_ARG;
    for (int b = 0; b < w.prior_n_blocks; b++) {
      const int k = w.prior_block_kind[b];
      if (k == UVS_BLOCK_POSE || k == UVS_BLOCK_EXPOSE) x0_len += 7;
      else if (k == UVS_BLOCK_SPEEDBIAS) x0_len += 9;
      else if (k == UVS_BLOCK_TD) x0_len += 1;
      else return UVS_ERR_INVALID_ARG;
    }
  }
  std::FILE *f = std::fopen(path, "wb");
  if (!f) return UVS_ERR_UNSUPPORTED;
  Writer o{f};
  o.raw("UVSWIN01", 8);
  o.i32(w.estimate_extrinsic); o.i32(w.estimate_td);
  const int F = w.n_frames, P = w.n_points, L = w.n_lines, np = w.n_proj, nl = w.n_line_obs, nv = w.n_vp_obs, ni = w.n_imu;
  const int ntd = (w.estimate_td && w.proj_vel_i) ? np : 0;   // the td extras may be absent (NULL) without estimate_td
  const int pn = w.prior_n > 0 ? w.prior_n : 0, pb = w.prior_n > 0 ? w.prior_n_blocks : 0;
  // ---- float64 arrays, order of window.py:_F64
  o.array(w.pose, {F, 7}); o.array(w.speed_bias, {F, 9}); o.array(w.ex_pose, {7}); o.array(w.td, {1});
  o.array(w.inv_depth, {P}); o.array(w.ortho, {L, 4});
  o.array(w.proj_pts_i, {np, 3}); o.array(w.proj_pts_j, {np, 3});
  o.array(w.proj_vel_i, {ntd, 2}); o.array(w.proj_vel_j, {ntd, 2});
  o.array(w.proj_td_i, {ntd}); o.array(w.proj_td_j, {ntd}); o.array(w.proj_row_i, {ntd}); o.array(w.proj_row_j, {ntd});
  o.array(w.line_sp, {nl, 2}); o.array(w.line_ep, {nl, 2}); o.array(w.vp_dir, {nv, 3});
  o.array(w.line_ric, {3, 3}); o.array(w.line_tic, {3});
  o.array(w.imu_delta_p, {ni, 3}); o.array(w.imu_delta_q, {ni, 4}); o.array(w.imu_delta_v, {ni, 3}); o.array(w.imu_sum_dt, {ni});
  o.array(w.imu_lin_ba, {ni, 3}); o.array(w.imu_lin_bg, {ni, 3}); o.array(w.imu_jacobian, {ni, 225}); o.array(w.imu_covariance, {ni, 225});
  o.array(w.prior_J, {pn, pn}); o.array(w.prior_r, {pn}); o.array(w.prior_x0, {x0_len});
  // ---- int32 arrays, order of window.py:_I32
  o.array(w.proj_frame_i, {np}); o.array(w.proj_frame_j, {np}); o.array(w.proj_point, {np});
  o.array(w.line_frame, {nl}); o.array(w.line_idx, {nl}); o.array(w.vp_frame, {nv}); o.array(w.vp_line, {nv});
  o.array(w.imu_frame_i, {ni}); o.array(w.prior_block_kind, {pb}); o.array(w.prior_block_id, {pb});
  if (w.n_relo > 0) {   // optional relocalisation section (window.py:_RELO_F64 / _RELO_I32)
    o.array(w.relo_pose, {7}); o.array(w.relo_pts_j, {w.n_relo, 3}); o.array(w.relo_point, {w.n_relo});
  }
  const bool closed = std::fclose(f) == 0;
  return (o.ok && closed) ? UVS_OK : UVS_ERR_UNSUPPORTED;
}

}  // namespace uvs_host

extern "C" int uvs_host_save_window(const UvsWindow *w, const char *path) {
  if (!w) return UVS_ERR_INVALID_ARG;
  return uvs_host::save_window(*w, path);
}
