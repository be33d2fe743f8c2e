// uvs_api.cu — the C ABI of libuvs_b200.so (include/uvs.h): host-side packing of the caller's
// windows into one flat HBM batch, launch sequencing of the sm_100a kernels, and the LM loop.
//
// Replaces the body of Estimator::optimization() between vector2double() and double2vector()
// (vins_estimator/src/estimator.cpp:800-999): problem assembly + ceres::Solve.  There is no CPU
// fallback: every entry point fails with UVS_ERR_CUDA when no device is available.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/uvs.h"
#include "uvs_device.cuh"
#include "uvs_handle.h"
#include "uvs_kernels.h"

using namespace uvs;

namespace {

int fail(UvsHandle *h, int status, const std::string &msg) {
  if (h) h->err = msg;
  return status;
}
int cuda_fail(UvsHandle *h, cudaError_t e, const char *where) {
  return fail(h, UVS_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                       \
  do {                                                                 \
    cudaError_t e_ = (call);                                           \
    if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);             \
  } while (0)

void fill_params(const UvsOptions &o, Params &P) {
  P.S = o.focal_length / 1.6;   // ProjectionFactor::sqrt_info = FOCAL_LENGTH / 1.6 * I2 (estimator.cpp:17)
  for (int k = 0; k < 3; k++) P.g[k] = o.gravity[k];
  P.line_factor = o.line_factor; P.vp_factor = o.vp_factor;
  P.cauchy_point = o.cauchy_point; P.cauchy_line = o.cauchy_line; P.cauchy_vp = o.cauchy_vp;
  P.tr_over_row = o.row != 0.0 ? o.tr / o.row : 0.0;
  P.half_row = o.row / 2.0;
  P.min_lm_diag = o.min_lm_diagonal; P.max_lm_diag = o.max_lm_diagonal;
  P.min_relative_decrease = o.min_relative_decrease;
  P.max_radius = o.max_radius; P.min_radius = o.min_radius; P.initial_radius = o.initial_radius;
  P.function_tolerance = o.function_tolerance; P.gradient_tolerance = o.gradient_tolerance;
  P.parameter_tolerance = o.parameter_tolerance;
  P.fixed_iterations = o.fixed_iterations;
  P.max_num_iterations = o.max_num_iterations;
}

template <class T>
void prefix(std::vector<T> &off, int B, const UvsWindow *w, T (*get)(const UvsWindow &)) {
  off.assign(B + 1, 0);
  for (int i = 0; i < B; i++) off[i + 1] = off[i] + get(w[i]);
}

int prior_local(int kind) { return (kind == UVS_BLOCK_POSE || kind == UVS_BLOCK_EXPOSE) ? 6 : (kind == UVS_BLOCK_SPEEDBIAS ? 9 : 1); }
int prior_global(int kind) { return (kind == UVS_BLOCK_POSE || kind == UVS_BLOCK_EXPOSE) ? 7 : (kind == UVS_BLOCK_SPEEDBIAS ? 9 : 1); }

// owned arrays of a window expanded by its relocalisation pose (see upload_enqueue)
struct RelocExpansion {
  std::vector<double> pose, sb, pts_i, pts_j;
  std::vector<int> fi, fj, pt;
};
int expand_relocalisation(const UvsWindow &x, RelocExpansion &E, UvsWindow &out) {
  if (!x.relo_pose || !x.relo_point || !x.relo_pts_j || !x.pose || !x.speed_bias) return 1;
  if (x.estimate_td) return 3;
  if (x.n_proj > 0 && (!x.proj_frame_i || !x.proj_frame_j || !x.proj_point || !x.proj_pts_i || !x.proj_pts_j)) return 1;
  const int F = x.n_frames, np = x.n_proj, nr = x.n_relo;
  for (int r = 0; r < nr; r++) {
    if (x.relo_point[r] < 0 || x.relo_point[r] >= x.n_points) return 1;
    if (r > 0 && x.relo_point[r] <= x.relo_point[r - 1]) return 1;
  }
  E.pose.assign(x.pose, x.pose + 7 * (size_t)F); E.pose.insert(E.pose.end(), x.relo_pose, x.relo_pose + 7);
  E.sb.assign(x.speed_bias, x.speed_bias + 9 * (size_t)F); E.sb.insert(E.sb.end(), 9, 0.0);
  std::vector<int> relo_of(x.n_points, -1);
  for (int r = 0; r < nr; r++) relo_of[x.relo_point[r]] = r;
  int placed = 0;
  for (int k = 0; k < np; k++) {
    E.fi.push_back(x.proj_frame_i[k]); E.fj.push_back(x.proj_frame_j[k]); E.pt.push_back(x.proj_point[k]);
    E.pts_i.insert(E.pts_i.end(), x.proj_pts_i + 3 * (size_t)k, x.proj_pts_i + 3 * (size_t)k + 3);
    E.pts_j.insert(E.pts_j.end(), x.proj_pts_j + 3 * (size_t)k, x.proj_pts_j + 3 * (size_t)k + 3);
    const int p = x.proj_point[k];
    if (p < 0 || p >= x.n_points) return 1;
    const bool last_of_point = k + 1 == np || x.proj_point[k + 1] != p;
    if (last_of_point && relo_of[p] >= 0) {   // the point's factor group ends here: its relocalisation factor joins it
      const int r = relo_of[p];
      E.fi.push_back(x.proj_frame_i[k]); E.fj.push_back(F); E.pt.push_back(p);
      E.pts_i.insert(E.pts_i.end(), x.proj_pts_i + 3 * (size_t)k, x.proj_pts_i + 3 * (size_t)k + 3);
      E.pts_j.insert(E.pts_j.end(), x.relo_pts_j + 3 * (size_t)r, x.relo_pts_j + 3 * (size_t)r + 3);
      relo_of[p] = -2;
      placed++;
    }
  }
  if (placed != nr) return 2;
  out = x;
  out.n_frames = F + 1; out.n_proj = np + nr; out.n_relo = 0;
  out.pose = E.pose.data(); out.speed_bias = E.sb.data();
  out.proj_frame_i = E.fi.data(); out.proj_frame_j = E.fj.data(); out.proj_point = E.pt.data();
  out.proj_pts_i = E.pts_i.data(); out.proj_pts_j = E.pts_j.data();
  return 0;
}

int post_launch(UvsHandle *h, const char *where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, where);
  return UVS_OK;
}

}  // namespace

extern "C" {

int uvs_abi_version(void) { return UVS_ABI_VERSION; }

void uvs_default_options(UvsOptions *o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  // config/euroc/euroc_config.yaml:20,55-56,64,85-87 and Ceres' trust-region defaults
  o->focal_length = 461.6;
  o->gravity[0] = 0.0; o->gravity[1] = 0.0; o->gravity[2] = 9.81007;
  o->line_factor = 300.0; o->vp_factor = 10.0;
  o->cauchy_point = 1.0; o->cauchy_line = 0.1; o->cauchy_vp = 1.0;
  o->tr = 0.0; o->row = 480.0;
  o->max_num_iterations = 10; o->fixed_iterations = 0; o->max_solver_time = 0.0;
  o->initial_radius = 1e4; o->max_radius = 1e16; o->min_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
}

const char *uvs_status_string(int s) {
  switch (s) {
    case UVS_OK: return "ok";
    case UVS_ERR_INVALID_ARG: return "invalid argument";
    case UVS_ERR_CUDA: return "CUDA error (no device, or a runtime failure; there is no CPU fallback)";
    case UVS_ERR_CAPACITY: return "capacity exceeded";
    case UVS_ERR_NOT_FINITE: return "non-finite cost or step";
    case UVS_ERR_NOT_PD: return "reduced camera system not positive definite";
    case UVS_ERR_NO_WINDOW: return "no window uploaded";
    case UVS_ERR_COMM: return "multi-GPU exchange failed";
    case UVS_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
  }
}

int uvs_create(int device, UvsHandle **out) {
  if (!out) return UVS_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) { cudaGetLastError(); return UVS_ERR_CUDA; }
  if (cudaSetDevice(device) != cudaSuccess) return UVS_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return UVS_ERR_CUDA;
  if (prop.major < 10) return UVS_ERR_CUDA;   // built for sm_100a only
  UvsHandle *h = new UvsHandle();
  h->device = device;
  h->stage.pinned_host = true;
  h->hscratch.pinned_host = true;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return UVS_ERR_CUDA; }
  cudaEventCreate(&h->ev_a); cudaEventCreate(&h->ev_b); cudaEventCreate(&h->ev_c); cudaEventCreate(&h->ev_d);
  for (int k = 0; k < 3; k++) {
    if (cudaStreamCreateWithFlags(&h->fork.aux[k], cudaStreamNonBlocking) != cudaSuccess) { delete h; return UVS_ERR_CUDA; }
    cudaEventCreateWithFlags(&h->fork.join[k], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&h->fork.fork, cudaEventDisableTiming);
  const size_t chol_smem = chol_max_dynamic_smem(prop.sharedMemPerBlockOptin);
  if (chol_smem == 0) { cudaGetLastError(); cudaStreamDestroy(h->stream); delete h; return UVS_ERR_CUDA; }
  h->packed_limit = chol_packed_limit(chol_smem);
  h->smem_optin = prop.sharedMemPerBlockOptin;
  cudaMalloc((void **)&h->d_active, sizeof(int));
  cudaMallocHost((void **)&h->h_active, sizeof(int));
  uvs_default_options(&h->opts);
  *out = h;
  return UVS_OK;
}

int uvs_destroy(UvsHandle *h) {
  if (!h) return UVS_ERR_INVALID_ARG;
  for (UvsHandle *c : h->children) uvs_destroy(c);
  h->children.clear();
  cudaSetDevice(h->device);
  uvs::resident_destroy(h);
  cudaStreamSynchronize(h->stream);
  h->dev.release(); h->stage.release(); h->scratch.release(); h->hscratch.release();
  if (h->d_active) cudaFree(h->d_active);
  if (h->h_active) cudaFreeHost(h->h_active);
  cudaEventDestroy(h->ev_a); cudaEventDestroy(h->ev_b); cudaEventDestroy(h->ev_c); cudaEventDestroy(h->ev_d);
  if (h->iter_exec) cudaGraphExecDestroy(h->iter_exec);
  if (h->nccl_comm) uvs::nccl_destroy(h->nccl_comm);
  for (int k = 0; k < 3; k++) { cudaStreamDestroy(h->fork.aux[k]); cudaEventDestroy(h->fork.join[k]); }
  cudaEventDestroy(h->fork.fork);
  for (cudaEvent_t e : h->stage_ev) cudaEventDestroy(e);
  cudaStreamDestroy(h->stream);
  delete h;
  return UVS_OK;
}

const char *uvs_last_error(const UvsHandle *h) { return h ? h->err.c_str() : "null handle"; }

static int upload_enqueue(UvsHandle *h, int32_t B, const UvsWindow *w, const UvsOptions *opts, const ResidentHook *hook = nullptr);
static int upload_finish(UvsHandle *h);

int uvs_upload_windows(UvsHandle *h, int32_t B, const UvsWindow *w, const UvsOptions *opts) {
  const int rc = upload_enqueue(h, B, w, opts);
  return rc ? rc : upload_finish(h);
}

// host packing + every device operation of an upload, enqueued on the handle's stream without waiting.  With a resident
// hook (uvs_window.cu) the factor / IMU / prior-matrix sections are not copied from the caller (their pointers are null)
// but written on the device by hook->fill.
static int upload_enqueue(UvsHandle *h, int32_t B, const UvsWindow *w, const UvsOptions *opts, const ResidentHook *hook) {
  if (!h || B <= 0 || !w) return fail(h, UVS_ERR_INVALID_ARG, "uvs_upload_windows: bad arguments");
  CK(cudaSetDevice(h->device));
  // Relocalisation factors (estimator.cpp:944-978): relo_Pose is a pose block without IMU factors, its factors are point
  // factors whose observing pose is that block.  Inside the library it is one more FRAME at the end of the window whose
  // speed-bias block has no factor (zero Hessian rows: the LM diagonal keeps them at a zero step, the rest of the system
  // is what Ceres sees), and every relocalisation factor joins the factor group of its point.
  std::vector<UvsWindow> expanded;
  std::vector<RelocExpansion> reloc_store;
  h->relo.assign(B, 0);
  {
    bool any = false;
    for (int i = 0; i < B; i++) any = any || w[i].n_relo > 0;
    if (any) {
      if (hook) return fail(h, UVS_ERR_UNSUPPORTED, "device-resident window: relocalisation factors are not supported");
      expanded.assign(w, w + B);
      reloc_store.resize(B);
      for (int i = 0; i < B; i++) {
        if (w[i].n_relo <= 0) continue;
        const int rc = expand_relocalisation(w[i], reloc_store[i], expanded[i]);
        if (rc == 1) return fail(h, UVS_ERR_INVALID_ARG, "relocalisation: null array / index out of range / not ascending in window " + std::to_string(i));
        if (rc == 2) return fail(h, UVS_ERR_INVALID_ARG, "relocalisation: a matched point owns no projection factor in window " + std::to_string(i));
        if (rc == 3) return fail(h, UVS_ERR_UNSUPPORTED, "relocalisation factors cannot be combined with estimate_td");
        h->relo[i] = 1;
      }
      w = expanded.data();
    }
  }
  static const bool trace = std::getenv("UVS_TRACE") != nullptr;   // host-side phase timings on stderr (adds syncs)
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  auto t_phase = now();
  if (opts) h->opts = *opts;
  fill_params(h->opts, h->P);
  h->have_window = false;
  const int td = w[0].estimate_td ? 1 : 0;
  int max_d = 0, max_prior_n = 0, max_frames = 0, max_lines = 0, max_lobs = 0;
  bool any_ex = false;
  for (int i = 0; i < B; i++) {
    const UvsWindow &x = w[i];
    if (x.n_frames < 1 || x.n_points < 0 || x.n_lines < 0 || x.n_proj < 0 || x.n_line_obs < 0 || x.n_vp_obs < 0 ||
        x.n_imu < 0 || x.prior_n < 0 || x.prior_n_blocks < 0)
      return fail(h, UVS_ERR_INVALID_ARG, "negative size in window " + std::to_string(i));
    if ((x.estimate_td ? 1 : 0) != td) return fail(h, UVS_ERR_UNSUPPORTED, "estimate_td must be uniform over the batch");
    if (!x.pose || !x.speed_bias || !x.ex_pose || (x.n_points && !x.inv_depth) || (x.n_lines && !x.ortho))
      return fail(h, UVS_ERR_INVALID_ARG, "null state pointer in window " + std::to_string(i));
    if (hook && (B != 1 || td)) return fail(h, UVS_ERR_UNSUPPORTED, "device-resident window: one window, no td");
    if (x.n_proj && !hook && (!x.proj_frame_i || !x.proj_frame_j || !x.proj_point || !x.proj_pts_i || !x.proj_pts_j))
      return fail(h, UVS_ERR_INVALID_ARG, "null projection-factor array");
    if (td && x.n_proj && (!x.proj_vel_i || !x.proj_vel_j || !x.proj_td_i || !x.proj_td_j || !x.proj_row_i || !x.proj_row_j || !x.td))
      return fail(h, UVS_ERR_INVALID_ARG, "estimate_td needs the td arrays");
    if (x.n_line_obs && !hook && (!x.line_frame || !x.line_idx || !x.line_sp || !x.line_ep)) return fail(h, UVS_ERR_INVALID_ARG, "null line-factor array");
    if (x.n_vp_obs && !hook && (!x.vp_frame || !x.vp_line || !x.vp_dir)) return fail(h, UVS_ERR_INVALID_ARG, "null VP-factor array");
    if ((x.n_line_obs || x.n_vp_obs) && (!x.line_ric || !x.line_tic)) return fail(h, UVS_ERR_INVALID_ARG, "line_ric / line_tic missing");
    if (x.n_imu && !hook && (!x.imu_frame_i || !x.imu_delta_p || !x.imu_delta_q || !x.imu_delta_v || !x.imu_sum_dt || !x.imu_lin_ba ||
                    !x.imu_lin_bg || !x.imu_jacobian || !x.imu_covariance))
      return fail(h, UVS_ERR_INVALID_ARG, "null IMU array");
    if (x.prior_n && ((!hook && (!x.prior_J || !x.prior_r)) || !x.prior_block_kind || !x.prior_block_id || !x.prior_x0 || !x.prior_n_blocks))
      return fail(h, UVS_ERR_INVALID_ARG, "null prior array");
    if (x.n_frames > 32) return fail(h, UVS_ERR_CAPACITY, "more than 32 frames per window (a landmark's factors must fit one warp)");
    const int d = 15 * x.n_frames + (x.estimate_extrinsic ? 6 : 0) + (td ? 1 : 0);
    max_d = std::max(max_d, d);
    max_frames = std::max(max_frames, (int)x.n_frames);
    max_lines = std::max(max_lines, (int)x.n_lines);
    max_lobs = std::max(max_lobs, (int)x.n_line_obs);
    any_ex = any_ex || x.estimate_extrinsic;
    max_prior_n = std::max(max_prior_n, (int)x.prior_n);
  }
  h->B = B; h->max_d = max_d; h->max_prior_n = max_prior_n;
  if (h->iter_exec) { cudaGraphExecDestroy(h->iter_exec); h->iter_exec = nullptr; }   // the graph holds the old batch's arguments
  h->max_frames = max_frames; h->any_ex = any_ex; h->max_lines = max_lines;
  // independent kernels of a stage run side by side on the auxiliary streams (fork / join events).  This pays for a single
  // window too: its kernels are latency-bound (2.45 -> 1.89 ms per 10-iteration solve); UVS_SERIAL=1 forces the plain
  // in-order sequence (profiling), UVS_CONCURRENT_MIN=<B> restores a batch-size threshold
  static const bool force_serial = std::getenv("UVS_SERIAL") != nullptr;
  static const int concurrent_min = std::getenv("UVS_CONCURRENT_MIN") ? std::atoi(std::getenv("UVS_CONCURRENT_MIN")) : 1;
  h->concurrent = B >= concurrent_min && !force_serial;
  // landmark path: thread-per-landmark elimination + per-window dense rank update when the pose-pose system
  // fits shared memory (the reference's window size); otherwise the general warp-per-landmark path
  h->use_build3 = !td && (max_frames + (any_ex ? 1 : 0)) <= 12 && max_frames >= 2;
  prefix<int>(h->frame_off, B, w, [](const UvsWindow &x) { return (int)x.n_frames; });
  prefix<int>(h->point_off, B, w, [](const UvsWindow &x) { return (int)x.n_points; });
  prefix<int>(h->line_off, B, w, [](const UvsWindow &x) { return (int)x.n_lines; });
  prefix<int>(h->proj_off, B, w, [](const UvsWindow &x) { return (int)x.n_proj; });
  prefix<int>(h->lobs_off, B, w, [](const UvsWindow &x) { return (int)x.n_line_obs; });
  prefix<int>(h->vobs_off, B, w, [](const UvsWindow &x) { return (int)x.n_vp_obs; });
  prefix<int>(h->imu_off, B, w, [](const UvsWindow &x) { return (int)x.n_imu; });
  prefix<int>(h->prior_off, B, w, [](const UvsWindow &x) { return (int)x.prior_n; });
  prefix<int>(h->pblk_off, B, w, [](const UvsWindow &x) { return x.prior_n > 0 ? (int)x.prior_n_blocks : 0; });
  std::vector<int> pw_off(B + 1, 0);   // warp slots of the fused point kernel: ceil(points / 32) + frames bounds the anchor-uniform warps of a window
  for (int i = 0; i < B; i++) pw_off[i + 1] = pw_off[i] + (w[i].n_points + 31) / 32 + w[i].n_frames;
  const int nPW = pw_off[B];
  h->cam_off.assign(B + 1, 0); h->S_off.assign(B + 1, 0); h->priorJ_off.assign(B + 1, 0); h->win_flags.assign(B, 0);
  for (int i = 0; i < B; i++) {
    const int d = 15 * w[i].n_frames + (w[i].estimate_extrinsic ? 6 : 0) + (td ? 1 : 0);
    h->cam_off[i + 1] = h->cam_off[i] + d;
    h->S_off[i + 1] = h->S_off[i] + (long long)d * d;
    h->priorJ_off[i + 1] = h->priorJ_off[i] + (long long)w[i].prior_n * w[i].prior_n;
    h->win_flags[i] = (w[i].estimate_extrinsic ? WF_EXTRINSIC : 0) | (td ? WF_TD : 0);
  }
  const int nF = h->frame_off[B], nP = h->point_off[B], nL = h->line_off[B], nProj = h->proj_off[B], nLobs = h->lobs_off[B],
            nVobs = h->vobs_off[B], nImu = h->imu_off[B], nCam = h->cam_off[B], nPriorR = h->prior_off[B], nBlk = h->pblk_off[B];
  const long long nS = h->S_off[B], nPJ = h->priorJ_off[B];

  // ---- input region: identical offsets in the pinned staging buffer and in the device arena
  Layout in;
  const size_t I = sizeof(int), Dd = sizeof(double);
  const size_t o_frame_off = in.take((B + 1) * I, 16), o_point_off = in.take((B + 1) * I, 16), o_line_off = in.take((B + 1) * I, 16),
               o_proj_off = in.take((B + 1) * I, 16), o_lobs_off = in.take((B + 1) * I, 16), o_vobs_off = in.take((B + 1) * I, 16),
               o_imu_off = in.take((B + 1) * I, 16), o_cam_off = in.take((B + 1) * I, 16), o_prior_off = in.take((B + 1) * I, 16),
               o_pblk_off = in.take((B + 1) * I, 16), o_S_off = in.take((B + 1) * 8, 16), o_pJ_off = in.take((B + 1) * 8, 16),
               o_flags = in.take(B * I, 16), o_frwin = in.take(nF * I, 16), o_pwoff = in.take((B + 1) * I, 16);
  const size_t o_pose = in.take(nF * 7 * Dd), o_sb = in.take(nF * 9 * Dd, 16), o_ex = in.take(B * 7 * Dd, 16), o_td = in.take(B * Dd, 16),
               o_inv = in.take(nP * Dd, 16), o_ortho = in.take(nL * 4 * Dd, 16);
  const size_t o_state_end = in.total;
  const size_t o_pfi = in.take(nProj * I), o_pfj = in.take(nProj * I), o_ppt = in.take(nProj * I),
               o_ppi = in.take(nProj * 3 * Dd), o_ppj = in.take(nProj * 3 * Dd);
  const size_t o_pvi = in.take(td ? nProj * 2 * Dd : 0), o_pvj = in.take(td ? nProj * 2 * Dd : 0),
               o_ptdi = in.take(td ? nProj * Dd : 0), o_ptdj = in.take(td ? nProj * Dd : 0),
               o_prwi = in.take(td ? nProj * Dd : 0), o_prwj = in.take(td ? nProj * Dd : 0);
  const size_t o_lf = in.take(nLobs * I), o_li = in.take(nLobs * I), o_lsp = in.take(nLobs * 2 * Dd), o_lep = in.take(nLobs * 2 * Dd);
  const size_t o_vf = in.take(nVobs * I), o_vl = in.take(nVobs * I), o_vd = in.take(nVobs * 3 * Dd);
  const size_t o_ric = in.take(B * 9 * Dd), o_tic = in.take(B * 3 * Dd);
  const size_t o_if = in.take(nImu * I), o_idp = in.take(nImu * 3 * Dd), o_idq = in.take(nImu * 4 * Dd), o_idv = in.take(nImu * 3 * Dd),
               o_idt = in.take(nImu * Dd), o_iba = in.take(nImu * 3 * Dd), o_ibg = in.take(nImu * 3 * Dd),
               o_ijac = in.take((size_t)nImu * 225 * Dd), o_icov = in.take((size_t)nImu * 225 * Dd);
  const size_t o_prJ = in.take((size_t)nPJ * Dd), o_prr = in.take(nPriorR * Dd), o_prx = in.take((size_t)nBlk * 9 * Dd),
               o_bk = in.take(nBlk * I, 16), o_bi = in.take(nBlk * I, 16), o_bcol = in.take(nBlk * I, 16), o_bcam = in.take(nBlk * I, 16),
               o_brow = in.take(nBlk * I, 16);
  in.take(0);   // the region ends aligned
  h->input_bytes = in.total;

  // ---- work region
  Layout wk;
  wk.total = in.total;
  // candidate state buffer: the same relative layout as the input state sections (filled by ONE device copy of that range)
  const size_t w_pose = wk.take(nF * 7 * Dd), w_sb = wk.take(nF * 9 * Dd, 16), w_ex = wk.take(B * 7 * Dd, 16), w_td = wk.take(B * Dd, 16),
               w_inv = wk.take(nP * Dd, 16), w_ortho = wk.take(nL * 4 * Dd, 16);
  const size_t w_pristine = wk.take(o_state_end - o_pose);
  const size_t w_cur = wk.take(B * I), w_ctl = wk.take(B * sizeof(WinCtl)), w_sum = wk.take((size_t)B * sizeof(UvsSummary));
  const size_t w_pidx = wk.take(nProj * sizeof(int4)), w_lidx = wk.take(nLobs * sizeof(int4)), w_vidx = wk.take(nVobs * sizeof(int4)),
               w_iidx = wk.take(nImu * sizeof(int2));
  const size_t w_ptb = wk.take(nP * I), w_lnb = wk.take(nL * I);            // begin arrays (memset 0x7f together)
  const size_t w_pte = wk.take(nP * I), w_lne = wk.take(nL * I), w_ptw = wk.take(nP * I), w_lnw = wk.take(nL * I);
  const size_t w_pto = wk.take((size_t)nPW * 32 * I), w_ptk = wk.take(nP * I);
  const size_t w_ft0 = wk.take((size_t)nF * 48 * Dd), w_ft1 = wk.take((size_t)nF * 48 * Dd), w_ls0 = wk.take((size_t)nL * 8 * Dd), w_ls1 = wk.take((size_t)nL * 8 * Dd);
  const size_t w_icomp = wk.take((size_t)nImu * (108 + 16) * Dd);   // compact Jacobian blocks + unweighted residual (uvs_sweep.cu IMU_CS)
  const size_t w_clw = wk.take(h->use_build3 ? (size_t)B * chol_chain_lw_doubles(max_frames) * Dd : 0);
  const size_t w_cfrag = wk.take(max_d > h->packed_limit ? (size_t)B * chol_frag_doubles(max_d) * Dd : 0);
  const size_t w_sqi = wk.take((size_t)nImu * 225 * Dd), w_prH = wk.take((size_t)nPJ * Dd), w_err = wk.take(I);
  const size_t w_rp = wk.take((size_t)nProj * 48 * Dd), w_rl = wk.take((size_t)nLobs * 24 * Dd), w_rv = wk.take((size_t)nVobs * 12 * Dd),
               w_ri = wk.take((size_t)nImu * REC_IMU * Dd), w_rpr = wk.take(nPriorR * Dd);
  const size_t w_scc = wk.take(nCam * Dd), w_scp = wk.take(nP * Dd), w_scl = wk.take(nL * 4 * Dd);
  // [S | gS | gfull | colsq | acc]: ONE contiguous block = one all-reduce per iteration in the factor-parallel mode
  const size_t w_S = wk.take((size_t)nS * Dd), w_gS = wk.take(nCam * Dd), w_gf = wk.take(nCam * Dd), w_csq = wk.take(nCam * Dd);
  const size_t w_acc = wk.take((size_t)B * ACC_STRIDE * Dd);
  const size_t w_reduce_end = wk.total;
  const size_t w_dc = wk.take(nCam * Dd), w_dp = wk.take(nP * Dd), w_dl = wk.take(nL * 4 * Dd);
  size_t w_b3 = 0;
  if (h->use_build3) {
    Dev tmp{}; tmp.B = B; tmp.nP = nP; tmp.nL = nL; tmp.nProj = nProj; tmp.nLobs = nLobs; tmp.nVobs = nVobs;
    w_b3 = wk.take(build3_bytes(tmp, max_frames, any_ex, &h->b3));
    if (build3_smem(max_frames, any_ex, max_prior_n) > h->smem_optin - 1024) h->use_build3 = false;
  }

  CK(h->stage.reserve(in.total));
  CK(h->dev.reserve(wk.total));
  if (trace) { std::fprintf(stderr, "[uvs] upload: layout %.3f ms (input %.1f MB, work %.1f MB)\n", ms_since(t_phase), in.total / 1e6, wk.total / 1e6); t_phase = now(); }
  char *S = h->stage.base;
  auto cpI = [&](size_t off, const std::vector<int> &v) { std::memcpy(S + off, v.data(), v.size() * sizeof(int)); };
  cpI(o_frame_off, h->frame_off); cpI(o_point_off, h->point_off); cpI(o_line_off, h->line_off); cpI(o_proj_off, h->proj_off);
  cpI(o_lobs_off, h->lobs_off); cpI(o_vobs_off, h->vobs_off); cpI(o_imu_off, h->imu_off); cpI(o_cam_off, h->cam_off);
  cpI(o_pwoff, pw_off); cpI(o_prior_off, h->prior_off); cpI(o_pblk_off, h->pblk_off); cpI(o_flags, h->win_flags);
  std::memcpy(S + o_S_off, h->S_off.data(), (B + 1) * 8);
  std::memcpy(S + o_pJ_off, h->priorJ_off.data(), (B + 1) * 8);
  auto put = [&](size_t off, size_t elem_off, const void *src, size_t bytes) { if (bytes && src) std::memcpy(S + off + elem_off, src, bytes); };
  std::atomic<int> pack_err(0), chain_bad(0), line_run_max(0);
  const bool want_line_runs = h->use_build3 && !any_ex;
  auto pack_range = [&](int lo, int hi) {
  for (int i = lo; i < hi; i++) {
    const UvsWindow &x = w[i];
    const size_t f0 = h->frame_off[i], p0 = h->point_off[i], l0 = h->line_off[i], j0 = h->proj_off[i], a0 = h->lobs_off[i],
                 v0 = h->vobs_off[i], m0 = h->imu_off[i], b0 = h->pblk_off[i];
    { int *fw = (int *)(S + o_frwin) + f0; for (int k = 0; k < x.n_frames; k++) fw[k] = i; }
    put(o_pose, f0 * 7 * Dd, x.pose, x.n_frames * 7 * Dd);
    put(o_sb, f0 * 9 * Dd, x.speed_bias, x.n_frames * 9 * Dd);
    put(o_ex, (size_t)i * 7 * Dd, x.ex_pose, 7 * Dd);
    { const double tdv = x.td ? x.td[0] : 0.0; put(o_td, (size_t)i * Dd, &tdv, Dd); }
    put(o_inv, p0 * Dd, x.inv_depth, x.n_points * Dd);
    put(o_ortho, l0 * 4 * Dd, x.ortho, x.n_lines * 4 * Dd);
    put(o_pfi, j0 * I, x.proj_frame_i, x.n_proj * I); put(o_pfj, j0 * I, x.proj_frame_j, x.n_proj * I);
    put(o_ppt, j0 * I, x.proj_point, x.n_proj * I);
    put(o_ppi, j0 * 3 * Dd, x.proj_pts_i, x.n_proj * 3 * Dd); put(o_ppj, j0 * 3 * Dd, x.proj_pts_j, x.n_proj * 3 * Dd);
    if (td) {
      put(o_pvi, j0 * 2 * Dd, x.proj_vel_i, x.n_proj * 2 * Dd); put(o_pvj, j0 * 2 * Dd, x.proj_vel_j, x.n_proj * 2 * Dd);
      put(o_ptdi, j0 * Dd, x.proj_td_i, x.n_proj * Dd); put(o_ptdj, j0 * Dd, x.proj_td_j, x.n_proj * Dd);
      put(o_prwi, j0 * Dd, x.proj_row_i, x.n_proj * Dd); put(o_prwj, j0 * Dd, x.proj_row_j, x.n_proj * Dd);
    }
    put(o_lf, a0 * I, x.line_frame, x.n_line_obs * I); put(o_li, a0 * I, x.line_idx, x.n_line_obs * I);
    put(o_lsp, a0 * 2 * Dd, x.line_sp, x.n_line_obs * 2 * Dd); put(o_lep, a0 * 2 * Dd, x.line_ep, x.n_line_obs * 2 * Dd);
    if (hook) {
      int seen = line_run_max.load();
      while (hook->line_run_max > seen && !line_run_max.compare_exchange_weak(seen, hook->line_run_max)) {}
    } else if (want_line_runs && x.n_line_obs > 0) {   // most observations of one line (they are contiguous): the fused path stages them per line
      int run = 1, best = 1;
      for (int k = 1; k < x.n_line_obs; k++) { run = x.line_idx[k] == x.line_idx[k - 1] ? run + 1 : 1; best = std::max(best, run); }
      int seen = line_run_max.load();
      while (best > seen && !line_run_max.compare_exchange_weak(seen, best)) {}
    }
    put(o_vf, v0 * I, x.vp_frame, x.n_vp_obs * I); put(o_vl, v0 * I, x.vp_line, x.n_vp_obs * I);
    put(o_vd, v0 * 3 * Dd, x.vp_dir, x.n_vp_obs * 3 * Dd);
    if (x.line_ric) put(o_ric, (size_t)i * 9 * Dd, x.line_ric, 9 * Dd); else std::memset(S + o_ric + (size_t)i * 9 * Dd, 0, 9 * Dd);
    if (x.line_tic) put(o_tic, (size_t)i * 3 * Dd, x.line_tic, 3 * Dd); else std::memset(S + o_tic + (size_t)i * 3 * Dd, 0, 3 * Dd);
    put(o_if, m0 * I, x.imu_frame_i, x.n_imu * I);
    put(o_idp, m0 * 3 * Dd, x.imu_delta_p, x.n_imu * 3 * Dd); put(o_idq, m0 * 4 * Dd, x.imu_delta_q, x.n_imu * 4 * Dd);
    put(o_idv, m0 * 3 * Dd, x.imu_delta_v, x.n_imu * 3 * Dd); put(o_idt, m0 * Dd, x.imu_sum_dt, x.n_imu * Dd);
    put(o_iba, m0 * 3 * Dd, x.imu_lin_ba, x.n_imu * 3 * Dd); put(o_ibg, m0 * 3 * Dd, x.imu_lin_bg, x.n_imu * 3 * Dd);
    put(o_ijac, m0 * 225 * Dd, x.imu_jacobian, (size_t)x.n_imu * 225 * Dd);
    put(o_icov, m0 * 225 * Dd, x.imu_covariance, (size_t)x.n_imu * 225 * Dd);
    if (x.prior_n > 0) {
      if (x.prior_J) put(o_prJ, (size_t)h->priorJ_off[i] * Dd, x.prior_J, (size_t)x.prior_n * x.prior_n * Dd);
      put(o_prr, (size_t)h->prior_off[i] * Dd, x.prior_r, x.prior_n * Dd);
      int col = 0; size_t xo = 0;
      int sb_lo = 1 << 30, sb_hi = -1;   // frames whose speed-bias block the prior couples
      int *bk = (int *)(S + o_bk) + b0, *bi = (int *)(S + o_bi) + b0, *bcol = (int *)(S + o_bcol) + b0,
          *bcam = (int *)(S + o_bcam) + b0, *brow = (int *)(S + o_brow) + b0;
      double *bx = (double *)(S + o_prx) + 9 * b0;
      for (int b = 0; b < x.prior_n_blocks; b++) {
        const int kind = x.prior_block_kind[b], id = x.prior_block_id[b];
        if (kind < 0 || kind > 3) { pack_err = 1; return; }
        if ((kind <= 1) && (id < 0 || id >= x.n_frames)) { pack_err = 2; return; }
        bk[b] = kind; bi[b] = id; bcol[b] = col;
        if (kind == UVS_BLOCK_SPEEDBIAS) { sb_lo = std::min(sb_lo, id); sb_hi = std::max(sb_hi, id); }
        int cam = -1, row = i;
        if (kind == UVS_BLOCK_POSE) { cam = 15 * id; row = (int)f0 + id; }
        else if (kind == UVS_BLOCK_SPEEDBIAS) { cam = 15 * id + 6; row = (int)f0 + id; }
        else if (kind == UVS_BLOCK_EXPOSE) cam = x.estimate_extrinsic ? 15 * x.n_frames : -1;
        else cam = td ? 15 * x.n_frames + (x.estimate_extrinsic ? 6 : 0) : -1;
        bcam[b] = cam; brow[b] = row;
        const int gs = prior_global(kind);
        for (int k = 0; k < 9; k++) bx[9 * b + k] = k < gs ? x.prior_x0[xo + k] : 0.0;
        xo += gs; col += prior_local(kind);
      }
      if (col != x.prior_n) { pack_err = 3; return; }
      if (sb_hi - sb_lo > 1) chain_bad = 1;   // the prior ties speed-bias blocks of non-adjacent frames together
    }
  }
  };
  {
    // the copies are independent per window: spread them over host threads for large batches
    // pack threads: at most 16, and only this process's share of the host cores when several ranks run on one node
    // (LOCAL_WORLD_SIZE of torchrun / UVS_PACK_THREADS): N ranks x 16 threads on a 32-core host was what held the
    // end-to-end scaling at 0.60 on 8 GPUs in round 1
    int nt = 1;
    if (B >= 64) {
      static const int share = [] {
        const char *e = std::getenv("UVS_PACK_THREADS");
        if (e && std::atoi(e) > 0) return std::atoi(e);
        const char *lw = std::getenv("LOCAL_WORLD_SIZE");
        const int ranks = lw && std::atoi(lw) > 0 ? std::atoi(lw) : 1;
        const int hc = (int)std::max(1u, std::thread::hardware_concurrency());
        return std::max(1, std::min(16, hc / ranks));
      }();
      nt = share;
    }
    if (nt <= 1) pack_range(0, B);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back(pack_range, (int)((long long)B * t / nt), (int)((long long)B * (t + 1) / nt));
      for (auto &t : th) t.join();
    }
    {
      static const bool no_chain = std::getenv("UVS_NO_CHAIN") != nullptr;
      bool small = false;
      for (int i = 0; i < B; i++) small = small || w[i].n_frames < 2;
      h->chain_ok = h->use_build3 && !chain_bad && !small && !no_chain;
      // fused linearisation (no point / line Jacobian records): the reference's configuration - constant extrinsic, no td
      const bool no_fuse = std::getenv("UVS_NO_FUSE") != nullptr;   // read at every upload: tests switch paths
      // The fused kernels trade parallelism for traffic: a lane walks its point's whole track (ten factors = ten dependent
      // evaluations), which is what a batch that fills the GPU wants and what a handful of windows does not - there the
      // record path (a thread per factor, then 4 / 8 lanes per landmark) has the shorter critical path.  Measured per
      // 10-iteration solve, fused / record: B = 1 1.650 / 1.604, 4 1.713 / 1.637, 16 1.779 / 1.753, 64 2.259 / 2.362 ms.
      const int fuse_min = std::getenv("UVS_FUSE_MIN") ? std::atoi(std::getenv("UVS_FUSE_MIN")) : 32;
      h->fused = h->use_build3 && !any_ex && !no_fuse && B >= fuse_min && line_run_max.load() <= lin_max_line_obs() &&
                 lin_lines_smem(max_frames) + 1024 <= h->smem_optin;
    }
    if (pack_err == 1) return fail(h, UVS_ERR_INVALID_ARG, "bad prior block kind");
    if (pack_err == 2) return fail(h, UVS_ERR_INVALID_ARG, "prior block id out of range");
    if (pack_err == 3) return fail(h, UVS_ERR_INVALID_ARG, "prior blocks do not add up to prior_n");
  }
  char *Dv = h->dev.base;
  if (trace) { std::fprintf(stderr, "[uvs] upload: pack %.3f ms\n", ms_since(t_phase)); t_phase = now(); }
  if (!hook) {
    CK(cudaMemcpyAsync(Dv, S, in.total, cudaMemcpyHostToDevice, h->stream));
    h->h2d_bytes += (int64_t)in.total;
  } else {
    // host-provided sections only: tables + state, the extrinsic frozen into the line functors, the prior's block tables;
    // the factor / IMU / prior-matrix sections are written on the device from the resident store
    InputOffsets io{o_state_end, o_pfi, o_pfj, o_ppt, o_ppi, o_ppj, o_lf, o_li, o_lsp, o_lep, o_vf, o_vl, o_vd, o_ric, o_if, o_idp, o_idq, o_idv,
                    o_idt, o_iba, o_ibg, o_ijac, o_icov, o_prJ, o_prr, o_prx, in.total};
    CK(cudaMemcpyAsync(Dv, S, o_state_end, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(Dv + o_ric, S + o_ric, o_if - o_ric, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(Dv + o_prx, S + o_prx, in.total - o_prx, cudaMemcpyHostToDevice, h->stream));
    h->h2d_bytes += (int64_t)(o_state_end + (o_if - o_ric) + (in.total - o_prx));
    const int frc = hook->fill(hook->user, h, Dv, S, io);
    if (frc) return frc;
  }
  if (trace) { cudaStreamSynchronize(h->stream); std::fprintf(stderr, "[uvs] upload: H2D %.3f ms\n", ms_since(t_phase)); t_phase = now(); }
  // zero / preset the derived region that needs it
  CK(cudaMemsetAsync(Dv + w_cur, 0, w_pidx - w_cur, h->stream));          // cur, ctl, acc, summary
  CK(cudaMemsetAsync(Dv + w_ptb, 0x7f, w_pte - w_ptb, h->stream));        // pt_begin, ln_begin
  CK(cudaMemsetAsync(Dv + w_pte, 0, w_sqi - w_pte, h->stream));           // pt_end .. ln_win
  CK(cudaMemsetAsync(Dv + w_err, 0, ALIGN, h->stream));
  CK(cudaMemsetAsync(Dv + w_pto, 0xff, (size_t)nPW * 32 * I, h->stream));   // empty lanes of the point order
  CK(cudaMemsetAsync(Dv + w_scc, 0, wk.total - w_scc, h->stream));        // scales, system, deltas
  CK(cudaMemcpyAsync(Dv + w_pose, Dv + o_pose, o_state_end - o_pose, cudaMemcpyDeviceToDevice, h->stream));  // candidate buffer = copy
  CK(cudaMemcpyAsync(Dv + w_pristine, Dv + o_pose, o_state_end - o_pose, cudaMemcpyDeviceToDevice, h->stream));

  Dev &D = h->D;
  std::memset(&D, 0, sizeof(D));
  D.B = B; D.nF = nF; D.nP = nP; D.nL = nL; D.nProj = nProj; D.nLobs = nLobs; D.nVobs = nVobs; D.nImu = nImu; D.nCam = nCam;
  D.nPriorR = nPriorR; D.nPriorBlk = nBlk; D.estimate_td = td; D.rank = h->rank; D.nranks = h->nranks;
  D.max_frames = max_frames; D.max_lines = max_lines; D.max_lobs = max_lobs;
#define PI(o) ((const int *)(Dv + (o)))
#define PD(o) ((const double *)(Dv + (o)))
#define WD(o) ((double *)(Dv + (o)))
#define WI(o) ((int *)(Dv + (o)))
  D.frame_off = PI(o_frame_off); D.point_off = PI(o_point_off); D.line_off = PI(o_line_off); D.proj_off = PI(o_proj_off);
  D.lobs_off = PI(o_lobs_off); D.vobs_off = PI(o_vobs_off); D.imu_off = PI(o_imu_off); D.cam_off = PI(o_cam_off);
  D.prior_off = PI(o_prior_off); D.pblk_off = PI(o_pblk_off);
  D.S_off = (const long long *)(Dv + o_S_off); D.priorJ_off = (const long long *)(Dv + o_pJ_off); D.win_flags = PI(o_flags);
  D.pose[0] = WD(o_pose); D.sb[0] = WD(o_sb); D.ex[0] = WD(o_ex); D.td[0] = WD(o_td); D.inv_depth[0] = WD(o_inv); D.ortho[0] = WD(o_ortho);
  D.pose[1] = WD(w_pose); D.sb[1] = WD(w_sb); D.ex[1] = WD(w_ex); D.td[1] = WD(w_td); D.inv_depth[1] = WD(w_inv); D.ortho[1] = WD(w_ortho);
  D.cur = WI(w_cur); D.ctl = (WinCtl *)(Dv + w_ctl); D.acc = WD(w_acc); D.summary = (UvsSummary *)(Dv + w_sum);
  D.proj_fi = PI(o_pfi); D.proj_fj = PI(o_pfj); D.proj_pt = PI(o_ppt); D.proj_pts_i = PD(o_ppi); D.proj_pts_j = PD(o_ppj);
  D.proj_vel_i = PD(o_pvi); D.proj_vel_j = PD(o_pvj); D.proj_td_i = PD(o_ptdi); D.proj_td_j = PD(o_ptdj);
  D.proj_row_i = PD(o_prwi); D.proj_row_j = PD(o_prwj);
  D.line_frame = PI(o_lf); D.line_idx = PI(o_li); D.line_sp = PD(o_lsp); D.line_ep = PD(o_lep);
  D.vp_frame = PI(o_vf); D.vp_line = PI(o_vl); D.vp_dir = PD(o_vd); D.ric = PD(o_ric); D.tic = PD(o_tic);
  D.imu_frame = PI(o_if); D.imu_dp = PD(o_idp); D.imu_dq = PD(o_idq); D.imu_dv = PD(o_idv); D.imu_sum_dt = PD(o_idt);
  D.imu_lin_ba = PD(o_iba); D.imu_lin_bg = PD(o_ibg); D.imu_jac = PD(o_ijac); D.imu_cov = PD(o_icov);
  D.prior_J = PD(o_prJ); D.prior_r0 = PD(o_prr); D.prior_x0 = PD(o_prx); D.pblk_kind = PI(o_bk); D.pblk_id = PI(o_bi);
  D.pblk_col = WI(o_bcol); D.pblk_cam = WI(o_bcam); D.pblk_row = WI(o_brow);
  D.proj_idx = (int4 *)(Dv + w_pidx); D.line_idx4 = (int4 *)(Dv + w_lidx); D.vp_idx4 = (int4 *)(Dv + w_vidx); D.imu_idx = (int2 *)(Dv + w_iidx);
  D.pt_begin = WI(w_ptb); D.ln_begin = WI(w_lnb); D.pt_end = WI(w_pte); D.ln_end = WI(w_lne); D.pt_win = WI(w_ptw); D.ln_win = WI(w_lnw);
  D.pt_order = WI(w_pto); D.fr_win = PI(o_frwin); D.pw_off = PI(o_pwoff); D.nPW = nPW;
  D.ftab[0] = WD(w_ft0); D.ftab[1] = WD(w_ft1); D.lsc[0] = WD(w_ls0); D.lsc[1] = WD(w_ls1);
  D.imu_sqrt_info = WD(w_sqi); D.imu_comp = WD(w_icomp); D.chain_lw = WD(w_clw); D.chain_lw_stride = chol_chain_lw_doubles(max_frames);
  D.chol_frag = WD(w_cfrag); D.chol_frag_stride = chol_frag_doubles(max_d); D.prior_H = WD(w_prH); D.err = WI(w_err);
  D.rec_proj = WD(w_rp); D.rec_line = WD(w_rl); D.rec_vp = WD(w_rv); D.rec_imu = WD(w_ri); D.rec_prior = WD(w_rpr);
  D.scale_cam = WD(w_scc); D.scale_pt = WD(w_scp); D.scale_ln = WD(w_scl);
  D.Smat = WD(w_S); D.gS = WD(w_gS); D.gfull = WD(w_gf); D.colsq_cam = WD(w_csq);
  D.colsq_pt = nullptr; D.colsq_ln = nullptr;
  D.delta_cam = WD(w_dc); D.delta_pt = WD(w_dp); D.delta_ln = WD(w_dl);
#undef PI
#undef PD
#undef WD
#undef WI
  h->o_pose0 = o_pose; h->o_state_bytes = o_state_end - o_pose;
  h->in_sb = o_sb; h->in_ex = o_ex; h->in_td = o_td; h->in_inv = o_inv; h->in_ortho = o_ortho; h->in_ric = o_ric; h->in_tic = o_tic;
  h->o_pristine = w_pristine; h->o_cur = w_cur; h->cur_bytes = B * I;
  h->o_reduce = w_S; h->reduce_doubles = (w_reduce_end - w_S) / Dd;
  h->o_b3 = w_b3;

  if (trace) { cudaStreamSynchronize(h->stream); std::fprintf(stderr, "[uvs] upload: memsets + state copies %.3f ms\n", ms_since(t_phase)); t_phase = now(); }
  h->launches += launch_prep(D, h->stream);
  if (h->use_build3) {
    CK(cudaMemsetAsync(Dv + w_b3 + h->b3.o_Y, 0, h->b3.o_ph - h->b3.o_Y, h->stream));   // dense landmark columns start as zeros
    h->launches += launch_stash_init(D, Dv + w_b3, h->b3, h->stream);
    if (h->fused) h->launches += launch_prep_point_order(D, (int *)(Dv + w_ptk), h->stream);
    else h->launches += launch_build3_prep(D, Dv + w_b3, h->b3, h->any_ex, h->stream);
  }
  int rc = post_launch(h, "prep kernels");
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_active, D.err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (trace) { cudaStreamSynchronize(h->stream); std::fprintf(stderr, "[uvs] upload: prep kernels %.3f ms\n", ms_since(t_phase)); t_phase = now(); }
  return UVS_OK;
}

// waits for the upload and reports what the device-side validation found
static int upload_finish(UvsHandle *h) {
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const int err = *h->h_active;
  if (err) {
    static const char *msg[] = {"", "projection factor index out of range", "projection factors of one point are not contiguous",
                                "projection factors of one point have different anchor frames", "line factor index out of range",
                                "line factors of one line are not contiguous", "VP factor index out of range",
                                "VP factor without the line factor of the same observation", "two VP factors on one line observation",
                                "IMU factor frame out of range", "IMU covariance not invertible / not positive definite"};
    return fail(h, err == 10 ? UVS_ERR_NOT_PD : (err == 2 || err == 3 || err == 5 || err == 7 || err == 8 ? UVS_ERR_UNSUPPORTED : UVS_ERR_INVALID_ARG),
                std::string("uvs_upload_windows: ") + msg[err < 11 ? err : 0]);
  }
  h->have_window = true;
  h->records_epoch++;
  return UVS_OK;
}

}  // extern "C"

namespace uvs {
int handle_upload(UvsHandle *h, int32_t B, const UvsWindow *w, const UvsOptions *opts, const ResidentHook *hook) {
  const int rc = upload_enqueue(h, B, w, opts, hook);
  return rc ? rc : upload_finish(h);
}
int handle_fail(UvsHandle *h, int status, const std::string &msg) { return fail(h, status, msg); }
int handle_ensure_scratch(UvsHandle *h, size_t bytes) {
  if (h->scratch.reserve(bytes) != cudaSuccess) return fail(h, UVS_ERR_CUDA, "scratch allocation failed");
  return UVS_OK;
}
int handle_ensure_hscratch(UvsHandle *h, size_t bytes) {
  if (h->hscratch.reserve(bytes) != cudaSuccess) return fail(h, UVS_ERR_CUDA, "pinned scratch allocation failed");
  return UVS_OK;
}
}  // namespace uvs
namespace {
int ensure_scratch(UvsHandle *h, size_t bytes) { return handle_ensure_scratch(h, bytes); }
int ensure_hscratch(UvsHandle *h, size_t bytes) { return handle_ensure_hscratch(h, bytes); }
int all_reduce(UvsHandle *h, double *buf, size_t count) {
  if (h->nranks <= 1) return UVS_OK;
  if (h->nccl_comm) {
    const int rc = uvs::nccl_all_reduce_sum(h->nccl_comm, buf, count, h->stream);
    if (rc != 0) return fail(h, UVS_ERR_COMM, std::string("ncclAllReduce: ") + uvs::nccl_error_string(rc));
    h->collectives++;
    return UVS_OK;
  }
  if (!h->reduce) return fail(h, UVS_ERR_COMM, "multi-rank mode without a communicator");
  const int rc = h->reduce(h->reduce_user, buf, (int64_t)count, (void *)h->stream);
  if (rc != 0) return fail(h, UVS_ERR_COMM, "reduce callback failed with " + std::to_string(rc));
  h->collectives++;
  return UVS_OK;
}
}  // namespace

extern "C" {

int uvs_download_state(UvsHandle *h, int32_t B, UvsWindow *w) {
  if (!h || !w) return fail(h, UVS_ERR_INVALID_ARG, "uvs_download_state: bad arguments");
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_download_state: no window uploaded");
  if (B != h->B) return fail(h, UVS_ERR_INVALID_ARG, "uvs_download_state: batch size differs from the upload");
  CK(cudaSetDevice(h->device));
  const Dev &D = h->D;
  const size_t Dd = sizeof(double);
  // same section layout as the input state region
  Layout L;
  const size_t o_pose = L.take(D.nF * 7 * Dd), o_sb = L.take(D.nF * 9 * Dd), o_ex = L.take(D.B * 7 * Dd), o_td = L.take(D.B * Dd),
               o_inv = L.take(D.nP * Dd), o_ortho = L.take(D.nL * 4 * Dd);
  int rc = ensure_scratch(h, L.total); if (rc) return rc;
  rc = ensure_hscratch(h, L.total); if (rc) return rc;
  char *ds = h->scratch.base;
  CK(cudaMemsetAsync(ds, 0, L.total, h->stream));
  h->launches += launch_gather_state(D, (double *)(ds + o_pose), (double *)(ds + o_sb), (double *)(ds + o_ex), (double *)(ds + o_td),
                                     (double *)(ds + o_inv), (double *)(ds + o_ortho), h->stream);
  rc = post_launch(h, "gather_state"); if (rc) return rc;
  rc = all_reduce(h, (double *)ds, L.total / Dd); if (rc) return rc;
  CK(cudaMemcpyAsync(h->hscratch.base, ds, L.total, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const char *S = h->hscratch.base;
  for (int i = 0; i < B; i++) {
    UvsWindow &x = w[i];
    const int relo = (int)h->relo.size() > i ? h->relo[i] : 0;
    if (x.n_frames + relo != h->frame_off[i + 1] - h->frame_off[i] || x.n_points != h->point_off[i + 1] - h->point_off[i] ||
        x.n_lines != h->line_off[i + 1] - h->line_off[i] || (relo && !x.relo_pose))
      return fail(h, UVS_ERR_INVALID_ARG, "uvs_download_state: window sizes differ from the upload");
    if (relo) std::memcpy(x.relo_pose, S + o_pose + (size_t)(h->frame_off[i] + x.n_frames) * 7 * Dd, 7 * Dd);
    std::memcpy(x.pose, S + o_pose + (size_t)h->frame_off[i] * 7 * Dd, x.n_frames * 7 * Dd);
    std::memcpy(x.speed_bias, S + o_sb + (size_t)h->frame_off[i] * 9 * Dd, x.n_frames * 9 * Dd);
    std::memcpy(x.ex_pose, S + o_ex + (size_t)i * 7 * Dd, 7 * Dd);
    if (x.td) std::memcpy(x.td, S + o_td + (size_t)i * Dd, Dd);
    if (x.n_points) std::memcpy(x.inv_depth, S + o_inv + (size_t)h->point_off[i] * Dd, x.n_points * Dd);
    if (x.n_lines) std::memcpy(x.ortho, S + o_ortho + (size_t)h->line_off[i] * 4 * Dd, x.n_lines * 4 * Dd);
  }
  return UVS_OK;
}

int uvs_upload_state(UvsHandle *h, int32_t B, const UvsWindow *w) {
  if (!h || !w) return fail(h, UVS_ERR_INVALID_ARG, "uvs_upload_state: bad arguments");
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_upload_state: no window uploaded");
  if (B != h->B) return fail(h, UVS_ERR_INVALID_ARG, "uvs_upload_state: batch size differs from the upload");
  CK(cudaSetDevice(h->device));
  h->records_epoch++;
  CK(cudaStreamSynchronize(h->stream));   // the staging buffer may still feed an earlier copy
  const size_t Dd = sizeof(double);
  char *S = h->stage.base, *Dv = h->dev.base;
  for (int i = 0; i < B; i++) {
    const UvsWindow &x = w[i];
    const int relo = (int)h->relo.size() > i ? h->relo[i] : 0;
    if (x.n_frames + relo != h->frame_off[i + 1] - h->frame_off[i] || x.n_points != h->point_off[i + 1] - h->point_off[i] ||
        x.n_lines != h->line_off[i + 1] - h->line_off[i])
      return fail(h, UVS_ERR_INVALID_ARG, "uvs_upload_state: window sizes differ from the upload");
    if (!x.pose || !x.speed_bias || !x.ex_pose || (x.n_points && !x.inv_depth) || (x.n_lines && !x.ortho) || (relo && !x.relo_pose))
      return fail(h, UVS_ERR_INVALID_ARG, "uvs_upload_state: null state pointer");
    std::memcpy(S + h->o_pose0 + (size_t)h->frame_off[i] * 7 * Dd, x.pose, x.n_frames * 7 * Dd);
    std::memcpy(S + h->in_sb + (size_t)h->frame_off[i] * 9 * Dd, x.speed_bias, x.n_frames * 9 * Dd);
    if (relo) {
      std::memcpy(S + h->o_pose0 + (size_t)(h->frame_off[i] + x.n_frames) * 7 * Dd, x.relo_pose, 7 * Dd);
      std::memset(S + h->in_sb + (size_t)(h->frame_off[i] + x.n_frames) * 9 * Dd, 0, 9 * Dd);
    }
    std::memcpy(S + h->in_ex + (size_t)i * 7 * Dd, x.ex_pose, 7 * Dd);
    if (x.td) std::memcpy(S + h->in_td + (size_t)i * Dd, x.td, Dd);
    if (x.n_points) std::memcpy(S + h->in_inv + (size_t)h->point_off[i] * Dd, x.inv_depth, x.n_points * Dd);
    if (x.n_lines) std::memcpy(S + h->in_ortho + (size_t)h->line_off[i] * 4 * Dd, x.ortho, x.n_lines * 4 * Dd);
    if (x.line_ric) std::memcpy(S + h->in_ric + (size_t)i * 9 * Dd, x.line_ric, 9 * Dd);
    if (x.line_tic) std::memcpy(S + h->in_tic + (size_t)i * 3 * Dd, x.line_tic, 3 * Dd);
  }
  // buffer 0 becomes the current iterate of every window again; the pristine copy follows (uvs_reset_state)
  CK(cudaMemcpyAsync(Dv + h->o_pose0, S + h->o_pose0, h->o_state_bytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(Dv + h->in_ric, S + h->in_ric, (size_t)B * 9 * Dd, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(Dv + h->in_tic, S + h->in_tic, (size_t)B * 3 * Dd, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(Dv + h->o_pristine, Dv + h->o_pose0, h->o_state_bytes, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemsetAsync(Dv + h->o_cur, 0, h->cur_bytes, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UVS_OK;
}

// ---- factor sweeps through the ABI ---------------------------------------------------------------
static int eval_common(UvsHandle *h, int type, double *residuals, double *jacobians, int32_t flags) {
  if (!h) return UVS_ERR_INVALID_ARG;
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_eval_*: no window uploaded");
  h->records_epoch++;
  CK(cudaSetDevice(h->device));
  const Dev &D = h->D;
  const bool local = (flags & UVS_EVAL_LOCAL_LAYOUT) != 0, dev_out = (flags & UVS_EVAL_DEVICE_OUT) != 0;
  const bool ceres = !local;
  const int td = D.estimate_td;
  long long n = 0; int NR = 0, JD = 0, REC = 0;
  double *rec = nullptr;
  cudaStream_t st = h->stream;
  switch (type) {
    case 0: n = D.nProj; NR = 2; REC = ceres ? (td ? CREC_PROJ_TD : CREC_PROJ) : (td ? REC_PROJ_TD : REC_PROJ); rec = D.rec_proj;
      h->launches += launch_proj(D, h->P, true, ceres, 0, 0, rec, nullptr, nullptr, 0, st); break;
    case 1: n = D.nLobs; NR = 2; REC = ceres ? CREC_LINE : REC_LINE; rec = D.rec_line;
      // tangent layout: the table-based line + VP kernel the solver and the marginalization use; Ceres layout (raw qw column): k_line
      if (ceres) h->launches += launch_line(D, h->P, true, true, 0, 0, rec, nullptr, nullptr, 0, st);
      else { h->launches += launch_line_tables(D, 0, st); h->launches += launch_line_vp(D, h->P, true, 0, 0, D.rec_line, D.rec_vp, nullptr, 0, st); }
      break;
    case 2: n = D.nVobs; NR = 1; REC = ceres ? CREC_VP : REC_VP; rec = D.rec_vp;
      if (ceres || D.nLobs == 0) h->launches += launch_vp(D, h->P, true, ceres, 0, 0, rec, nullptr, nullptr, 0, st);
      else { h->launches += launch_line_tables(D, 0, st); h->launches += launch_line_vp(D, h->P, true, 0, 0, D.rec_line, D.rec_vp, nullptr, 0, st); }
      break;
    case 3: n = D.nImu; NR = 15; REC = 15 + 15 * (2 * (ceres ? 7 : 6) + 18); rec = D.rec_imu;
      h->launches += launch_imu(D, h->P, true, 0, 0, rec, nullptr, nullptr, 0, st); break;
    default: return UVS_ERR_INVALID_ARG;
  }
  int rc = post_launch(h, "uvs_eval sweep"); if (rc) return rc;
  if (n == 0) return UVS_OK;
  JD = REC - NR;
  double *dr = residuals, *dj = jacobians;
  if (!dev_out) {
    rc = ensure_scratch(h, (size_t)n * REC * sizeof(double) + 2 * ALIGN); if (rc) return rc;
    dr = (double *)h->scratch.base;
    dj = (double *)(h->scratch.base + align_up((size_t)n * NR * sizeof(double)));
  }
  if (type == 3) h->launches += launch_export_imu(rec, (int)n, ceres ? 7 : 6, residuals ? dr : nullptr, jacobians ? dj : nullptr, st);
  else h->launches += launch_split(rec, n, REC, NR, residuals ? dr : nullptr, jacobians ? dj : nullptr, st);
  rc = post_launch(h, "uvs_eval export"); if (rc) return rc;
  if (!dev_out) {
    if (residuals) CK(cudaMemcpyAsync(residuals, dr, (size_t)n * NR * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (jacobians) CK(cudaMemcpyAsync(jacobians, dj, (size_t)n * JD * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return UVS_OK;
}

int uvs_eval_proj(UvsHandle *h, double *r, double *J, int32_t flags) { return eval_common(h, 0, r, J, flags); }
int uvs_eval_line(UvsHandle *h, double *r, double *J, int32_t flags) { return eval_common(h, 1, r, J, flags); }
int uvs_eval_vp(UvsHandle *h, double *r, double *J, int32_t flags) { return eval_common(h, 2, r, J, flags); }
int uvs_eval_imu(UvsHandle *h, double *r, double *J, int32_t flags) { return eval_common(h, 3, r, J, flags); }

int uvs_eval_prior(UvsHandle *h, double *residuals, double *jacobians, int32_t flags) {
  if (!h) return UVS_ERR_INVALID_ARG;
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_eval_prior: no window uploaded");
  h->records_epoch++;
  CK(cudaSetDevice(h->device));
  const Dev &D = h->D;
  if (D.nPriorR == 0) return UVS_OK;
  const bool local = (flags & UVS_EVAL_LOCAL_LAYOUT) != 0, dev_out = (flags & UVS_EVAL_DEVICE_OUT) != 0;
  cudaStream_t st = h->stream;
  h->launches += launch_prior(D, h->max_prior_n, false, 0, 0, D.rec_prior, nullptr, 0, st);
  int rc = post_launch(h, "uvs_eval_prior"); if (rc) return rc;
  if (residuals) CK(cudaMemcpyAsync(residuals, D.rec_prior, (size_t)D.nPriorR * sizeof(double), dev_out ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (jacobians) {
    // per window: n x sum(widths)
    std::vector<long long> off(h->B + 1, 0);
    for (int i = 0; i < h->B; i++) {
      long long cols = 0;
      const int n = h->prior_off[i + 1] - h->prior_off[i];
      // widths come from the staging copy of the block kinds
      const int *bk = nullptr; (void)bk;
      off[i + 1] = off[i];
      if (n > 0) {
        // block kinds live in the pinned staging buffer at the same offset as on the device
        const int *kinds = (const int *)(h->stage.base + ((const char *)D.pblk_kind - h->dev.base));
        for (int b = h->pblk_off[i]; b < h->pblk_off[i + 1]; b++) cols += local ? prior_local(kinds[b]) : prior_global(kinds[b]);
        off[i + 1] += (long long)n * cols;
      }
    }
    const size_t jbytes = (size_t)off[h->B] * sizeof(double), obytes = align_up((h->B + 1) * sizeof(long long));
    rc = ensure_scratch(h, obytes + jbytes + ALIGN); if (rc) return rc;
    CK(cudaMemcpyAsync(h->scratch.base, off.data(), (h->B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    double *dj = dev_out ? jacobians : (double *)(h->scratch.base + obytes);
    h->launches += launch_export_prior(D, local ? 6 : 7, dj, (const long long *)h->scratch.base, st);
    rc = post_launch(h, "uvs_eval_prior export"); if (rc) return rc;
    if (!dev_out) CK(cudaMemcpyAsync(jacobians, dj, jbytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));   // `off` must outlive the H2D copy
  }
  CK(cudaStreamSynchronize(st));
  return UVS_OK;
}

static int launch_resid_sweep(UvsHandle *h, int mode, int cand, int slot) {
  const Dev &D = h->D;
  double *cost = D.acc + slot;
  cudaStream_t st = h->stream;
  const Fork *fk = h->concurrent ? &h->fork : nullptr;
  if (fk) fork_from(fk, st, 3);
  h->launches += launch_proj(D, h->P, false, false, mode, cand, nullptr, nullptr, cost, ACC_STRIDE, st);
  h->launches += launch_line_tables(D, cand, fk ? fk->aux[0] : st);   // tables of the candidate's state buffer (current ones after an accepted step)
  h->launches += launch_line_vp(D, h->P, false, mode, cand, nullptr, nullptr, cost, ACC_STRIDE, fk ? fk->aux[0] : st);
  h->launches += launch_imu(D, h->P, false, mode, cand, nullptr, nullptr, cost, ACC_STRIDE, fk ? fk->aux[1] : st);
  // residual-only: the prior residual of the CURRENT iterate (rec_prior) must survive a rejected step
  h->launches += launch_prior(D, h->max_prior_n, false, mode, cand, nullptr, cost, ACC_STRIDE, fk ? fk->aux[2] : st);
  if (fk) for (int k = 0; k < 3; k++) join_to(fk, st, k);
  return post_launch(h, "residual sweep");
}

int uvs_eval_cost(UvsHandle *h, double *cost) {
  if (!h || !cost) return fail(h, UVS_ERR_INVALID_ARG, "uvs_eval_cost: bad arguments");
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_eval_cost: no window uploaded");
  CK(cudaSetDevice(h->device));
  int rc = ensure_scratch(h, h->B * sizeof(double)); if (rc) return rc;
  h->launches += launch_copy_acc(h->D, ACC_CAND_COST, (double *)h->scratch.base, 1, h->stream);   // clear
  rc = launch_resid_sweep(h, 0, 0, ACC_CAND_COST); if (rc) return rc;
  h->launches += launch_copy_acc(h->D, ACC_CAND_COST, (double *)h->scratch.base, 1, h->stream);
  CK(cudaMemcpyAsync(cost, h->scratch.base, h->B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UVS_OK;
}

// ---- Levenberg-Marquardt + Schur solve (ceres::Solve replacement, estimator.cpp:982-994) ---------
int uvs_solve(UvsHandle *h, UvsSummary *summaries) {
  if (!h) return UVS_ERR_INVALID_ARG;
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_solve: no window uploaded");
  CK(cudaSetDevice(h->device));
  h->records_epoch++;
  const Dev &D = h->D;
  const Params &P = h->P;
  cudaStream_t st = h->stream;
  if (h->opts.max_num_iterations + 1 > UVS_MAX_ITER_LOG) return fail(h, UVS_ERR_CAPACITY, "max_num_iterations exceeds the iteration log");
  const auto t0 = std::chrono::steady_clock::now();
  CK(cudaEventRecord(h->ev_a, st));
  h->launches += launch_solve_init(D, P, st);
  h->launches += launch_line_tables(D, 0, st);   // line / VP tables of the starting state
  const bool prof = h->profiling > 0;
  const int NE = UVS_N_STAGES + 1;
  if (prof) {
    const size_t need = (size_t)NE * h->opts.max_num_iterations;
    while (h->stage_ev.size() < need) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->stage_ev.push_back(e); }
  }
  const bool check_exit = !h->opts.fixed_iterations;
  const Fork *fk = h->concurrent ? &h->fork : nullptr;
  int rc = UVS_OK, iters_run = 0;
  double *cost0 = D.acc + ACC_COST0, *costc = D.acc + ACC_CAND_COST;
#define STAGE(k) do { if (prof) CK(cudaEventRecord(h->stage_ev[(size_t)it * NE + (k)], st)); } while (0)
  // one LM iteration = the same ~20 launches (+ fork / join events) with the same arguments every time: everything that
  // changes lives on the device.  With uvs_set_graph_replay(h, 1) the sequence is captured once per upload into a CUDA
  // graph (iteration 1; iteration 0 runs eagerly and also does the one-time attribute settings) and replayed: one
  // driver call per iteration instead of ~45.  Opt-in: instantiating the graph costs ~1.7 ms per upload (measured), more
  // than a single 10-iteration solve saves (0.3 ms) - it pays when one upload is solved repeatedly.
  auto enqueue_iteration = [&](int it) -> int {
    bool merged = false;
    STAGE(0);
    if (h->fused) {
      // fused path: the point / line / VP factors are evaluated inside the landmark elimination (uvs_lin.cu); only the
      // IMU and prior sweeps still write records.  The whole linearisation is accounted to the build stage.
      STAGE(1); STAGE(2); STAGE(3); STAGE(4); STAGE(5);
      h->launches += launch_build3_fused(D, P, h->dev.base + h->o_b3, h->b3, h->max_frames, h->max_lines, h->max_prior_n, st,
                                         h->profiling < 2 ? fk : nullptr);
    } else if (fk && h->profiling == 0 && h->use_build3 && !h->any_ex && D.B < 74 && !std::getenv("UVS_TWO_STAGE")) {
      // a handful of windows: sweep and build stage as one dependency graph over the streams (uvs_build3.cu)
      h->launches += launch_record_linearisation(D, P, h->dev.base + h->o_b3, h->b3, h->max_frames, h->max_prior_n, st, fk);
      merged = true;
    } else if (fk && h->profiling < 2) {
      // the four factor-type kernels side by side; the stage events then see the sweep as one interval
      fork_from(fk, st, 3);
      h->launches += launch_proj(D, P, true, false, 1, 0, D.rec_proj, nullptr, cost0, ACC_STRIDE, st);
      h->launches += launch_line_vp(D, P, true, 1, 0, D.rec_line, D.rec_vp, cost0, ACC_STRIDE, fk->aux[0]);
      h->launches += launch_imu(D, P, true, 1, 0, D.rec_imu, nullptr, cost0, ACC_STRIDE, fk->aux[1]);
      h->launches += launch_prior(D, h->max_prior_n, true, 1, 0, D.rec_prior, cost0, ACC_STRIDE, fk->aux[2]);
      for (int k = 0; k < 3; k++) join_to(fk, st, k);
      STAGE(1); STAGE(2); STAGE(3); STAGE(4); STAGE(5);
    } else {
      h->launches += launch_proj(D, P, true, false, 1, 0, D.rec_proj, nullptr, cost0, ACC_STRIDE, st); STAGE(1);
      h->launches += launch_line_vp(D, P, true, 1, 0, D.rec_line, D.rec_vp, cost0, ACC_STRIDE, st); STAGE(2);
      STAGE(3);   // the VP factors ride in the line kernel
      h->launches += launch_imu(D, P, true, 1, 0, D.rec_imu, nullptr, cost0, ACC_STRIDE, st); STAGE(4);
      h->launches += launch_prior(D, h->max_prior_n, true, 1, 0, D.rec_prior, cost0, ACC_STRIDE, st); STAGE(5);
    }
    rc = post_launch(h, "Jacobian sweep"); if (rc) return rc;
    if (h->fused || merged) {
    } else if (h->use_build3) {
      h->launches += launch_build3(D, P, h->dev.base + h->o_b3, h->b3, h->max_frames, h->any_ex, h->max_prior_n, st, fk);
    } else {
      h->launches += launch_build(D, P, h->max_prior_n, st);
    }
    rc = post_launch(h, "build"); if (rc) return rc;
    if (h->nranks > 1) {   // partial reduced systems + per-window accumulators of all ranks: one collective
      rc = all_reduce(h, (double *)(h->dev.base + h->o_reduce), h->reduce_doubles); if (rc) return rc;
    }
    STAGE(6);
    if (h->chain_ok) h->launches += launch_chol_chain(D, P, h->max_frames, h->any_ex, h->use_build3, st);
    else h->launches += launch_chol(D, P, h->max_d, h->packed_limit, h->use_build3, st);
    STAGE(7);
    rc = post_launch(h, "chol"); if (rc) return rc;
    if (h->use_build3) h->launches += launch_back3(D, h->dev.base + h->o_b3, h->b3, st, fk, h->fused);
    else h->launches += launch_backsub(D, P, st);
    STAGE(8);
    rc = post_launch(h, "backsub"); if (rc) return rc;
    rc = launch_resid_sweep(h, 1, 1, ACC_CAND_COST); if (rc) return rc;
    (void)costc;
    if (h->nranks > 1) { rc = all_reduce(h, D.acc, (size_t)D.B * ACC_STRIDE); if (rc) return rc; }
    STAGE(9);
    h->launches += launch_step(D, P, !h->chain_ok, st); STAGE(10);
    rc = post_launch(h, "step"); if (rc) return rc;
    return UVS_OK;
  };
  const bool use_graph = h->graph_replay && !prof && h->nranks <= 1 && h->opts.max_num_iterations >= 3;
  for (int it = 0; it < h->opts.max_num_iterations; it++) {
    if (use_graph && h->iter_exec) {
      CK(cudaGraphLaunch(h->iter_exec, st));
      h->launches += h->iter_launches;
    } else if (use_graph && it == 1) {
      const int64_t l0 = h->launches;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      rc = enqueue_iteration(it);
      cudaGraph_t g = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (ce != cudaSuccess) return cuda_fail(h, ce, "graph capture of the LM iteration");
      const cudaError_t ie = cudaGraphInstantiate(&h->iter_exec, g, 0);
      cudaGraphDestroy(g);
      if (ie != cudaSuccess) { h->iter_exec = nullptr; return cuda_fail(h, ie, "graph instantiation"); }
      h->iter_launches = h->launches - l0;
      CK(cudaGraphLaunch(h->iter_exec, st));
    } else {
      rc = enqueue_iteration(it); if (rc) return rc;
    }
    iters_run++;
    if (check_exit || h->opts.max_solver_time > 0.0) {
      h->launches += launch_count_active(D, h->d_active, st);
      CK(cudaMemcpyAsync(h->h_active, h->d_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (*h->h_active == 0) break;
      if (h->opts.max_solver_time > 0.0 && !h->opts.fixed_iterations) {
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (el >= h->opts.max_solver_time) break;   // remaining windows report UVS_TERM_NO_CONVERGENCE / time
      }
    }
  }
#undef STAGE
  h->launches += launch_finish(D, st);
  CK(cudaEventRecord(h->ev_b, st));
  if (summaries) CK(cudaMemcpyAsync(summaries, D.summary, (size_t)D.B * sizeof(UvsSummary), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&h->last_solve_ms, h->ev_a, h->ev_b);
  for (int k = 0; k < UVS_N_STAGES; k++) h->stage_ms[k] = 0.f;
  h->stage_iters = prof ? iters_run : 0;
  if (prof)
    for (int it = 0; it < iters_run; it++)
      for (int k = 0; k < UVS_N_STAGES; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->stage_ev[(size_t)it * NE + k], h->stage_ev[(size_t)it * NE + k + 1]) == cudaSuccess) h->stage_ms[k] += ms;
      }
  h->last_sweep_ms = h->stage_ms[0] + h->stage_ms[1] + h->stage_ms[2] + h->stage_ms[3] + h->stage_ms[4];
  h->n_sweeps = h->stage_iters;
  if (summaries && h->opts.max_solver_time > 0.0 && !h->opts.fixed_iterations) {
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (el >= h->opts.max_solver_time)
      for (int i = 0; i < D.B; i++) if (summaries[i].termination == UVS_TERM_NO_CONVERGENCE && summaries[i].num_iterations <= h->opts.max_num_iterations) summaries[i].termination = UVS_TERM_TIME;
  }
  return UVS_OK;
}

int uvs_batch_solve(UvsHandle *h, int32_t B, UvsWindow *w, const UvsOptions *opts, UvsSummary *summaries) {
  static const bool trace = std::getenv("UVS_TRACE") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!trace) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[uvs] batch_solve: %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  };
  int rc = uvs_upload_windows(h, B, w, opts); if (rc) return rc;
  lap("upload");
  rc = uvs_solve(h, summaries); if (rc) return rc;
  lap("solve");
  rc = uvs_download_state(h, B, w);
  lap("download");
  return rc;
}

int uvs_batch_solve_pipelined(UvsHandle *h, int32_t B, UvsWindow *w, const UvsOptions *opts, UvsSummary *summaries, int32_t n_groups) {
  if (!h || B <= 0 || !w) return fail(h, UVS_ERR_INVALID_ARG, "uvs_batch_solve_pipelined: bad arguments");
  if (h->nranks > 1) return fail(h, UVS_ERR_UNSUPPORTED, "uvs_batch_solve_pipelined: not available in factor-parallel mode");
  // every sub-batch pays the fixed latency of an LM solve (~20 dependent launches per iteration) again, so few groups:
  // measured on B200, 1184 C2 windows: 1 group 27.8 ms, 2: 23.0, 3: 22.8, 4: 24.1, 6: 27.0, 8: 28.7 (tools/e2e_probe.py)
  int G = n_groups > 0 ? n_groups : (B >= 768 ? 3 : (B >= 128 ? 2 : 1));
  G = std::min(G, B);
  if (G <= 1) return uvs_batch_solve(h, B, w, opts, summaries);
  while ((int)h->children.size() < G) {
    UvsHandle *c = nullptr;
    const int rc = uvs_create(h->device, &c);
    if (rc) return fail(h, rc, "uvs_batch_solve_pipelined: cannot create a sub-batch handle");
    h->children.push_back(c);
  }
  h->have_window = false;   // one-shot service call: the parent keeps no batch
  // sub-batch boundaries: the first group gets `first` times the share of the others (the GPU idles until its data has
  // landed, so a smaller first group starts the device earlier; UVS_PIPE_FIRST overrides for experiments)
  std::vector<int> bounds(G + 1, 0);
  {
    static const char *env = std::getenv("UVS_PIPE_FIRST");
    const double first = env ? std::atof(env) : (G >= 3 ? 0.6 : 1.0);   // measured (tools/e2e_split.py): G = 3: 22.84 -> 22.44 ms, G = 2: no gain
    const double unit = (double)B / (first + (G - 1));
    for (int g = 1; g < G; g++) bounds[g] = std::min(B, std::max(g, (int)(unit * (first + (g - 1)) + 0.5)));
    bounds[G] = B;
  }
  std::vector<int> rcs(G, UVS_OK);
  std::vector<std::thread> workers;
  const UvsOptions o = opts ? *opts : h->opts;
  static const bool trace = std::getenv("UVS_TRACE") != nullptr;
  const auto t_call = std::chrono::steady_clock::now();
  auto since = [t_call] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
  for (int g = 0; g < G; g++) {
    UvsHandle *c = h->children[g];
    const int lo = bounds[g], hi = bounds[g + 1];
    // uploads go one after another (they share the host cores and the PCIe link); each sub-batch starts its LM loop
    // on its own stream as soon as its data has landed, overlapping the next sub-batch's pack + H2D
    rcs[g] = upload_enqueue(c, hi - lo, w + lo, &o);
    if (trace) std::fprintf(stderr, "[uvs] pipelined: group %d upload enqueued at %.2f ms\n", g, since());
    if (rcs[g]) break;
    workers.emplace_back([c, lo, hi, w, summaries, g, since, &rcs] {
      int rc = upload_finish(c);
      const double t_up = since();
      if (!rc) rc = uvs_solve(c, summaries ? summaries + lo : nullptr);
      const double t_solve = since();
      if (!rc) rc = uvs_download_state(c, hi - lo, w + lo);
      if (trace) std::fprintf(stderr, "[uvs] pipelined: group %d data on device at %.2f ms, solved at %.2f ms (device %.2f ms), downloaded at %.2f ms\n",
                              g, t_up, t_solve, c->last_solve_ms, since());
      rcs[g] = rc;
    });
  }
  for (auto &t : workers) t.join();
  float ms = 0.f;
  for (int g = 0; g < G; g++) {
    UvsHandle *c = h->children[g];
    ms = std::max(ms, c->last_solve_ms);
    h->launches += c->launches; c->launches = 0;
  }
  h->last_solve_ms = ms;
  for (int g = 0; g < G; g++)
    if (rcs[g]) return fail(h, rcs[g], "sub-batch " + std::to_string(g) + ": " + h->children[g]->err);
  return UVS_OK;
}

int uvs_marginalize(UvsHandle *h, int32_t window_index, int32_t flag, UvsPrior *out) {
  if (!h || !out) return fail(h, UVS_ERR_INVALID_ARG, "uvs_marginalize: bad arguments");
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_marginalize: no window uploaded");
  return uvs_marginalize_impl(h, window_index, flag, out);
}

int uvs_sweep_bytes(UvsHandle *h, int64_t *jac, int64_t *res) {
  if (!h) return UVS_ERR_INVALID_ARG;
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_sweep_bytes: no window uploaded");
  // algorithmic bytes per factor (reads of observations + indices, writes of r and local J)
  const Dev &D = h->D;
  long long pr = 0;
  for (int i = 0; i < h->B; i++) { const long long n = h->prior_off[i + 1] - h->prior_off[i]; pr += 8 * (n * n + 3 * n); }
  const long long state = 8LL * (16LL * D.nF + 8LL * D.B + D.nP + 4LL * D.nL);
  if (jac) *jac = 384LL * D.nProj + 232LL * D.nLobs + 120LL * D.nVobs + 6024LL * D.nImu + pr + state;
  if (res) *res = 80LL * D.nProj + 72LL * D.nLobs + 40LL * D.nVobs + 2424LL * D.nImu + pr + state;
  return UVS_OK;
}

int uvs_jacobian_sweep(UvsHandle *h, int32_t repeats, float *ms_group, float *ms_each) {
  if (!h || repeats <= 0) return fail(h, UVS_ERR_INVALID_ARG, "uvs_jacobian_sweep: bad arguments");
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_jacobian_sweep: no window uploaded");
  CK(cudaSetDevice(h->device));
  h->records_epoch++;
  const Dev &D = h->D;
  const Params &P = h->P;
  cudaStream_t st = h->stream;
  const Fork *fk = &h->fork;
  auto one = [&](int k, cudaStream_t s) {
    if (k == 0) h->launches += launch_proj(D, P, true, false, 0, 0, D.rec_proj, nullptr, nullptr, 0, s);
    else if (k == 1) h->launches += launch_line_vp(D, P, true, 0, 0, D.rec_line, D.rec_vp, nullptr, 0, s);
    else if (k == 2) h->launches += launch_imu(D, P, true, 0, 0, D.rec_imu, nullptr, nullptr, 0, s);
    else h->launches += launch_prior(D, h->max_prior_n, true, 0, 0, D.rec_prior, nullptr, 0, s);
  };
  h->launches += launch_line_tables(D, 0, st);
  for (int k = 0; k < 4; k++) one(k, st);   // warm-up (shared-memory attributes, instruction cache)
  CK(cudaEventRecord(h->ev_c, st));
  for (int r = 0; r < repeats; r++) {
    fork_from(fk, st, 3);
    one(0, st); one(1, fk->aux[0]); one(2, fk->aux[1]); one(3, fk->aux[2]);
    for (int k = 0; k < 3; k++) join_to(fk, st, k);
  }
  CK(cudaEventRecord(h->ev_d, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev_c, h->ev_d);
  if (ms_group) *ms_group = ms / repeats;
  if (ms_each)
    for (int k = 0; k < 4; k++) {
      CK(cudaEventRecord(h->ev_c, st));
      for (int r = 0; r < repeats; r++) one(k, st);
      CK(cudaEventRecord(h->ev_d, st));
      CK(cudaStreamSynchronize(st));
      cudaEventElapsedTime(&ms, h->ev_c, h->ev_d);
      ms_each[k] = ms / repeats;
    }
  return post_launch(h, "uvs_jacobian_sweep");
}

int uvs_reset_state(UvsHandle *h) {
  if (!h) return UVS_ERR_INVALID_ARG;
  if (!h->have_window) return fail(h, UVS_ERR_NO_WINDOW, "uvs_reset_state: no window uploaded");
  CK(cudaSetDevice(h->device));
  h->records_epoch++;
  char *Dv = h->dev.base;
  CK(cudaMemcpyAsync(Dv + h->o_pose0, Dv + h->o_pristine, h->o_state_bytes, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemsetAsync(Dv + h->o_cur, 0, h->cur_bytes, h->stream));
  return UVS_OK;
}

int uvs_set_graph_replay(UvsHandle *h, int32_t enable) {
  if (!h) return UVS_ERR_INVALID_ARG;
  h->graph_replay = enable != 0;
  if (!h->graph_replay && h->iter_exec) { cudaGraphExecDestroy(h->iter_exec); h->iter_exec = nullptr; }
  return UVS_OK;
}

int uvs_set_profiling(UvsHandle *h, int32_t level) {
  if (!h) return UVS_ERR_INVALID_ARG;
  h->profiling = level;
  return UVS_OK;
}

int uvs_last_stage_ms(const UvsHandle *h, float ms[UVS_N_STAGES], int32_t *n_iterations) {
  if (!h || !ms) return UVS_ERR_INVALID_ARG;
  for (int k = 0; k < UVS_N_STAGES; k++) ms[k] = h->stage_ms[k];
  if (n_iterations) *n_iterations = h->stage_iters;
  return UVS_OK;
}

int64_t uvs_launch_count(const UvsHandle *h) { return h ? h->launches : 0; }

int uvs_last_solve_ms(const UvsHandle *h, float *ms) {
  if (!h || !ms) return UVS_ERR_INVALID_ARG;
  *ms = h->last_solve_ms;
  return UVS_OK;
}

int uvs_last_sweep_ms(const UvsHandle *h, float *ms, int32_t *n) {
  if (!h || !ms) return UVS_ERR_INVALID_ARG;
  *ms = h->last_sweep_ms;
  if (n) *n = h->n_sweeps;
  return UVS_OK;
}

int uvs_comm_unique_id(unsigned char *id) {
  if (!id) return UVS_ERR_INVALID_ARG;
  return uvs::nccl_unique_id(id) == 0 ? UVS_OK : UVS_ERR_COMM;
}

int uvs_comm_init_nccl(UvsHandle *h, const unsigned char *id, int32_t rank, int32_t nranks) {
  if (!h || !id || nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) return fail(h, UVS_ERR_INVALID_ARG, "uvs_comm_init_nccl: bad arguments");
  CK(cudaSetDevice(h->device));
  if (h->iter_exec) { cudaGraphExecDestroy(h->iter_exec); h->iter_exec = nullptr; }
  if (h->nccl_comm) { uvs::nccl_destroy(h->nccl_comm); h->nccl_comm = nullptr; }
  if (nranks > 1) {
    const int rc = uvs::nccl_init_rank(&h->nccl_comm, id, rank, nranks);
    if (rc != 0) return fail(h, UVS_ERR_COMM, std::string("ncclCommInitRank: ") + uvs::nccl_error_string(rc));
  }
  h->rank = rank; h->nranks = nranks; h->reduce = nullptr; h->reduce_user = nullptr;
  h->D.rank = rank; h->D.nranks = nranks;
  return UVS_OK;
}

int64_t uvs_collective_count(const UvsHandle *h) { return h ? h->collectives : 0; }

int uvs_comm_init(UvsHandle *h, int32_t rank, int32_t nranks, UvsAllReduceFn reduce, void *user) {
  if (!h || nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) return fail(h, UVS_ERR_INVALID_ARG, "uvs_comm_init: bad arguments");
  if (nranks > 1 && !reduce) return fail(h, UVS_ERR_INVALID_ARG, "uvs_comm_init: reduce callback missing");
  if (h->iter_exec) { cudaGraphExecDestroy(h->iter_exec); h->iter_exec = nullptr; }
  h->rank = rank; h->nranks = nranks; h->reduce = reduce; h->reduce_user = user;
  h->D.rank = rank; h->D.nranks = nranks;
  return UVS_OK;
}

}  // extern "C"
