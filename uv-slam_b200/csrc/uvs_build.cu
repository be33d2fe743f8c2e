// uvs_build.cu — normal equations + Schur complement + back-substitution.
//
// Replaces what Ceres does inside ceres::Solve for linear_solver_type = SPARSE_SCHUR
// (vins_estimator/src/estimator.cpp:984): block J^T J, elimination of the landmark blocks (inverse
// depths 1x1, orthonormal lines 4x4) into the reduced camera system, and the landmark
// back-substitution.  Jacobi scaling and the LM diagonal follow Ceres' documented behaviour
// (SURVEY.md 8c / Appendix B).
//
// One warp owns one landmark: lane f holds the record of the landmark's f-th factor (a landmark's
// records are contiguous), the 1x1 / 4x4 landmark block and the coupling blocks live in registers,
// warp shuffles do the reductions over the landmark's factors ("warp-reduced block elimination"),
// and the warp adds its finished contribution  H_cc - W (E + D^2)^-1 W^T  to the window's reduced
// system with FP64 reductions (RED.ADD.F64 at L2).  Camera-side Jacobi scaling is applied later
// (k_chol); landmark-side scaling is applied here.
#include "uvs_device.cuh"
#include "uvs_kernels.h"

namespace uvs {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ void addS(double *S, int d, int r, int c, double v) {
  if (r <= c) atomicAdd(S + (size_t)r * d + c, v);
  else atomicAdd(S + (size_t)c * d + r, v);
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

struct WinView {
  int w, d, F, ex_off, td_off;
  double *S, *gS, *gfull, *colsq;
  const double *delta;
  double radius;
  bool have_scale;
};

__device__ __forceinline__ WinView window_view(const Dev &D, int w) {
  WinView v;
  v.w = w;
  const int co = D.cam_off[w];
  v.d = D.cam_off[w + 1] - co;
  v.F = D.frame_off[w + 1] - D.frame_off[w];
  const int fl = D.win_flags[w];
  v.ex_off = (fl & WF_EXTRINSIC) ? 15 * v.F : -1;
  v.td_off = (fl & WF_TD) ? 15 * v.F + ((fl & WF_EXTRINSIC) ? 6 : 0) : -1;
  v.S = D.Smat + D.S_off[w];
  v.gS = D.gS + co; v.gfull = D.gfull + co; v.colsq = D.colsq_cam + co;
  v.delta = D.delta_cam + co;
  v.radius = D.ctl[w].radius;
  v.have_scale = D.ctl[w].have_scale != 0;
  return v;
}

// ------------------------------------------------------------------------------------------------
// Points.  Per lane: one projection factor record.  "Common" camera columns are shared by all
// factors of the point (anchor pose, extrinsic, td); the "own" block is the observing pose.
template <bool kEx, bool kTd>
struct PointLane {
  static constexpr int NC = 6 + (kEx ? 6 : 0) + (kTd ? 1 : 0);
  double r[2], Ac[2][NC], Jo[2][6], Jl[2];
  int coff[NC];   // camera offset of every common column
  int cj;         // camera offset of the own block
};

template <bool kEx, bool kTd>
__device__ __forceinline__ void load_point_lane(const Dev &D, const WinView &V, int f, bool valid, PointLane<kEx, kTd> &L) {
  constexpr int NC = PointLane<kEx, kTd>::NC;
  const int REC = D.estimate_td ? REC_PROJ_TD : REC_PROJ;
  if (valid) {
    const double *R = D.rec_proj + (size_t)f * REC;
    const int4 ix = D.proj_idx[f];
    const int fo = D.frame_off[V.w];
    const int ci = 15 * (ix.x - fo);
    L.cj = 15 * (ix.y - fo);
    L.r[0] = R[0]; L.r[1] = R[1];
#pragma unroll
    for (int row = 0; row < 2; row++) {
#pragma unroll
      for (int c = 0; c < 6; c++) {
        L.Ac[row][c] = R[2 + row * 6 + c];
        L.Jo[row][c] = R[14 + row * 6 + c];
        if (kEx) L.Ac[row][6 + c] = R[26 + row * 6 + c];
      }
      L.Jl[row] = R[38 + row];
      if (kTd) L.Ac[row][NC - 1] = R[40 + row];
    }
#pragma unroll
    for (int c = 0; c < 6; c++) { L.coff[c] = ci + c; if (kEx) L.coff[6 + c] = V.ex_off + c; }
    if (kTd) L.coff[NC - 1] = V.td_off;
  } else {
    L.r[0] = L.r[1] = 0.0; L.Jl[0] = L.Jl[1] = 0.0; L.cj = 0;
#pragma unroll
    for (int row = 0; row < 2; row++) {
#pragma unroll
      for (int c = 0; c < NC; c++) L.Ac[row][c] = 0.0;
#pragma unroll
      for (int c = 0; c < 6; c++) L.Jo[row][c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < NC; c++) L.coff[c] = 0;
  }
  // common offsets are the same for every valid lane: take lane 0's
#pragma unroll
  for (int c = 0; c < NC; c++) L.coff[c] = __shfl_sync(FULL, L.coff[c], 0);
}

// landmark scalars shared by build and back-substitution
struct PointCore {
  double sk, hinv, gk;   // Jacobi scale, 1/(E~ + D^2), unscaled landmark gradient
};

template <bool kEx, bool kTd>
__device__ __forceinline__ PointCore point_core(const Dev &D, const Params &P, const WinView &V, int gp,
                                                const PointLane<kEx, kTd> &L, int lane) {
  PointCore c;
  const double colsq = wsum(L.Jl[0] * L.Jl[0] + L.Jl[1] * L.Jl[1]);
  c.gk = wsum(L.Jl[0] * L.r[0] + L.Jl[1] * L.r[1]);
  if (!V.have_scale) {
    c.sk = 1.0 / (1.0 + sqrt(colsq));
    if (lane == 0) D.scale_pt[gp] = c.sk;
  } else {
    c.sk = D.scale_pt[gp];
  }
  const double Et = c.sk * c.sk * colsq;
  c.hinv = 1.0 / (Et + clampd(Et, P.min_lm_diag, P.max_lm_diag) / V.radius);
  return c;
}

template <bool kEx, bool kTd>
__device__ void build_point(const Dev &D, const Params &P, const WinView &V, int gp, int f0, int n, int lane) {
  constexpr int NC = PointLane<kEx, kTd>::NC;
  PointLane<kEx, kTd> L;
  const bool valid = lane < n;
  load_point_lane<kEx, kTd>(D, V, f0 + lane, valid, L);
  const PointCore pc = point_core<kEx, kTd>(D, P, V, gp, L, lane);
  const double hinv = pc.hinv, gkt = pc.sk * pc.gk;
  if (lane == 0) atomic_max_nonneg(D.acc + (size_t)V.w * ACC_STRIDE + ACC_GMAX + D.rank, fabs(pc.gk));
  const int d = V.d;

  // coupling vectors W = H_ck D_sk (camera side unscaled)
  double wc[NC], wo[6];
#pragma unroll
  for (int c = 0; c < NC; c++) wc[c] = pc.sk * wsum(L.Ac[0][c] * L.Jl[0] + L.Ac[1][c] * L.Jl[1]);
#pragma unroll
  for (int c = 0; c < 6; c++) wo[c] = pc.sk * (L.Jo[0][c] * L.Jl[0] + L.Jo[1][c] * L.Jl[1]);

  // gradient and squared column norms, common columns (warp-reduced, one lane commits)
#pragma unroll
  for (int c = 0; c < NC; c++) {
    const double g = wsum(L.Ac[0][c] * L.r[0] + L.Ac[1][c] * L.r[1]);
    const double q = wsum(L.Ac[0][c] * L.Ac[0][c] + L.Ac[1][c] * L.Ac[1][c]);
    if (lane == (c & 31)) {
      atomicAdd(V.gfull + L.coff[c], g);
      atomicAdd(V.gS + L.coff[c], g - wc[c] * hinv * gkt);
      atomicAdd(V.colsq + L.coff[c], q);
    }
  }
  // own columns
  if (valid) {
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const double g = L.Jo[0][c] * L.r[0] + L.Jo[1][c] * L.r[1];
      atomicAdd(V.gfull + L.cj + c, g);
      atomicAdd(V.gS + L.cj + c, g - wo[c] * hinv * gkt);
      atomicAdd(V.colsq + L.cj + c, L.Jo[0][c] * L.Jo[0][c] + L.Jo[1][c] * L.Jo[1][c]);
    }
  }
  // common x common (upper triangle; offsets ascend: anchor pose < extrinsic < td)
  {
    int idx = 0;
#pragma unroll
    for (int a = 0; a < NC; a++) {
#pragma unroll
      for (int b = a; b < NC; b++) {
        const double h = wsum(L.Ac[0][a] * L.Ac[0][b] + L.Ac[1][a] * L.Ac[1][b]) - wc[a] * wc[b] * hinv;
        if (lane == (idx & 31)) addS(V.S, d, L.coff[a], L.coff[b], h);
        idx++;
      }
    }
  }
  if (valid) {
    // common x own
#pragma unroll
    for (int a = 0; a < NC; a++) {
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const double h = L.Ac[0][a] * L.Jo[0][c] + L.Ac[1][a] * L.Jo[1][c] - wc[a] * wo[c] * hinv;
        addS(V.S, d, L.coff[a], L.cj + c, h);
      }
    }
    // own x own (same factor)
#pragma unroll
    for (int p = 0; p < 6; p++) {
#pragma unroll
      for (int q = p; q < 6; q++) {
        const double h = L.Jo[0][p] * L.Jo[0][q] + L.Jo[1][p] * L.Jo[1][q] - wo[p] * wo[q] * hinv;
        atomicAdd(V.S + (size_t)(L.cj + p) * d + L.cj + q, h);
      }
    }
  }
  // own x own' (different observing frames): pure Schur term
  for (int step = 1; step < n; step++) {
    double w2[6];
#pragma unroll
    for (int c = 0; c < 6; c++) w2[c] = __shfl_down_sync(FULL, wo[c], step);
    const int cj2 = __shfl_down_sync(FULL, L.cj, step);
    if (lane + step < n) {
#pragma unroll
      for (int p = 0; p < 6; p++) {
        const double t = -wo[p] * hinv;
#pragma unroll
        for (int q = 0; q < 6; q++) addS(V.S, d, L.cj + p, cj2 + q, t * w2[q]);
      }
    }
  }
}

template <bool kEx, bool kTd>
__device__ void backsub_point(const Dev &D, const Params &P, const WinView &V, int gp, int f0, int n, int lane) {
  constexpr int NC = PointLane<kEx, kTd>::NC;
  PointLane<kEx, kTd> L;
  const bool valid = lane < n;
  load_point_lane<kEx, kTd>(D, V, f0 + lane, valid, L);
  const PointCore pc = point_core<kEx, kTd>(D, P, V, gp, L, lane);
  double u[2] = {0.0, 0.0};
  if (valid) {
#pragma unroll
    for (int c = 0; c < NC; c++) { const double dc = V.delta[L.coff[c]]; u[0] += L.Ac[0][c] * dc; u[1] += L.Ac[1][c] * dc; }
#pragma unroll
    for (int c = 0; c < 6; c++) { const double dc = V.delta[L.cj + c]; u[0] += L.Jo[0][c] * dc; u[1] += L.Jo[1][c] * dc; }
  }
  const double t = pc.sk * (pc.gk + wsum(L.Jl[0] * u[0] + L.Jl[1] * u[1]));
  const double dk = pc.sk * (-pc.hinv * t);
  const double jd0 = u[0] + L.Jl[0] * dk, jd1 = u[1] + L.Jl[1] * dk;
  const double mc = wsum(jd0 * (L.r[0] + 0.5 * jd0) + jd1 * (L.r[1] + 0.5 * jd1));
  if (lane == 0) {
    const int w = V.w;
    const int cur = D.cur[w];
    const double lam = D.inv_depth[cur][gp];
    D.delta_pt[gp] = dk;
    D.inv_depth[cur ^ 1][gp] = lam + dk;
    double *acc = D.acc + (size_t)w * ACC_STRIDE;
    atomicAdd(acc + ACC_MODEL, -mc);
    atomicAdd(acc + ACC_STEP2, dk * dk);
    atomicAdd(acc + ACC_XNORM2, lam * lam);
  }
}

template <bool kBack>
__global__ void __launch_bounds__(128) k_points(Dev D, Params P) {
  const int gp = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gp >= D.nP) return;
  if (D.nranks > 1 && (gp % D.nranks) != D.rank) return;
  const int w = D.pt_win[gp];
  const int st = D.ctl[w].state;
  if (!(st & WS_ACTIVE)) return;
  const WinView V = window_view(D, w);
  const int f0 = D.pt_begin[gp], n = D.pt_end[gp] - f0;
  if (kBack) {
    if (D.acc[(size_t)w * ACC_STRIDE + ACC_FAIL] != 0.0) return;
    if (n <= 0) {  // unobserved landmark: gradient 0 -> step 0
      if (lane == 0) { const int cur = D.cur[w]; D.delta_pt[gp] = 0.0; D.inv_depth[cur ^ 1][gp] = D.inv_depth[cur][gp]; }
      return;
    }
  } else if (n <= 0) return;
  const int fl = D.win_flags[w];
  const bool ex = fl & WF_EXTRINSIC, td = fl & WF_TD;
#define UVS_DISPATCH(FN)                                            \
  if (ex) { if (td) FN<true, true>(D, P, V, gp, f0, n, lane); else FN<true, false>(D, P, V, gp, f0, n, lane); } \
  else { if (td) FN<false, true>(D, P, V, gp, f0, n, lane); else FN<false, false>(D, P, V, gp, f0, n, lane); }
  if (kBack) { UVS_DISPATCH(backsub_point) } else { UVS_DISPATCH(build_point) }
#undef UVS_DISPATCH
}

// ------------------------------------------------------------------------------------------------
// Lines.  Per lane: one line observation (2 rows) plus the VP factor of the same observation (1 row).
struct LineLane {
  double r[3], Jp[3][6], Jl[3][4];
  int cj;
};

__device__ __forceinline__ void load_line_lane(const Dev &D, const WinView &V, int f, bool valid, LineLane &L) {
#pragma unroll
  for (int row = 0; row < 3; row++) {
    L.r[row] = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) L.Jp[row][c] = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++) L.Jl[row][c] = 0.0;
  }
  L.cj = 0;
  if (!valid) return;
  const int4 ix = D.line_idx4[f];
  L.cj = 15 * (ix.x - D.frame_off[V.w]);
  const double *R = D.rec_line + (size_t)f * REC_LINE;
  L.r[0] = R[0]; L.r[1] = R[1];
#pragma unroll
  for (int row = 0; row < 2; row++) {
#pragma unroll
    for (int c = 0; c < 6; c++) L.Jp[row][c] = R[2 + row * 6 + c];
#pragma unroll
    for (int c = 0; c < 4; c++) L.Jl[row][c] = R[14 + row * 4 + c];
  }
  if (ix.w >= 0) {
    const double *Q = D.rec_vp + (size_t)ix.w * REC_VP;
    L.r[2] = Q[0];
#pragma unroll
    for (int c = 0; c < 6; c++) L.Jp[2][c] = Q[1 + c];
#pragma unroll
    for (int c = 0; c < 4; c++) L.Jl[2][c] = Q[7 + c];
  }
}

struct LineCore {
  double s[4], Minv[16], gl[4];   // Jacobi scales, (E~ + D^2)^-1, unscaled landmark gradient
  bool ok;
};

__device__ __forceinline__ LineCore line_core(const Dev &D, const Params &P, const WinView &V, int gl_idx,
                                              const LineLane &L, int lane) {
  LineCore c;
  double E[16];
#pragma unroll
  for (int p = 0; p < 4; p++) {
#pragma unroll
    for (int q = p; q < 4; q++) {
      const double e = wsum(L.Jl[0][p] * L.Jl[0][q] + L.Jl[1][p] * L.Jl[1][q] + L.Jl[2][p] * L.Jl[2][q]);
      E[4 * p + q] = e; E[4 * q + p] = e;
    }
    c.gl[p] = wsum(L.Jl[0][p] * L.r[0] + L.Jl[1][p] * L.r[1] + L.Jl[2][p] * L.r[2]);
  }
  if (!V.have_scale) {
#pragma unroll
    for (int p = 0; p < 4; p++) c.s[p] = 1.0 / (1.0 + sqrt(E[5 * p]));
    if (lane < 4) D.scale_ln[4 * (size_t)gl_idx + lane] = c.s[lane];
  } else {
#pragma unroll
    for (int p = 0; p < 4; p++) c.s[p] = D.scale_ln[4 * (size_t)gl_idx + p];
  }
  double M[16];
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int q = 0; q < 4; q++) M[4 * p + q] = c.s[p] * c.s[q] * E[4 * p + q];
#pragma unroll
  for (int p = 0; p < 4; p++) M[5 * p] += clampd(M[5 * p], P.min_lm_diag, P.max_lm_diag) / V.radius;
  // 4x4 Cholesky M = L L^T, then Minv = L^-T L^-1
  double Lc[16];
  c.ok = true;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double dj = M[5 * j];
#pragma unroll
    for (int k = 0; k < j; k++) dj -= Lc[4 * j + k] * Lc[4 * j + k];
    if (!(dj > 0.0)) c.ok = false;
    dj = sqrt(dj);
    Lc[5 * j] = dj;
#pragma unroll
    for (int i = j + 1; i < 4; i++) {
      double s = M[4 * i + j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= Lc[4 * i + k] * Lc[4 * j + k];
      Lc[4 * i + j] = s / dj;
    }
  }
  double Li[16];   // L^-1 (lower)
#pragma unroll
  for (int col = 0; col < 4; col++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < col) { Li[4 * i + col] = 0.0; continue; }
      double s = (i == col) ? 1.0 : 0.0;
#pragma unroll
      for (int k = col; k < i; k++) s -= Lc[4 * i + k] * Li[4 * k + col];
      Li[4 * i + col] = s / Lc[5 * i];
    }
  }
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      double s = 0.0;
#pragma unroll
      for (int k = (p > q ? p : q); k < 4; k++) s += Li[4 * k + p] * Li[4 * k + q];
      c.Minv[4 * p + q] = s;
    }
  return c;
}

__device__ void build_line(const Dev &D, const Params &P, const WinView &V, int gl, int f0, int n, int lane) {
  LineLane L;
  const bool valid = lane < n;
  load_line_lane(D, V, f0 + lane, valid, L);
  const LineCore lc = line_core(D, P, V, gl, L, lane);
  if (!lc.ok) { if (lane == 0) atomicAdd(D.acc + (size_t)V.w * ACC_STRIDE + ACC_FAIL, 1.0); return; }
  if (lane == 0) {
    double m = fmax(fmax(fabs(lc.gl[0]), fabs(lc.gl[1])), fmax(fabs(lc.gl[2]), fabs(lc.gl[3])));
    atomic_max_nonneg(D.acc + (size_t)V.w * ACC_STRIDE + ACC_GMAX + D.rank, m);
  }
  const int d = V.d;
  double gt[4];
#pragma unroll
  for (int c = 0; c < 4; c++) gt[c] = lc.s[c] * lc.gl[c];
  // W = Jp^T Jl D_s (6x4), T = W Minv
  double W[6][4], T[6][4];
#pragma unroll
  for (int p = 0; p < 6; p++) {
#pragma unroll
    for (int c = 0; c < 4; c++)
      W[p][c] = lc.s[c] * (L.Jp[0][p] * L.Jl[0][c] + L.Jp[1][p] * L.Jl[1][c] + L.Jp[2][p] * L.Jl[2][c]);
#pragma unroll
    for (int c = 0; c < 4; c++)
      T[p][c] = W[p][0] * lc.Minv[c] + W[p][1] * lc.Minv[4 + c] + W[p][2] * lc.Minv[8 + c] + W[p][3] * lc.Minv[12 + c];
  }
  if (valid) {
#pragma unroll
    for (int p = 0; p < 6; p++) {
      const double g = L.Jp[0][p] * L.r[0] + L.Jp[1][p] * L.r[1] + L.Jp[2][p] * L.r[2];
      atomicAdd(V.gfull + L.cj + p, g);
      atomicAdd(V.gS + L.cj + p, g - (T[p][0] * gt[0] + T[p][1] * gt[1] + T[p][2] * gt[2] + T[p][3] * gt[3]));
      atomicAdd(V.colsq + L.cj + p, L.Jp[0][p] * L.Jp[0][p] + L.Jp[1][p] * L.Jp[1][p] + L.Jp[2][p] * L.Jp[2][p]);
#pragma unroll
      for (int q = p; q < 6; q++) {
        const double h = L.Jp[0][p] * L.Jp[0][q] + L.Jp[1][p] * L.Jp[1][q] + L.Jp[2][p] * L.Jp[2][q] -
                         (T[p][0] * W[q][0] + T[p][1] * W[q][1] + T[p][2] * W[q][2] + T[p][3] * W[q][3]);
        atomicAdd(V.S + (size_t)(L.cj + p) * d + L.cj + q, h);
      }
    }
  }
  for (int step = 1; step < n; step++) {
    double W2[6][4];
#pragma unroll
    for (int p = 0; p < 6; p++)
#pragma unroll
      for (int c = 0; c < 4; c++) W2[p][c] = __shfl_down_sync(FULL, W[p][c], step);
    const int cj2 = __shfl_down_sync(FULL, L.cj, step);
    if (lane + step < n) {
#pragma unroll
      for (int p = 0; p < 6; p++)
#pragma unroll
        for (int q = 0; q < 6; q++)
          addS(V.S, d, L.cj + p, cj2 + q, -(T[p][0] * W2[q][0] + T[p][1] * W2[q][1] + T[p][2] * W2[q][2] + T[p][3] * W2[q][3]));
    }
  }
}

__device__ void backsub_line(const Dev &D, const Params &P, const WinView &V, int gl, int f0, int n, int lane) {
  LineLane L;
  const bool valid = lane < n;
  load_line_lane(D, V, f0 + lane, valid, L);
  const LineCore lc = line_core(D, P, V, gl, L, lane);
  double u[3] = {0.0, 0.0, 0.0};
  if (valid) {
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const double dc = V.delta[L.cj + c];
      u[0] += L.Jp[0][c] * dc; u[1] += L.Jp[1][c] * dc; u[2] += L.Jp[2][c] * dc;
    }
  }
  double t[4], dk[4];
#pragma unroll
  for (int c = 0; c < 4; c++) t[c] = lc.s[c] * (lc.gl[c] + wsum(L.Jl[0][c] * u[0] + L.Jl[1][c] * u[1] + L.Jl[2][c] * u[2]));
#pragma unroll
  for (int c = 0; c < 4; c++)
    dk[c] = -lc.s[c] * (lc.Minv[4 * c] * t[0] + lc.Minv[4 * c + 1] * t[1] + lc.Minv[4 * c + 2] * t[2] + lc.Minv[4 * c + 3] * t[3]);
  double mc = 0.0;
#pragma unroll
  for (int row = 0; row < 3; row++) {
    const double jd = u[row] + L.Jl[row][0] * dk[0] + L.Jl[row][1] * dk[1] + L.Jl[row][2] * dk[2] + L.Jl[row][3] * dk[3];
    mc += jd * (L.r[row] + 0.5 * jd);
  }
  mc = wsum(mc);
  if (lane == 0) {
    const int w = V.w;
    const int cur = D.cur[w];
    double s2 = 0.0, x2 = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const double x = D.ortho[cur][4 * (size_t)gl + c];
      D.delta_ln[4 * (size_t)gl + c] = dk[c];
      D.ortho[cur ^ 1][4 * (size_t)gl + c] = x + dk[c];
      s2 += dk[c] * dk[c]; x2 += x * x;
    }
    double *acc = D.acc + (size_t)w * ACC_STRIDE;
    atomicAdd(acc + ACC_MODEL, -mc);
    atomicAdd(acc + ACC_STEP2, s2);
    atomicAdd(acc + ACC_XNORM2, x2);
  }
}

template <bool kBack>
__global__ void __launch_bounds__(128) k_lines(Dev D, Params P) {
  const int gl = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gl >= D.nL) return;
  if (D.nranks > 1 && (gl % D.nranks) != D.rank) return;
  const int w = D.ln_win[gl];
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  const WinView V = window_view(D, w);
  const int f0 = D.ln_begin[gl], n = D.ln_end[gl] - f0;
  if (kBack) {
    if (D.acc[(size_t)w * ACC_STRIDE + ACC_FAIL] != 0.0) return;
    if (n <= 0) {
      if (lane < 4) { const int cur = D.cur[w]; D.delta_ln[4 * (size_t)gl + lane] = 0.0; D.ortho[cur ^ 1][4 * (size_t)gl + lane] = D.ortho[cur][4 * (size_t)gl + lane]; }
      return;
    }
    backsub_line(D, P, V, gl, f0, n, lane);
  } else {
    if (n <= 0) return;
    build_line(D, P, V, gl, f0, n, lane);
  }
}

// ------------------------------------------------------------------------------------------------
// IMU factors: 30x30 dense block at camera offset 15*frame_i.  One warp per factor.
__global__ void __launch_bounds__(128) k_build_imu(Dev D) {
  const int f = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= D.nImu) return;
  if (D.nranks > 1 && D.rank != 0) return;
  const int2 ix = D.imu_idx[f];
  const int w = ix.y;
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  const WinView V = window_view(D, w);
  const int c0 = 15 * (ix.x - D.frame_off[w]);
  const double *R = D.rec_imu + (size_t)f * REC_IMU;
  const double *J = R + 15;
  // 465 upper-triangle entries of J^T J, spread over the lanes
  for (int e = lane; e < 465; e += 32) {
    // unrank e -> (p <= q) in a 30x30 upper triangle, row-major
    int p = 0, rem = e;
    while (rem >= 30 - p) { rem -= 30 - p; p++; }
    const int q = p + rem;
    double h = 0.0;
#pragma unroll
    for (int i = 0; i < 15; i++) h += J[i * 30 + p] * J[i * 30 + q];
    atomicAdd(V.S + (size_t)(c0 + p) * V.d + c0 + q, h);
  }
  if (lane < 30) {
    double g = 0.0, q2 = 0.0;
#pragma unroll
    for (int i = 0; i < 15; i++) { const double j = J[i * 30 + lane]; g += j * R[i]; q2 += j * j; }
    atomicAdd(V.gfull + c0 + lane, g);
    atomicAdd(V.gS + c0 + lane, g);
    atomicAdd(V.colsq + c0 + lane, q2);
  }
}

// Prior: H += J0^T J0 (precomputed), g += J0^T r.  One CTA per window.
__global__ void __launch_bounds__(256) k_build_prior(Dev D) {
  const int w = blockIdx.x;
  if (D.nranks > 1 && D.rank != 0) return;
  const int n = D.prior_off[w + 1] - D.prior_off[w];
  if (n <= 0 || !(D.ctl[w].state & WS_ACTIVE)) return;
  extern __shared__ int cmap[];   // J0 column -> camera offset (-1: constant block)
  const WinView V = window_view(D, w);
  for (int c = threadIdx.x; c < n; c += blockDim.x) cmap[c] = -1;
  __syncthreads();
  for (int b = D.pblk_off[w] + threadIdx.x; b < D.pblk_off[w + 1]; b += blockDim.x) {
    const int kind = D.pblk_kind[b], cam = D.pblk_cam[b], col = D.pblk_col[b];
    const int ls = (kind == 0 || kind == 2) ? 6 : (kind == 1 ? 9 : 1);
    if (cam >= 0) for (int c = 0; c < ls; c++) cmap[col + c] = cam + c;
  }
  __syncthreads();
  const double *H = D.prior_H + D.priorJ_off[w];
  const double *J0 = D.prior_J + D.priorJ_off[w];
  const double *r = D.rec_prior + D.prior_off[w];
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int p = e / n, q = e - p * n;
    const int cp = cmap[p], cq = cmap[q];
    if (cp < 0 || cq < 0 || cp > cq) continue;
    atomicAdd(V.S + (size_t)cp * V.d + cq, H[e]);
  }
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const int cp = cmap[p];
    if (cp < 0) continue;
    double g = 0.0;
    for (int i = 0; i < n; i++) g += J0[(size_t)i * n + p] * r[i];
    atomicAdd(V.gfull + cp, g);
    atomicAdd(V.gS + cp, g);
    atomicAdd(V.colsq + cp, H[(size_t)p * n + p]);
  }
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int launch_build(const Dev &D, const Params &P, int max_prior_n, cudaStream_t st) {
  int n = 0;
  if (D.nP) { k_points<false><<<cdiv(D.nP, 4), 128, 0, st>>>(D, P); n++; }
  if (D.nL) { k_lines<false><<<cdiv(D.nL, 4), 128, 0, st>>>(D, P); n++; }
  if (D.nImu) { k_build_imu<<<cdiv(D.nImu, 4), 128, 0, st>>>(D); n++; }
  if (D.nPriorR) { k_build_prior<<<D.B, 256, (size_t)max_prior_n * sizeof(int), st>>>(D); n++; }
  return n;
}

// camera-only factors (IMU, prior) of the normal equations
int launch_build_cam(const Dev &D, int max_prior_n, cudaStream_t st) {
  int n = 0;
  if (D.nImu) { k_build_imu<<<cdiv(D.nImu, 4), 128, 0, st>>>(D); n++; }
  if (D.nPriorR) { k_build_prior<<<D.B, 256, (size_t)max_prior_n * sizeof(int), st>>>(D); n++; }
  return n;
}

int launch_backsub(const Dev &D, const Params &P, cudaStream_t st) {
  int n = 0;
  if (D.nP) { k_points<true><<<cdiv(D.nP, 4), 128, 0, st>>>(D, P); n++; }
  if (D.nL) { k_lines<true><<<cdiv(D.nL, 4), 128, 0, st>>>(D, P); n++; }
  return n;
}

}  // namespace uvs
