// uvs_build3.cu — normal equations + Schur complement + back-substitution for the reference's window
// size (<= 12 six-wide camera blocks: 11 poses + extrinsic, no td): the production path.
//
// Same algebra as uvs_build.cu (Ceres SPARSE_SCHUR restated), organised so that every record is read a fixed number of
// times (elimination + direct terms) and the dense contractions run on the FP64 tensor cores:
//
//   k_core_points / k_core_lines   one lane GROUP (4 / 8 lanes) per landmark: the warp stages the contiguous record span of
//        its landmarks in shared memory (cp.async), the lanes split the observations, merge by shuffles, eliminate the
//        landmark block and write the "stash": Y = W (E + D^2)^-1/2 (6 values per camera block of the landmark; 6x4 per
//        line observation), z = (E + D^2)^-1/2 g, Jacobi scale, LM diagonal.  W (E+D^2)^-1 W^T = Y Y^T,  W (E+D^2)^-1 g = Y z.
//   k_direct_fused (k_direct when an extrinsic is free)   direct terms  sum_f J_a^T J_b, J^T r, column norms per camera-block
//        pair from lists sorted by block pair at upload (k_prep_direct): a warp owns a pair, the records are gathered
//        with cp.async and contracted on the FP64 tensor cores (G^T G with G = [J_a | J_b | r]).
//   k_window_system                one CTA per window: Schur terms as ONE dense rank update  V -= Y Y^T  over all landmark
//        columns, streamed with TMA bulk copies (4 stages), FP64 mma.sync, eight balanced tile units for the 66-row window.
//   k_window_tail                  IMU blocks ([J r]^T [J r] on the tensor cores) and the prior.
//   k_back_points / k_back_lines   one lane group per point / line: delta_k = -(E+D^2)^-1/2 (z + Y^T delta_c).
// All of them add into the window's reduced system with FP64 reductions (they run side by side on auxiliary streams).
// The model cost change uses  -(g^T y + y^T H y / 2) = (y^T D^2 y - g^T y) / 2  (y solves (H + D^2) y = -g),
// which needs no Jacobians; k_chol / k_chol_chain add the camera part.
#include <algorithm>

#include "uvs_device.cuh"
#include "uvs_kernels.h"
#include "uvs_stash.cuh"

namespace uvs {

// The records of a warp's eight points are one contiguous span of rec_proj (factors of a landmark are contiguous,
// landmarks consecutive): the warp copies it 32 records at a time into shared memory with coalesced 16-byte
// asynchronous copies (a lane reading its own 320-byte record with 8-byte loads costs 32 L1 wavefronts per
// instruction), then every lane reads its record as 128-bit words (row stride 42 doubles: conflict-free).
constexpr int PSTR = 42;

__global__ void __launch_bounds__(128) k_core_points(Dev D, Params P, Stash S) {
  __shared__ __align__(16) double stage_all[4][32 * PSTR];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double *stage = stage_all[threadIdx.x >> 5];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int gp = t / LPP, sub = t - gp * LPP;
  const unsigned gmask = ((1u << LPP) - 1u) << (lane & ~(LPP - 1));
  bool act = gp < D.nP;   // uniform over the lane group
  int w = 0, f0 = 0, n = 0;
  if (act) { w = D.pt_win[gp]; act = (D.ctl[w].state & WS_ACTIVE) != 0; }
  const int mp = S.mp;
  double *Y = nullptr, *ph = nullptr;
  if (act) {
    Y = S.Y + (colbase(D, w) + (gp - D.point_off[w])) * mp;
    ph = S.ph + 4 * (size_t)gp;
    f0 = D.pt_begin[gp]; n = D.pt_end[gp] - f0;
    const bool mine = D.nranks <= 1 || (gp % D.nranks) == D.rank;
    // the column was zero-filled at upload and its sparsity pattern never changes: only the blocks are rewritten
    if (n <= 0 || !mine) { if (sub == 0) ph[1] = 0.0; act = false; }
  }
  const int lo = __reduce_min_sync(full, act ? f0 : 0x7fffffff), hi = __reduce_max_sync(full, act ? f0 + n : 0);
  if (lo >= hi) return;   // uniform over the warp
  const bool ex = act && (D.win_flags[w] & WF_EXTRINSIC) != 0;
  const int fo = act ? D.frame_off[w] : 0, F = act ? D.frame_off[w + 1] - fo : 0;
  double colsq = 0.0, gk = 0.0;
  double wa[6] = {0, 0, 0, 0, 0, 0}, we[6] = {0, 0, 0, 0, 0, 0}, u0[6] = {0, 0, 0, 0, 0, 0};
  for (int c0 = lo; c0 < hi; c0 += 32) {
    const int cnt = min(32, hi - c0);
    const double *src = D.rec_proj + (size_t)c0 * REC_PROJ;
    for (int p = lane; p < cnt * (REC_PROJ / 2); p += 32) {
      const int r = p / (REC_PROJ / 2), q = p - (REC_PROJ / 2) * r;
      cp_async16(stage + r * PSTR + 2 * q, src + 2 * p);
    }
    cp_async_wait_all();
    __syncwarp();
    if (act) {
      for (int f = sub; f < n; f += LPP) {
        const int g = f0 + f - c0;
        if (g < 0 || g >= 32) continue;
        const double2 *r2 = reinterpret_cast<const double2 *>(stage + g * PSTR);
        const double2 rr = r2[0], jl = r2[19];
        const double j0 = jl.x, j1 = jl.y;
        colsq += j0 * j0 + j1 * j1;
        gk += j0 * rr.x + j1 * rr.y;
        const bool first = f < LPP;
        double *yj = Y + 6 * (D.proj_idx[f0 + f].y - fo);
        double ji[12], jj[12];
#pragma unroll
        for (int c = 0; c < 6; c++) { const double2 a = r2[1 + c], b2 = r2[7 + c]; ji[2 * c] = a.x; ji[2 * c + 1] = a.y; jj[2 * c] = b2.x; jj[2 * c + 1] = b2.y; }
        double u[6];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          wa[c] += ji[c] * j0 + ji[6 + c] * j1;
          u[c] = jj[c] * j0 + jj[6 + c] * j1;
          if (first) u0[c] = u[c];
        }
        if (!first) {   // later observations of this lane: parked unscaled, rescaled below (a block is 48 bytes, 16-byte aligned)
          double2 *y2 = reinterpret_cast<double2 *>(yj);
          y2[0] = make_double2(u[0], u[1]); y2[1] = make_double2(u[2], u[3]); y2[2] = make_double2(u[4], u[5]);
        }
        if (ex) {
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const double2 a = r2[13 + c], b2 = r2[16 + c];
            we[2 * c] += a.x * j0 + b2.x * j1; we[2 * c + 1] += a.y * j0 + b2.y * j1;
          }
        }
      }
    }
    __syncwarp();
  }
  if (!act) return;   // uniform over the lane group
  colsq = group_sum<LPP>(gmask, colsq);
  gk = group_sum<LPP>(gmask, gk);
#pragma unroll
  for (int c = 0; c < 6; c++) { wa[c] = group_sum<LPP>(gmask, wa[c]); if (ex) we[c] = group_sum<LPP>(gmask, we[c]); }
  double sk;
  if (!D.ctl[w].have_scale) { sk = 1.0 / (1.0 + sqrt(colsq)); if (sub == 0) D.scale_pt[gp] = sk; }
  else sk = D.scale_pt[gp];
  const double Et = sk * sk * colsq;
  const double D2 = clamp4(Et, P.min_lm_diag, P.max_lm_diag) / D.ctl[w].radius;
  const double sh = rsqrt(Et + D2);
  const double ysc = sk * sh;
  for (int f = sub; f < n; f += LPP) {
    double2 *y2 = reinterpret_cast<double2 *>(Y + 6 * (D.proj_idx[f0 + f].y - fo));
    if (f < LPP) {
      y2[0] = make_double2(ysc * u0[0], ysc * u0[1]); y2[1] = make_double2(ysc * u0[2], ysc * u0[3]); y2[2] = make_double2(ysc * u0[4], ysc * u0[5]);
    } else {
#pragma unroll
      for (int c = 0; c < 3; c++) { const double2 v = y2[c]; y2[c] = make_double2(ysc * v.x, ysc * v.y); }
    }
  }
  if (sub != 0) return;
  const int bi = D.proj_idx[f0].x - fo;
  {
    double2 *y2 = reinterpret_cast<double2 *>(Y + 6 * bi);
    y2[0] = make_double2(ysc * wa[0], ysc * wa[1]); y2[1] = make_double2(ysc * wa[2], ysc * wa[3]); y2[2] = make_double2(ysc * wa[4], ysc * wa[5]);
  }
  if (ex) {
#pragma unroll
    for (int c = 0; c < 6; c++) Y[6 * F + c] = ysc * we[c];
  }
  Y[mp - 2] = sk * gk * sh;
  ph[0] = sk; ph[1] = sh; ph[2] = D2;
  atomic_max_nn3(D.acc + (size_t)w * ACC_STRIDE + ACC_GMAX + D.rank, fabs(gk));
}

// One lane GROUP per line: the lanes split the line's observations (E = sum Jl^T Jl, g), merge by shuffles, every lane
// factors the damped 4x4 block, then each lane writes  Y_f = (Jp^T Jl D_s) L^-T  of its own observations (+ VP factors).
// The line records of a warp's four lines are one contiguous span of rec_line, their VP records one span of rec_vp
// (k_prep_vp keeps the VP observations in line-observation order): both are copied into shared memory with coalesced
// asynchronous copies and read from there by BOTH passes; spans that do not fit (never with <= 12 frames and dense
// ranges) are read from global memory as before.
constexpr int LCAP = 44;   // staged records per warp (4 lines x 11 frames; 46 KB of static shared memory per CTA)

__global__ void __launch_bounds__(128) k_core_lines(Dev D, Params P, Stash S) {
  __shared__ __align__(16) double lstage_all[4][LCAP * REC_LINE];
  __shared__ __align__(16) double vstage_all[4][LCAP * REC_VP];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double *lstage = lstage_all[threadIdx.x >> 5], *vstage = vstage_all[threadIdx.x >> 5];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int gl = t / LPL, sub = t - gl * LPL;
  const unsigned gmask = ((1u << LPL) - 1u) << (lane & ~(LPL - 1));
  bool act = gl < D.nL;   // uniform over the lane group
  int w = 0, f0 = 0, n = 0;
  if (act) { w = D.ln_win[gl]; act = (D.ctl[w].state & WS_ACTIVE) != 0; }
  const int mp = S.mp;
  double *Y = nullptr, *hd = nullptr;
  if (act) {
    f0 = D.ln_begin[gl]; n = D.ln_end[gl] - f0;
    Y = S.Y + (colbase(D, w) + (D.point_off[w + 1] - D.point_off[w]) + 4LL * (gl - D.line_off[w])) * mp;   // 4 columns
    hd = S.lh + 24 * (size_t)gl;
    const bool mine = D.nranks <= 1 || (gl % D.nranks) == D.rank;
    // Linv[0][0] = 0 marks "no step" for the back-substitution
    if (n <= 0 || !mine) { if (sub == 0) hd[8] = 0.0; act = false; }
  }
  int myvlo = 0x7fffffff, myvhi = -1;
  if (act) for (int f = sub; f < n; f += LPL) { const int vi = D.line_idx4[f0 + f].w; if (vi >= 0) { myvlo = min(myvlo, vi); myvhi = max(myvhi, vi); } }
  const int lo = __reduce_min_sync(full, act ? f0 : 0x7fffffff), hi = __reduce_max_sync(full, act ? f0 + n : 0);
  if (lo >= hi) return;   // uniform over the warp
  const int vlo = __reduce_min_sync(full, myvlo), vhi = __reduce_max_sync(full, myvhi) + 1;
  const bool staged_l = hi - lo <= LCAP, staged_v = vhi > vlo && vhi - vlo <= LCAP;
  if (staged_l) {
    const double *src = D.rec_line + (size_t)lo * REC_LINE;
    for (int p = lane; p < (hi - lo) * (REC_LINE / 2); p += 32) cp_async16(lstage + 2 * p, src + 2 * p);
  }
  if (staged_v) {
    const double *src = D.rec_vp + (size_t)vlo * REC_VP;
    for (int p = lane; p < (vhi - vlo) * REC_VP; p += 32) cp_async8(vstage + p, src + p);
  }
  cp_async_wait_all();
  __syncwarp();
  if (!act) return;   // uniform over the lane group
  auto line_rec = [&](int f) { return staged_l ? lstage + (f0 + f - lo) * REC_LINE : D.rec_line + (size_t)(f0 + f) * REC_LINE; };
  auto vp_rec = [&](int vi) { return staged_v ? vstage + (vi - vlo) * REC_VP : D.rec_vp + (size_t)vi * REC_VP; };
  const int fo = D.frame_off[w];
  double E[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, g[4] = {0, 0, 0, 0};
  for (int f = sub; f < n; f += LPL) {
    const double2 *r2 = reinterpret_cast<const double2 *>(line_rec(f));
    const double2 rr = r2[0];
#pragma unroll
    for (int row = 0; row < 2; row++) {
      const double2 a01 = r2[7 + 2 * row], a23 = r2[8 + 2 * row];
      const double a0 = a01.x, a1 = a01.y, a2 = a23.x, a3 = a23.y, rw = row ? rr.y : rr.x;
      E[0] += a0 * a0; E[1] += a0 * a1; E[2] += a0 * a2; E[3] += a0 * a3; E[4] += a1 * a1; E[5] += a1 * a2; E[6] += a1 * a3;
      E[7] += a2 * a2; E[8] += a2 * a3; E[9] += a3 * a3;
      g[0] += a0 * rw; g[1] += a1 * rw; g[2] += a2 * rw; g[3] += a3 * rw;
    }
    const int vi = D.line_idx4[f0 + f].w;
    if (vi >= 0) {
      const double *q = vp_rec(vi);
      const double a0 = q[7], a1 = q[8], a2 = q[9], a3 = q[10], rw = q[0];
      E[0] += a0 * a0; E[1] += a0 * a1; E[2] += a0 * a2; E[3] += a0 * a3; E[4] += a1 * a1; E[5] += a1 * a2; E[6] += a1 * a3;
      E[7] += a2 * a2; E[8] += a2 * a3; E[9] += a3 * a3;
      g[0] += a0 * rw; g[1] += a1 * rw; g[2] += a2 * rw; g[3] += a3 * rw;
    }
  }
#pragma unroll
  for (int k = 0; k < 10; k++) E[k] = group_sum<LPL>(gmask, E[k]);
#pragma unroll
  for (int k = 0; k < 4; k++) g[k] = group_sum<LPL>(gmask, g[k]);
  double s[4];
  const double Ed[4] = {E[0], E[4], E[7], E[9]};
  if (!D.ctl[w].have_scale) {
#pragma unroll
    for (int c = 0; c < 4; c++) { s[c] = 1.0 / (1.0 + sqrt(Ed[c])); if (sub == 0) D.scale_ln[4 * (size_t)gl + c] = s[c]; }
  } else {
#pragma unroll
    for (int c = 0; c < 4; c++) s[c] = D.scale_ln[4 * (size_t)gl + c];
  }
  // M = D_s E D_s + D^2 (lower), Cholesky, inverse of the factor
  const double radius = D.ctl[w].radius;
  double M[4][4], D2[4];
  M[0][0] = s[0] * s[0] * E[0]; M[1][0] = s[1] * s[0] * E[1]; M[2][0] = s[2] * s[0] * E[2]; M[3][0] = s[3] * s[0] * E[3];
  M[1][1] = s[1] * s[1] * E[4]; M[2][1] = s[2] * s[1] * E[5]; M[3][1] = s[3] * s[1] * E[6];
  M[2][2] = s[2] * s[2] * E[7]; M[3][2] = s[3] * s[2] * E[8]; M[3][3] = s[3] * s[3] * E[9];
#pragma unroll
  for (int c = 0; c < 4; c++) { D2[c] = clamp4(M[c][c], P.min_lm_diag, P.max_lm_diag) / radius; M[c][c] += D2[c]; }
  double L[4][4], Li[4][4];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double dj = M[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) dj -= L[j][k] * L[j][k];
    if (!(dj > 0.0)) ok = false;
    const double id = rsqrt(dj);
    L[j][j] = dj * id;
#pragma unroll
    for (int i = j + 1; i < 4; i++) {
      double tt = M[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) tt -= L[i][k] * L[j][k];
      L[i][j] = tt * id;
    }
  }
  if (!ok) {
    if (sub == 0) { hd[8] = 0.0; atomicAdd(D.acc + (size_t)w * ACC_STRIDE + ACC_FAIL, 1.0); }
    return;
  }
#pragma unroll
  for (int col = 0; col < 4; col++)
#pragma unroll
    for (int i = col; i < 4; i++) {
      double tt = (i == col) ? 1.0 : 0.0;
#pragma unroll
      for (int k = col; k < i; k++) tt -= L[i][k] * Li[k][col];
      Li[i][col] = tt / L[i][i];
    }
  if (sub == 0) {
    // z = L^-1 (D_s g)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      double tt = 0.0;
#pragma unroll
      for (int k = 0; k <= c; k++) tt += Li[c][k] * s[k] * g[k];
      Y[c * mp + mp - 2] = tt; hd[c] = s[c]; hd[4 + c] = D2[c];
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int k = 0; k < 4; k++) hd[8 + 4 * i + k] = k <= i ? Li[i][k] : 0.0;
    atomic_max_nn3(D.acc + (size_t)w * ACC_STRIDE + ACC_GMAX + D.rank, fmax(fmax(fabs(g[0]), fabs(g[1])), fmax(fabs(g[2]), fabs(g[3]))));
  }
  // Y blocks of this lane's observations
  for (int f = sub; f < n; f += LPL) {
    const double2 *r2 = reinterpret_cast<const double2 *>(line_rec(f));
    const int4 ix = D.line_idx4[f0 + f];
    const double *q = ix.w >= 0 ? vp_rec(ix.w) : nullptr;
    double *Yf = Y + 6 * (ix.x - fo);
    double jl[3][4], jp[2][6];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const double2 a = r2[7 + c], b2 = r2[9 + c];
      jl[0][2 * c] = a.x * s[2 * c]; jl[0][2 * c + 1] = a.y * s[2 * c + 1];
      jl[1][2 * c] = b2.x * s[2 * c]; jl[1][2 * c + 1] = b2.y * s[2 * c + 1];
    }
#pragma unroll
    for (int c = 0; c < 4; c++) jl[2][c] = q ? q[7 + c] * s[c] : 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double2 a = r2[1 + c], b2 = r2[4 + c];
      jp[0][2 * c] = a.x; jp[0][2 * c + 1] = a.y; jp[1][2 * c] = b2.x; jp[1][2 * c + 1] = b2.y;
    }
#pragma unroll
    for (int p = 0; p < 6; p += 2) {   // two rows at a time: the 48-byte block of a column is written as three 128-bit words
      double tt[2][4];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const double a0 = jp[0][p + h], a1 = jp[1][p + h], a2 = q ? q[1 + p + h] : 0.0;
        double W4[4];
#pragma unroll
        for (int c = 0; c < 4; c++) W4[c] = a0 * jl[0][c] + a1 * jl[1][c] + a2 * jl[2][c];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          double t1 = 0.0;
#pragma unroll
          for (int k = 0; k <= c; k++) t1 += W4[k] * Li[c][k];
          tt[h][c] = t1;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) *reinterpret_cast<double2 *>(Yf + c * mp + p) = make_double2(tt[0][c], tt[1][c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// back-substitution
// The columns of a warp's eight points are contiguous when they belong to one window: the warp copies them into shared
// memory with coalesced 16-byte asynchronous copies and the lane groups (4 lanes per point) read their blocks from
// there; warps that straddle windows read global memory.
constexpr int MP_MAX = 84;   // largest column stride (12 camera blocks + z, padded to 4 mod 16)

__global__ void __launch_bounds__(128) k_back_points(Dev D, Stash S) {
  __shared__ __align__(16) double stage_all[4][(32 / LPP) * MP_MAX];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double *stage = stage_all[threadIdx.x >> 5];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int gp = t / LPP, sub = t - gp * LPP;
  const unsigned gmask = ((1u << LPP) - 1u) << (lane & ~(LPP - 1));
  bool valid = gp < D.nP && (D.nranks <= 1 || (gp % D.nranks) == D.rank);
  int w = 0;
  if (valid) {
    w = D.pt_win[gp];
    valid = (D.ctl[w].state & WS_ACTIVE) && D.acc[(size_t)w * ACC_STRIDE + ACC_FAIL] == 0.0;
  }
  const int mp = S.mp;
  const double *Y = valid ? S.Y + (colbase(D, w) + (gp - D.point_off[w])) * mp : nullptr;
  // staged path: every group of the warp valid and in the window of lane 0 (columns contiguous from lane 0's on)
  const int w0 = __shfl_sync(full, w, 0);
  const bool staged = __all_sync(full, valid && w == w0) && mp <= MP_MAX;
  if (staged) {
    const double *src = reinterpret_cast<const double *>(__shfl_sync(full, (unsigned long long)Y, 0));
    for (int p = lane; p < (32 / LPP) * mp / 2; p += 32) cp_async16(stage + 2 * p, src + 2 * p);
    cp_async_wait_all();
    __syncwarp();
  }
  double mc = 0.0, s2 = 0.0, x2 = 0.0;
  if (valid) {   // uniform over the lane group
    const int cur = D.cur[w];
    const double *y = staged ? stage + (lane / LPP) * mp : Y;
    const double *ph = S.ph + 4 * (size_t)gp;
    const double sk = ph[0], sh = ph[1], D2 = ph[2];
    double dk = 0.0;
    if (sh != 0.0) {
      const int F = D.frame_off[w + 1] - D.frame_off[w];
      const int nb = F + ((D.win_flags[w] & WF_EXTRINSIC) ? 1 : 0);
      const double *dl = D.delta_cam + D.cam_off[w];
      double u = 0.0;
      for (int b = sub; b < nb; b += LPP) {   // the lanes split the camera blocks; the extrinsic block (b == F) sits at 15 F as well
#pragma unroll
        for (int c = 0; c < 6; c++) u += y[6 * b + c] * dl[15 * b + c];
      }
      u = group_sum<LPP>(gmask, u);
      const double z = y[mp - 2];
      // scaled column: z = sk sh g, entries sk sh W;  unscaled column (fused path): z = g, entries W
      const double yk = S.unscaled_pts ? -sh * sk * sh * (z + u) : -sh * (z + u);
      dk = sk * yk;
      mc = 0.5 * (D2 * yk * yk - (S.unscaled_pts ? sk * z : z / sh) * yk);
    }
    if (sub == 0) {
      const double lam = D.inv_depth[cur][gp];
      D.delta_pt[gp] = dk;
      D.inv_depth[cur ^ 1][gp] = lam + dk;
      s2 = dk * dk; x2 = lam * lam;
    }
  }
  add_win3(D.acc, w, valid && sub == 0, mc, s2, x2);
}

// back-substitution of the lines, one lane group per line (4 columns).  Staging the columns like k_back_points was
// measured slower (43 KB of shared memory per CTA halves the resident warps of a latency-bound kernel: 76 -> 90 us).
__global__ void __launch_bounds__(128) k_back_lines(Dev D, Stash S) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int gl = t / LPL, sub = t - gl * LPL;
  const unsigned gmask = ((1u << LPL) - 1u) << ((threadIdx.x & 31) & ~(LPL - 1));
  bool valid = gl < D.nL && (D.nranks <= 1 || (gl % D.nranks) == D.rank);
  int w = 0;
  if (valid) {
    w = D.ln_win[gl];
    valid = (D.ctl[w].state & WS_ACTIVE) && D.acc[(size_t)w * ACC_STRIDE + ACC_FAIL] == 0.0;
  }
  double mc = 0.0, s2 = 0.0, x2 = 0.0;
  if (valid) {   // uniform over the lane group
    const int mp = S.mp, cur = D.cur[w];
    const double *Y = S.Y + (colbase(D, w) + (D.point_off[w + 1] - D.point_off[w]) + 4LL * (gl - D.line_off[w])) * mp;
    const double *hd = S.lh + 24 * (size_t)gl;
    double dk[4] = {0, 0, 0, 0};
    if (hd[8] != 0.0) {
      const int F = D.frame_off[w + 1] - D.frame_off[w];
      const double *dl = D.delta_cam + D.cam_off[w];
      double tz[4], yk[4];
      // the lanes split the camera blocks of  u = Y^T delta_c: a block of a column is 48 bytes, 16-byte aligned -> three
      // 128-bit loads; the deltas of the lane's blocks are loaded once for the four columns (the kernel sat at 86 % L1 throughput)
      double u4[4] = {0.0, 0.0, 0.0, 0.0};
      for (int b = sub; b < F; b += LPL) {
        double dv[6];
#pragma unroll
        for (int p = 0; p < 6; p++) dv[p] = dl[15 * b + p];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const double2 *y2 = reinterpret_cast<const double2 *>(Y + c * mp + 6 * b);
          const double2 a0 = y2[0], a1 = y2[1], a2 = y2[2];
          u4[c] += a0.x * dv[0] + a0.y * dv[1] + a1.x * dv[2] + a1.y * dv[3] + a2.x * dv[4] + a2.y * dv[5];
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) tz[c] = Y[c * mp + mp - 2] + group_sum<LPL>(gmask, u4[c]);   // z + u
      // y_k = -L^-T (z + u)
#pragma unroll
      for (int c = 0; c < 4; c++) {
        double a = 0.0;
#pragma unroll
        for (int k = c; k < 4; k++) a += hd[8 + 4 * k + c] * tz[k];
        yk[c] = -a;
      }
      // g~^T y_k = (L z)^T y_k = -z^T (z + u)
      double gy = 0.0, dy = 0.0;
#pragma unroll
      for (int c = 0; c < 4; c++) { gy -= Y[c * mp + mp - 2] * tz[c]; dy += hd[4 + c] * yk[c] * yk[c]; dk[c] = hd[c] * yk[c]; }
      mc = 0.5 * (dy - gy);
    }
    if (sub == 0) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double x = D.ortho[cur][4 * (size_t)gl + c];
        D.delta_ln[4 * (size_t)gl + c] = dk[c];
        D.ortho[cur ^ 1][4 * (size_t)gl + c] = x + dk[c];
        s2 += dk[c] * dk[c]; x2 += x * x;
      }
    }
  }
  add_win3(D.acc, w, valid && sub == 0, mc, s2, x2);
}

// ------------------------------------------------------------------------------------------------
// Direct-term lists: per window, items {factor, kind} sorted (stable) by camera-block pair.
//   kinds: 0 proj (i,i)  1 proj (j,j)  2 proj (i,j) i<j  3 proj (j,i) j<i  4 proj (i,ex)  5 proj (j,ex)  6 proj (ex,ex)
//          7 line (j,j)  8 vp (j,j)
constexpr int KMAX = 80;   // >= 12 * 13 / 2 + 1 keys per window

struct DirectLists {
  int2 *items;      // [6 nProj + nLobs + nVobs]
  int *off;         // [B][KMAX]  start of every key's list (exclusive scan; off[key+1] is its end)
};

__device__ __forceinline__ long long direct_base(const Dev &D, int w) {
  return 6LL * D.proj_off[w] + D.lobs_off[w] + D.vobs_off[w];
}

// `fused` (batches without a free extrinsic): a projection factor appears ONCE, under its off-diagonal pair (i, j) -
// k_direct_fused derives the (i,i), (j,j) blocks and the gradient from the same read of the record.
__global__ void __launch_bounds__(128) k_prep_direct(Dev D, DirectLists L, int fused) {
  __shared__ int cnt_all[4][KMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * 4 + warp;
  if (w >= D.B) return;
  int *cnt = cnt_all[warp];
  const int fo = D.frame_off[w], F = D.frame_off[w + 1] - fo;
  const bool ex = (D.win_flags[w] & WF_EXTRINSIC) != 0;
  const int j0 = D.proj_off[w], j1 = D.proj_off[w + 1], a0 = D.lobs_off[w], a1 = D.lobs_off[w + 1], v0 = D.vobs_off[w], v1 = D.vobs_off[w + 1];
  int2 *items = L.items + direct_base(D, w);
  int *off = L.off + (size_t)w * KMAX;
  const unsigned full = 0xffffffffu;
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 0) { for (int k = lane; k < KMAX; k += 32) cnt[k] = 0; }
    else {
      // exclusive scan of the counts -> cursors
      if (lane == 0) { int run = 0; for (int k = 0; k < KMAX; k++) { const int c = cnt[k]; cnt[k] = run; off[k] = run; run += c; } }
    }
    __syncwarp();
    // emit(key, factor, kind): pass 0 counts, pass 1 scatters stably in (slot, factor) order
    auto emit = [&](bool active, int key, int fidx, int kind) {
      if (pass == 0) { if (active) atomicAdd(&cnt[key], 1); return; }
      const unsigned act = __ballot_sync(full, active);
      if (active) {
        const unsigned same = __match_any_sync(act, key);
        const int rank = __popc(same & ((1u << lane) - 1u));
        const int pos = cnt[key] + rank;
        items[pos] = make_int2(fidx, kind);
      }
      __syncwarp();
      if (active) {
        const unsigned same = __match_any_sync(act, key);
        if ((same & ((1u << lane) - 1u)) == 0) cnt[key] += __popc(same);
      }
      __syncwarp();
    };
    for (int slot = fused ? 2 : 0; slot < (fused ? 3 : (ex ? 6 : 3)); slot++) {
      for (int base = j0; base < j1; base += 32) {
        const int f = base + lane;
        const bool act = f < j1;
        int key = 0, kind = 0;
        if (act) {
          const int4 ix = D.proj_idx[f];
          const int i = ix.x - fo, j = ix.y - fo;
          if (slot == 0) { key = pair_key(i, i); kind = 0; }
          else if (slot == 1) { key = pair_key(j, j); kind = 1; }
          else if (slot == 2) { key = pair_key(i, j); kind = i < j ? 2 : 3; }
          else if (slot == 3) { key = pair_key(i, F); kind = 4; }
          else if (slot == 4) { key = pair_key(j, F); kind = 5; }
          else { key = pair_key(F, F); kind = 6; }
        }
        emit(act, key, f, kind);
      }
    }
    for (int base = a0; base < a1; base += 32) {
      const int f = base + lane;
      const bool act = f < a1;
      int key = 0;
      if (act) { const int j = D.line_idx4[f].x - fo; key = pair_key(j, j); }
      emit(act, key, f, 7);
    }
    for (int base = v0; base < v1; base += 32) {
      const int f = base + lane;
      const bool act = f < v1;
      int key = 0;
      if (act) { const int j = D.vp_idx4[f].x - fo; key = pair_key(j, j); }
      emit(act, key, f, 8);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int WT = 256;       // threads of k_window_system
constexpr int CH = 32;        // landmark columns per chunk (one TMA bulk copy)
constexpr int NSTAGE = 4;     // chunks in flight (TMA bulk copies)
constexpr int YSLACK = 64;    // fragment rows of the last 8-row tiles run past mp into the next column / this slack

// upper-triangle unranking of a 6x6 symmetric block: e in [0,21) -> (p <= q)
__constant__ unsigned char c_sym_p[21] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5};
__constant__ unsigned char c_sym_q[21] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5};

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (sm_90+ PTX)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Direct terms sum_f J_a^T J_b per camera-block pair, gradient J^T r and squared column norms: one WARP per
// (window, block pair[, segment of a diagonal pair's list]) so that thousands of warps hide the gather latency;
// lanes own the entries of the 6x6 block.  Results go straight into the window's (zeroed) system.
constexpr int SEGS = 8;   // diagonal pairs see ~40x more factors than off-diagonal ones: split their lists

__global__ void __launch_bounds__(128) k_direct(Dev D, DirectLists L, int nb_max) {
  const int lane = threadIdx.x & 31;
  const int U = nb_max * SEGS + nb_max * (nb_max - 1) / 2;
  const long long unit = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int w = (int)(unit / U), local = (int)(unit - (long long)w * U);
  if (w >= D.B) return;
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  const int fo = D.frame_off[w], F = D.frame_off[w + 1] - fo;
  const bool ex = (D.win_flags[w] & WF_EXTRINSIC) != 0;
  const int nb = F + (ex ? 1 : 0);
  int a, b, seg = 0;
  if (local < nb_max * SEGS) { a = b = local / SEGS; seg = local - a * SEGS; }
  else {
    const int o = local - nb_max * SEGS;          // strictly upper pair: o = b (b - 1) / 2 + a, a < b
    int i = (int)((sqrtf(8.0f * (float)o + 1.0f) + 1.0f) * 0.5f);
    while (i * (i - 1) / 2 > o) i--;
    while ((i + 1) * i / 2 <= o) i++;
    b = i; a = o - i * (i - 1) / 2;
  }
  if (b >= nb) return;
  const int key = pair_key(a, b);
  const int2 *items = L.items + direct_base(D, w);
  const int *off = L.off + (size_t)w * KMAX;
  int i0 = off[key], i1 = off[key + 1];
  if (a == b) {
    const int len = i1 - i0, per = (len + SEGS - 1) / SEGS;
    i0 += seg * per; i1 = min(i1, i0 + per);
  }
  if (i0 >= i1) return;
  const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
  double *Sg = D.Smat + D.S_off[w];
  const int ra = a < F ? 15 * a : 15 * F, rb = b < F ? 15 * b : 15 * F;
  if (a == b) {
    // 21 symmetric entries + 6 gradient entries
    if (lane >= 27) return;
    int p, q = 0;
    const bool isg = lane >= 21;
    if (!isg) { p = c_sym_p[lane]; q = c_sym_q[lane]; } else p = lane - 21;
    double acc = 0.0;
    for (int it = i0; it < i1; it++) {
      const int2 item = items[it];
      if (D.nranks > 1) {   // factor-parallel mode: only the records of this rank's landmarks are current
        const int lm = item.y <= 6 ? D.proj_idx[item.x].z : (item.y == 7 ? D.line_idx4[item.x].y : D.vp_idx4[item.x].y);
        if ((lm % D.nranks) != D.rank) continue;
      }
      const double *rec; int baseA, nrows;
      if (item.y <= 6) { rec = D.rec_proj + (size_t)item.x * REC_PROJ; nrows = 2; baseA = item.y == 0 ? 2 : (item.y == 1 ? 14 : 26); }
      else if (item.y == 7) { rec = D.rec_line + (size_t)item.x * REC_LINE; nrows = 2; baseA = 2; }
      else { rec = D.rec_vp + (size_t)item.x * REC_VP; nrows = 1; baseA = 1; }
      double t = rec[baseA + p] * (isg ? rec[0] : rec[baseA + q]);
      if (nrows == 2) t += rec[baseA + 6 + p] * (isg ? rec[1] : rec[baseA + 6 + q]);
      acc += t;
    }
    if (!isg) {
      atomicAdd(Sg + (size_t)(ra + p) * d + ra + q, acc);
      if (p == q) atomicAdd(D.colsq_cam + co + ra + p, acc);
    } else {
      atomicAdd(D.gS + co + ra + p, acc);
      atomicAdd(D.gfull + co + ra + p, acc);
    }
  } else {
    // off-diagonal pair (a < b): 36 entries, lanes 0..31 + a second entry on lanes 0..3; one warp owns the block
    const int p0 = lane / 6, q0 = lane - 6 * p0, q1 = 2 + lane;
    double acc0 = 0.0, acc1 = 0.0;
    for (int it = i0; it < i1; it++) {
      const int2 item = items[it];
      if (D.nranks > 1 && (D.proj_idx[item.x].z % D.nranks) != D.rank) continue;
      const double *rec = D.rec_proj + (size_t)item.x * REC_PROJ;
      int baseA, baseB;
      if (item.y == 2) { baseA = 2; baseB = 14; } else if (item.y == 3) { baseA = 14; baseB = 2; }
      else if (item.y == 4) { baseA = 2; baseB = 26; } else { baseA = 14; baseB = 26; }
      acc0 += rec[baseA + p0] * rec[baseB + q0] + rec[baseA + 6 + p0] * rec[baseB + 6 + q0];
      if (lane < 4) acc1 += rec[baseA + 5] * rec[baseB + q1] + rec[baseA + 11] * rec[baseB + 6 + q1];
    }
    atomicAdd(Sg + (size_t)(ra + p0) * d + rb + q0, acc0);
    if (lane < 4) atomicAdd(Sg + (size_t)(ra + 5) * d + rb + q1, acc1);
  }
}


// Fused variant for batches without a free extrinsic (the reference's EuRoC configuration): every factor record is
// read exactly once.  Units per window: one warp per off-diagonal camera-block pair (a < b) takes the projection
// factors of that pair and accumulates the (a,b) block (36 entries, owned by this warp), its share of the (a,a) and
// (b,b) blocks (21 + 21) and of the gradient (6 + 6) - 90 outputs, three per lane; SEGS_D warps per diagonal block
// take the line / VP factors.  The lists are read 32 items at a time (coalesced) and handed round by shuffles, so the
// record loads of consecutive items are independent and stay in flight together.
constexpr int SEGS_D = 4;
constexpr int DSTR = 26;   // doubles staged per item: [r | Ji | Jj] of a projection record; 14 of a line record, 7 of a VP record

__global__ void __launch_bounds__(128) k_direct_fused(Dev D, DirectLists L, int nb_max) {
  __shared__ __align__(16) double stage[4][32 * DSTR];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double *buf = stage[threadIdx.x >> 5];
  const int npair = nb_max * (nb_max - 1) / 2;
  const int U = npair + nb_max * SEGS_D;
  const long long unit = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int w = (int)(unit / U), local = (int)(unit - (long long)w * U);
  if (w >= D.B) return;
  // which unit (pure arithmetic), then ALL the metadata loads at once, then the tests: the warp's life is a chain of
  // dependent memory round trips (metadata -> list -> records), every one that can be merged is ~1 us saved
  int a, b, seg = 0;
  if (local < npair) {
    b = (int)((sqrtf(8.0f * (float)local + 1.0f) + 1.0f) * 0.5f);   // local = b (b - 1) / 2 + a, a < b
    while (b * (b - 1) / 2 > local) b--;
    while ((b + 1) * b / 2 <= local) b++;
    a = local - b * (b - 1) / 2;
  } else {
    a = b = (local - npair) / SEGS_D;
    seg = (local - npair) - a * SEGS_D;
  }
  const int key = pair_key(a, b);          // < KMAX for every unit of the grid
  const int *off = L.off + (size_t)w * KMAX;
  const int wstate = D.ctl[w].state;
  const int fo = D.frame_off[w], nb = D.frame_off[w + 1] - fo;
  const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
  const long long s_off = D.S_off[w];
  const long long ibase = direct_base(D, w);
  const int o0 = off[key], o1 = off[key + 1];
  if (!(wstate & WS_ACTIVE)) return;
  double *Sg = D.Smat + s_off;
  const int2 *items = L.items + ibase;
  if (local < npair) {
    if (b >= nb) return;
    const int i0 = o0, i1 = o1;
    if (i0 >= i1) return;
    // G = [J_a | J_b | r | 0 0 0] (two rows per factor, 16 columns): G^T G holds the (a,a), (a,b), (b,b) blocks and both
    // gradients.  It is accumulated on the FP64 tensor cores, four rows (two factors) per step: the A fragment of a
    // column block and its B fragment are the same value G[k = lane % 4][8 blk + lane / 4], so a step is two
    // shared-memory reads and three DMMAs (blocks (0,0), (0,1), (1,1) of the upper triangle) per lane.
    // Offsets inside a staged record for a kind-2 item (a = anchor frame i: J_a = Ji at 2, J_b = Jj at 14) and a
    // kind-3 item (a = observing frame j: J_a = Jj, J_b = Ji); second row +6 (Jacobians) / +1 (residual).
    const int fcol = lane >> 2, frow = lane & 1, fsub = (lane >> 1) & 1;
    const int o0_k2 = (fcol < 6 ? 2 + fcol : 8 + fcol) + 6 * frow, o0_k3 = (fcol < 6 ? 14 + fcol : fcol - 4) + 6 * frow;
    const int o1_k2 = fcol < 4 ? 16 + fcol + 6 * frow : (fcol == 4 ? frow : -1);
    const int o1_k3 = fcol < 4 ? 4 + fcol + 6 * frow : (fcol == 4 ? frow : -1);
    double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
    for (int base = i0; base < i1; base += 32) {
      int kind = -1;
      if (base + lane < i1) {
        const int2 my = items[base + lane];
        kind = my.y;
        if (D.nranks > 1 && (D.proj_idx[my.x].z % D.nranks) != D.rank) kind = -1;   // factor-parallel: another rank's landmark
        if (kind >= 0) {   // every lane copies [r | Ji | Jj] of its own item's record: up to 32 records in flight per warp
          const double *rec = D.rec_proj + (size_t)my.x * REC_PROJ;
          double *dst = buf + lane * DSTR;
#pragma unroll
          for (int c = 0; c < 13; c++) cp_async16(dst + 2 * c, rec + 2 * c);
        }
      }
      cp_async_wait_all();
      __syncwarp();
      const int cnt = min(32, i1 - base);
#pragma unroll 2
      for (int s = 0; 2 * s < cnt; s++) {
        const int it = 2 * s + fsub;
        const int kd = __shfl_sync(full, kind, it);
        double g0 = 0.0, g1 = 0.0;
        if (kd >= 0) {
          const double *rec = buf + it * DSTR;
          const int o1 = kd == 3 ? o1_k3 : o1_k2;
          g0 = rec[kd == 3 ? o0_k3 : o0_k2];
          if (o1 >= 0) g1 = rec[o1];
        }
        dmma884(c00, g0, g0);
        dmma884(c01, g0, g1);
        dmma884(c11, g1, g1);
      }
      __syncwarp();
    }
    const int ra = 15 * a, rb = 15 * b;
    // entry (p, q) of G^T G, p <= q: columns 0..5 -> block a, 6..11 -> block b, 12 -> gradient
    auto put = [&](int p, int q, double v) {
      if (p >= 12 || q > 12) return;
      const int gp = p < 6 ? ra + p : rb + p - 6;
      if (q == 12) { atomicAdd(D.gS + co + gp, v); atomicAdd(D.gfull + co + gp, v); return; }
      if (p > q) return;
      const int gq = q < 6 ? ra + q : rb + q - 6;
      atomicAdd(Sg + (size_t)gp * d + gq, v);
      if (p == q) atomicAdd(D.colsq_cam + co + gp, v);
    };
    const int pr = lane >> 2, pc = 2 * (lane & 3);
    put(pr, pc, c00[0]); put(pr, pc + 1, c00[1]);
    put(pr, 8 + pc, c01[0]); put(pr, 9 + pc, c01[1]);
    put(8 + pr, 8 + pc, c11[0]); put(8 + pr, 9 + pc, c11[1]);
  } else {
    // line / VP factors of diagonal block a, one segment of the list
    if (a >= nb) return;
    int i0 = o0, i1 = o1;
    const int per = (i1 - i0 + SEGS_D - 1) / SEGS_D;
    i0 += seg * per; i1 = min(i1, i0 + per);
    if (i0 >= i1) return;
    // G = [J_pose (6) | r | 0], two rows per item (the second row of a VP item is zero): one DMMA per two items
    const int fcol = lane >> 2, frow = lane & 1, fsub = (lane >> 1) & 1;
    const int o_line = fcol < 6 ? 2 + 6 * frow + fcol : (fcol == 6 ? frow : -1);
    const int o_vp = frow ? -1 : (fcol < 6 ? 1 + fcol : (fcol == 6 ? 0 : -1));
    double cc[2] = {0.0, 0.0};
    for (int base = i0; base < i1; base += 32) {
      int mybase = -1;   // offset of the pose block inside the record (2: line, two rows; 1: VP, one row); -1 = skip
      if (base + lane < i1) {
        const int2 it = items[base + lane];
        const bool line = it.y == 7;
        bool mine = true;
        if (D.nranks > 1) mine = ((line ? D.line_idx4[it.x].y : D.vp_idx4[it.x].y) % D.nranks) == D.rank;
        if (mine) {
          double *dst = buf + lane * DSTR;
          if (line) {
            const double *rec = D.rec_line + (size_t)it.x * REC_LINE;   // [r(2) | Jpose 2x6]: 14 doubles, 16-byte aligned
#pragma unroll
            for (int c = 0; c < 7; c++) cp_async16(dst + 2 * c, rec + 2 * c);
          } else {
            const double *rec = D.rec_vp + (size_t)it.x * REC_VP;       // [r | Jpose 6]: 7 doubles
#pragma unroll
            for (int c = 0; c < 7; c++) cp_async8(dst + c, rec + c);
          }
          mybase = line ? 2 : 1;
        }
      }
      cp_async_wait_all();
      __syncwarp();
      const int cnt = min(32, i1 - base);
#pragma unroll 2
      for (int s = 0; 2 * s < cnt; s++) {
        const int it = 2 * s + fsub;
        const int bA = __shfl_sync(full, mybase, it);
        double g = 0.0;
        if (bA >= 0) {
          const int o = bA == 2 ? o_line : o_vp;
          if (o >= 0) g = buf[it * DSTR + o];
        }
        dmma884(cc, g, g);
      }
      __syncwarp();
    }
    const int ra = 15 * a;
    const int pr = lane >> 2, pc = 2 * (lane & 3);
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int p = pr, q = pc + e;
      if (p >= 6 || q > 6) continue;
      if (q == 6) { atomicAdd(D.gS + co + ra + p, cc[e]); atomicAdd(D.gfull + co + ra + p, cc[e]); }
      else if (p <= q) {
        atomicAdd(Sg + (size_t)(ra + p) * d + ra + q, cc[e]);
        if (p == q) atomicAdd(D.colsq_cam + co + ra + p, cc[e]);
      }
    }
  }
}

constexpr int SUP = 3;        // a warp owns a SUP x SUP block of 8x8 tiles of the rank update
constexpr int SUP_SETS = 2;   // and at most this many of them (8 warps x 2 covers the 10 blocks of 12 camera blocks)

// one chunk of the rank update for a 3 x 3 unit of 8x8 tiles: acc[u][v] += Y_rows(u) Y_rows(v)^T over c1 columns
template <unsigned MK>
__device__ __forceinline__ void rank_chunk(double (&acc)[SUP][SUP][2], const double *ya, const double *yb, const double *ysc, int c1, int mp) {
  constexpr unsigned rowm = (MK & 7u ? 1u : 0u) | (MK & 0x38u ? 2u : 0u) | (MK & 0x1c0u ? 4u : 0u), colm = (MK | MK >> 3 | MK >> 6) & 7u;
  for (int k0 = 0; k0 < c1; k0 += 4) {
    double fa[SUP], fb[SUP];
    const double cs = ysc[(size_t)k0 * mp];   // scale of this lane's column (entry mp - 1)
#pragma unroll
    for (int u = 0; u < SUP; u++) {
      if (rowm >> u & 1u) fa[u] = cs * ya[(size_t)k0 * mp + 8 * u];
      if (colm >> u & 1u) fb[u] = yb[(size_t)k0 * mp + 8 * u];
    }
#pragma unroll
    for (int u = 0; u < SUP; u++)
#pragma unroll
      for (int v = 0; v < SUP; v++)
        if (MK >> (3 * u + v) & 1u) dmma884(acc[u][v], fa[u], fb[v]);
  }
}
__device__ __forceinline__ void rank_chunk_any(unsigned mk, double (&acc)[SUP][SUP][2], const double *ya, const double *yb, const double *ysc, int c1, int mp) {
  for (int k0 = 0; k0 < c1; k0 += 4) {
    double fa[SUP], fb[SUP];
    const double cs = ysc[(size_t)k0 * mp];
#pragma unroll
    for (int u = 0; u < SUP; u++) { fa[u] = cs * ya[(size_t)k0 * mp + 8 * u]; fb[u] = yb[(size_t)k0 * mp + 8 * u]; }
#pragma unroll
    for (int u = 0; u < SUP; u++)
#pragma unroll
      for (int v = 0; v < SUP; v++)
        if (mk >> (3 * u + v) & 1u) dmma884(acc[u][v], fa[u], fb[v]);
  }
}

// cover of the 45 upper 8x8 tiles of a 9 x 9 tile grid by eight units {first row tile, first column tile, 3x3 mask}:
// 2x3 rectangles, the three diagonal triangles and one 1x3 strip - loads 12, 12, 12, 9 tiles on the four schedulers
__constant__ unsigned short c_units9[8][3] = {
    {0, 3, 0x03f}, {0, 6, 0x03f}, {2, 6, 0x03f}, {4, 6, 0x03f},   // rows 0-1 x cols 3-5 | rows 0-1 x 6-8 | rows 2-3 x 6-8 | rows 4-5 x 6-8
    {0, 0, 0x137}, {3, 3, 0x137}, {6, 6, 0x137},                  // upper triangles of tiles 0-2, 3-5, 6-8
    {2, 3, 0x007}};                                                // row 2 x cols 3-5

// Schur terms of one window as a dense rank update over its landmark columns, then IMU blocks and the prior.
//   [V | gsch] = Y [Y | z]^T  on the FP64 tensor cores: the columns (camera rows + the z row) are streamed chunk by
//   chunk with TMA bulk copies (cp.async.bulk + mbarrier, NSTAGE in flight); warps own 24 x 24 blocks of the upper
//   triangle and read the 8x4 / 4x8 fragments straight from the chunk (bank-conflict free for mp = 4 mod 16).
__global__ void __launch_bounds__(WT) k_window_system(Dev D, Stash S, int max_prior_n) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) unsigned long long bar[NSTAGE];
  const int w = blockIdx.x;
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  const int tid = threadIdx.x;
  const int fo = D.frame_off[w], F = D.frame_off[w + 1] - fo;
  const bool ex = (D.win_flags[w] & WF_EXTRINSIC) != 0;
  const int nb = F + (ex ? 1 : 0), m = 6 * nb, mp = S.mp;
  const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
  // shared layout: Ych[NSTAGE][CH][mp] (TMA destination, 16-byte aligned) + slack for fragment rows past mp
  double *Ych = sm;
  if (tid == 0) { for (int k = 0; k < NSTAGE; k++) mbar_init(&bar[k], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int e = tid; e < YSLACK; e += WT) Ych[(size_t)NSTAGE * CH * mp + e] = 0.0;
  __syncthreads();
  double *Sg = D.Smat + D.S_off[w];

  {
    const int np = D.point_off[w + 1] - D.point_off[w], nl = D.line_off[w + 1] - D.line_off[w];
    // small batches split a window's columns over gridDim.y CTAs (every CTA adds its partial update with reductions anyway)
    const int ncols_all = np + 4 * nl, nchunks_all = (ncols_all + CH - 1) / CH;
    const int cper = (nchunks_all + (int)gridDim.y - 1) / (int)gridDim.y, cbeg = (int)blockIdx.y * cper;
    if (cbeg >= nchunks_all) return;
    const int ncols = min(ncols_all - cbeg * CH, cper * CH), nchunks = (ncols + CH - 1) / CH;
    const double *Yw = S.Y + (colbase(D, w) + (size_t)cbeg * CH) * mp;
    auto issue = [&](int c) {   // one elected thread: arm the barrier with the byte count, start the bulk copy
      const int c0 = c * CH, c1 = min(ncols, c0 + CH);
      const unsigned bytes = (unsigned)((c1 - c0) * mp * sizeof(double));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar[c % NSTAGE], bytes);
      tma_bulk_g2s(Ych + (size_t)(c % NSTAGE) * CH * mp, Yw + (size_t)c0 * mp, bytes, &bar[c % NSTAGE]);
    };
    const int warp = tid >> 5, lane = tid & 31;
    const int zr = mp - 2;                           // the z row of a column (rows m .. zr-1 are zero padding)
    const int ntr = (zr + 1 + 7) / 8;                // 8-row tiles over the camera rows and the z row
    const int nsr = (ntr + SUP - 1) / SUP, nsuper = nsr * (nsr + 1) / 2;
    // work units: first row tile, first column tile and a 3 x 3 mask of 8x8 tiles (bit 3 u + v).  General shapes: SUP x SUP
    // super-blocks of the upper triangle in rank order.  The reference's window (nine row tiles: 11 frames x 6 + z) gets a
    // hand-balanced cover of the 45 upper tiles by EIGHT units (c_units9: seven of six tiles, one of three) - six
    // super-blocks of 6 / 9 tiles left two of the eight warps idle and two of the four schedulers with twice the DMMAs.
    int ta[SUP_SETS], tb[SUP_SETS];
    unsigned msk[SUP_SETS];
#pragma unroll
    for (int i = 0; i < SUP_SETS; i++) {
      ta[i] = tb[i] = 0; msk[i] = 0;
      if (ntr == 9 && WT == 256) {
        if (i == 0) { ta[0] = c_units9[warp][0]; tb[0] = c_units9[warp][1]; msk[0] = c_units9[warp][2]; }
      } else {
        const int idx = warp + i * (WT / 32);
        if (idx < nsuper) {
          int sa, sb;
          unrank_key(idx, sa, sb);
          ta[i] = SUP * sa; tb[i] = SUP * sb;
          for (int u = 0; u < SUP; u++)
            for (int v = 0; v < SUP; v++)
              if (!(sa == sb && v < u) && ta[i] + u < ntr && tb[i] + v < ntr) msk[i] |= 1u << (3 * u + v);
        }
      }
    }
    double acc[SUP_SETS][SUP][SUP][2];
#pragma unroll
    for (int i = 0; i < SUP_SETS; i++)
#pragma unroll
      for (int u = 0; u < SUP; u++)
#pragma unroll
        for (int v = 0; v < SUP; v++) acc[i][u][v][0] = acc[i][u][v][1] = 0.0;
    if (tid == 0) for (int c = 0; c < NSTAGE - 1 && c < nchunks; c++) issue(c);
    for (int c = 0; c < nchunks; c++) {
      // the stage of chunk c + NSTAGE - 1 held chunk c - 1, released by the barrier at the end of the last iteration
      if (tid == 0 && c + NSTAGE - 1 < nchunks) issue(c + NSTAGE - 1);
      mbar_wait(&bar[c % NSTAGE], (unsigned)((c / NSTAGE) & 1));
      double *Yb = Ych + (size_t)(c % NSTAGE) * CH * mp;
      const int c1 = min(ncols, (c + 1) * CH) - c * CH;
      if (c1 & 3) {   // last chunk: zero columns up to the next multiple of the MMA depth
        const int pad = ((c1 + 3) & ~3) - c1;
        for (int e = tid; e < pad * mp; e += WT) Yb[(size_t)c1 * mp + e] = 0.0;
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < SUP_SETS; i++) {
        const unsigned mk = msk[i];
        if (!mk) continue;
        const double *ya = Yb + (size_t)(lane & 3) * mp + (lane >> 2) + 8 * ta[i];
        const double *yb = Yb + (size_t)(lane & 3) * mp + (lane >> 2) + 8 * tb[i];
        const double *ysc = Yb + (size_t)(lane & 3) * mp + mp - 1;   // column scale (uvs_stash.cuh)
        // the tile mask is uniform over the warp: the common masks get loops with exactly their loads and DMMAs compiled
        // in (a predicated-off DMMA still takes its issue slot and tensor-pipe cycles)
        switch (mk) {
          case 0x03fu: rank_chunk<0x03fu>(acc[i], ya, yb, ysc, c1, mp); break;
          case 0x137u: rank_chunk<0x137u>(acc[i], ya, yb, ysc, c1, mp); break;
          case 0x007u: rank_chunk<0x007u>(acc[i], ya, yb, ysc, c1, mp); break;
          case 0x1ffu: rank_chunk<0x1ffu>(acc[i], ya, yb, ysc, c1, mp); break;
          default: rank_chunk_any(mk, acc[i], ya, yb, ysc, c1, mp); break;
        }
      }
      __syncthreads();
    }
    // ---- add the Schur part to the window's system (cleared by k_step / k_solve_init): fire-and-forget reductions,
    //      because the direct-term and IMU / prior kernels run side by side with this one on other streams
    auto grow = [&](int e) { const int a = e / 6; return (a < F ? 15 * a : 15 * F) + (e - 6 * a); };
#pragma unroll
    for (int i = 0; i < SUP_SETS; i++) {
      if (!msk[i]) continue;
#pragma unroll
      for (int u = 0; u < SUP; u++)
#pragma unroll
        for (int v = 0; v < SUP; v++) {
          if (!(msk[i] >> (3 * u + v) & 1u)) continue;
          const int row = 8 * (ta[i] + u) + (lane >> 2);
          if (row >= m) continue;
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int col = 8 * (tb[i] + v) + 2 * (lane & 3) + e;
            if (col < m) { if (row <= col) atomicAdd(Sg + (size_t)grow(row) * d + grow(col), -acc[i][u][v][e]); }
            else if (col == zr) atomicAdd(D.gS + co + grow(row), -acc[i][u][v][e]);
          }
        }
    }
  }
}

// IMU blocks and the prior of one window, added to the reduced system after k_window_system has stored the Schur part
// (a kernel of its own: the work is a handful of L2 round trips per window, which the two resident CTAs per SM of the
// rank-update kernel cannot hide, while here six to eight windows share an SM).  All additions are fire-and-forget
// FP64 reductions, so the IMU factors of a window are processed in one pass.
constexpr int TT = 256;       // threads of k_window_tail
constexpr int IMU_G = 4;      // IMU records staged per pass (15 KB: ten or more CTAs per SM)

__global__ void __launch_bounds__(TT) k_window_tail(Dev D, int max_prior_n, int ni) {
  extern __shared__ __align__(16) double sm[];
  const int w = blockIdx.x;
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  const int tid = threadIdx.x;
  const int fo = D.frame_off[w];
  const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
  double *Sg = D.Smat + D.S_off[w];
  // blockIdx.y: parts 0 .. ni-1 take the IMU factors (groups of IMU_G, round robin), the remaining np parts slices of the prior
  // (large batches: ni = np = 1; a handful of windows: several CTAs per window)
  const int part = blockIdx.y, np = (int)gridDim.y - ni;
  const int f0 = D.imu_off[w], nf = part < ni ? D.imu_off[w + 1] - f0 : 0;
  for (int g0 = part * IMU_G; g0 < nf; g0 += ni * IMU_G) {
    const int ng = min(IMU_G, nf - g0);
    const double *R = D.rec_imu + (size_t)(f0 + g0) * REC_IMU;
    for (int e = tid; e < ng * REC_IMU; e += TT) cp_async8(sm + e, R + e);
    cp_async_wait_all();
    __syncthreads();
    // G = [J (15 x 30) | r | 0], padded to 16 x 32: G^T G holds the upper triangle of the 30x30 block and, in column 30,
    // the gradient J^T r.  A warp takes half of a factor's ten upper 8x8 blocks (block rows {0, 3} or {1, 2}, five
    // blocks each) on the FP64 tensor cores: the A fragment of a column block and its B fragment are the same value
    // G[4 ks + lane % 4][8 blk + lane / 4].
    {
      const int warp = tid >> 5, lane = tid & 31, fr = lane & 3, fc = lane >> 2;
      for (int u = warp; u < 2 * ng; u += TT / 32) {
        const int k = u >> 1, hsel = u & 1;
        const double *r = sm + k * REC_IMU, *J = r + 15;
        const int c0 = 15 * (D.imu_idx[f0 + g0 + k].x - fo);
        auto frag = [&](int blk, int ks) {
          const int row = 4 * ks + fr, col = 8 * blk + fc;
          return row < 15 ? (col < 30 ? J[row * 30 + col] : (col == 30 ? r[row] : 0.0)) : 0.0;
        };
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
          const int I = hsel == 0 ? (pass == 0 ? 0 : 3) : (pass == 0 ? 1 : 2);
          double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const double a = frag(I, ks);
#pragma unroll
            for (int Jb = 0; Jb < 4; Jb++) {
              if (Jb < I) continue;   // uniform over the warp
              const double bv = Jb == I ? a : frag(Jb, ks);
              dmma884(acc[Jb], a, bv);
            }
          }
          const int p = 8 * I + fc;   // row of G^T G held by this lane
          if (p < 30) {
#pragma unroll
            for (int Jb = 0; Jb < 4; Jb++) {
              if (Jb < I) continue;
#pragma unroll
              for (int e = 0; e < 2; e++) {
                const int q = 8 * Jb + 2 * fr + e;
                const double v = acc[Jb][e];
                if (q < 30) {
                  if (p <= q) {
                    atomicAdd(Sg + (size_t)(c0 + p) * d + c0 + q, v);
                    if (p == q) atomicAdd(D.colsq_cam + co + c0 + p, v);
                  }
                } else if (q == 30) { atomicAdd(D.gfull + co + c0 + p, v); atomicAdd(D.gS + co + c0 + p, v); }
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // prior: H += J0^T J0 (precomputed at upload), g += J0^T r
  const int n = part >= ni ? D.prior_off[w + 1] - D.prior_off[w] : 0;
  if (n > 0) {
    const int sl = part - ni;   // slice of the prior
    int *cmap = reinterpret_cast<int *>(sm);
    for (int c = tid; c < n; c += TT) cmap[c] = -1;
    __syncthreads();
    for (int b = D.pblk_off[w] + tid; b < D.pblk_off[w + 1]; b += TT) {
      const int kind = D.pblk_kind[b], cam = D.pblk_cam[b], col = D.pblk_col[b];
      const int ls = (kind == 0 || kind == 2) ? 6 : (kind == 1 ? 9 : 1);
      if (cam >= 0) for (int c = 0; c < ls; c++) cmap[col + c] = cam + c;
    }
    __syncthreads();
    const double *H = D.prior_H + D.priorJ_off[w], *J0 = D.prior_J + D.priorJ_off[w], *r = D.rec_prior + D.prior_off[w];
    for (int e0 = sl * 4 * TT + tid; e0 < n * n; e0 += np * 4 * TT) {   // four loads in flight, then the four reductions
      double hv[4];
      int at[4];
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const int e = e0 + v * TT;
        at[v] = -1;
        hv[v] = 0.0;
        if (e < n * n) {
          const int p = e / n, q = e - p * n;
          const int cp = cmap[p], cq = cmap[q];
          if (cp >= 0 && cq >= 0 && cp <= cq) { at[v] = cp * d + cq; hv[v] = __ldg(H + e); }
        }
      }
#pragma unroll
      for (int v = 0; v < 4; v++) if (at[v] >= 0) atomicAdd(Sg + at[v], hv[v]);
    }
    // gradient: thread (p, part) sums every NP-th row of column p (coalesced over p), eight loads in flight; NP is chosen
    // so that NP n <= TT: one pass over the threads (4 n = 300 > 256 made a second, serial pass of 44 threads)
    const int NP = max(1, min(4, TT / n));
    for (int u = tid; u < NP * n; u += TT) {
      const int part = u / n, p = u - part * n;
      const int cp = cmap[p];
      if (cp < 0) continue;
      double g4[4] = {0, 0, 0, 0};
      const int RS = NP * np;   // row stride: NP parts in this CTA x np slices
      int i = sl * NP + part;
      for (; i + 7 * RS < n; i += 8 * RS) {
        double v8[8];
#pragma unroll
        for (int v = 0; v < 8; v++) v8[v] = __ldg(J0 + (size_t)(i + v * RS) * n + p);
#pragma unroll
        for (int v = 0; v < 8; v++) g4[v & 3] += v8[v] * r[i + v * RS];
      }
      for (; i < n; i += RS) g4[0] += __ldg(J0 + (size_t)i * n + p) * r[i];
      const double gg = (g4[0] + g4[1]) + (g4[2] + g4[3]);
      atomicAdd(D.gfull + co + cp, gg); atomicAdd(D.gS + co + cp, gg);
      if (part == 0 && sl == 0) atomicAdd(D.colsq_cam + co + cp, __ldg(H + (size_t)p * n + p));
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct Build3Ctx {
  Stash S;
  DirectLists L;
};

size_t build3_bytes(const Dev &D, int max_frames, bool any_ex, Build3Layout *lay) {
  const int nbmax = max_frames + (any_ex ? 1 : 0);
  lay->mp = 6 * nbmax + 2;
  while (lay->mp % 16 != 4) lay->mp++;   // fragment loads of the tensor-core rank update are bank-conflict free for mp = 4 (mod 16)
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t o = 0;
  lay->o_Y = o; o += al(((size_t)D.nP + 4 * (size_t)D.nL) * lay->mp * sizeof(double));
  lay->o_ph = o; o += al((size_t)D.nP * 4 * sizeof(double));
  lay->o_lh = o; o += al((size_t)D.nL * 24 * sizeof(double));
  lay->o_items = o; o += al(((size_t)6 * D.nProj + D.nLobs + D.nVobs + 1) * sizeof(int2));
  lay->o_off = o; o += al((size_t)D.B * KMAX * sizeof(int));
  return o;
}

static void make_ctx(char *base, const Build3Layout &lay, Build3Ctx &c) {
  c.S.Y = (double *)(base + lay.o_Y); c.S.ph = (double *)(base + lay.o_ph); c.S.lh = (double *)(base + lay.o_lh);
  c.S.mp = lay.mp; c.S.unscaled_pts = 0;
  c.L.items = (int2 *)(base + lay.o_items); c.L.off = (int *)(base + lay.o_off);
}

static inline int cdiv3(int a, int b) { return (a + b - 1) / b; }

// column scales of the stash start as 1 (the columns themselves as zeros: cudaMemset at upload)
__global__ void k_stash_init(double *Y, long long ncols, int mp) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c < ncols) Y[c * mp + mp - 1] = 1.0;
}

int launch_stash_init(const Dev &D, char *base, const Build3Layout &lay, cudaStream_t st) {
  const long long ncols = (long long)D.nP + 4LL * D.nL;
  if (ncols == 0) return 0;
  k_stash_init<<<(unsigned)((ncols + 255) / 256), 256, 0, st>>>((double *)(base + lay.o_Y), ncols, lay.mp);
  return 1;
}

int launch_build3_prep(const Dev &D, char *base, const Build3Layout &lay, bool any_ex, cudaStream_t st) {
  Build3Ctx c; make_ctx(base, lay, c);
  k_prep_direct<<<cdiv3(D.B, 4), 128, 0, st>>>(D, c.L, any_ex ? 0 : 1);
  return 1;
}

size_t build3_smem(int max_frames, bool any_ex, int max_prior_n) {
  const int nb = max_frames + (any_ex ? 1 : 0), m = 6 * nb;
  int mp = m + 2;
  while (mp % 16 != 4) mp++;
  (void)m;
  return ((size_t)NSTAGE * CH * mp + YSLACK) * sizeof(double) + (size_t)(max_prior_n + 2) * sizeof(int);
}

// CTAs per window of the per-window kernels: one for batches that fill the GPU, several for a handful of windows (the
// reference's use is ONE window per frame: a single CTA would stream all its landmark columns on one SM)
static int window_split(int B) { return B >= 74 ? 1 : std::min(8, cdiv3(148, 2 * B)); }

int launch_build3(const Dev &D, const Params &P, char *base, const Build3Layout &lay, int max_frames, bool any_ex, int max_prior_n,
                  cudaStream_t st, const Fork *fk) {
  Build3Ctx c; make_ctx(base, lay, c);
  int n = 0;
  // four independent strands: point elimination | line elimination | direct terms | IMU + prior; the rank update needs
  // the first two.  With auxiliary streams (fk) they run side by side, otherwise one after the other on st.
  cudaStream_t s_lines = fk ? fk->aux[0] : st, s_direct = fk ? fk->aux[1] : st, s_tail = fk ? fk->aux[2] : st;
  if (fk) fork_from(fk, st, 3);
  if (D.nP) { k_core_points<<<cdiv3(D.nP, 128 / LPP), 128, 0, st>>>(D, P, c.S); n++; }
  if (D.nL) { k_core_lines<<<cdiv3(D.nL, 128 / LPL), 128, 0, s_lines>>>(D, P, c.S); n++; }
  if (any_ex) {
    const int nb_max = max_frames + 1;
    const long long units = (long long)D.B * (nb_max * SEGS + nb_max * (nb_max - 1) / 2);
    k_direct<<<(unsigned)((units + 3) / 4), 128, 0, s_direct>>>(D, c.L, nb_max);
  } else {
    const long long units = (long long)D.B * (max_frames * SEGS_D + max_frames * (max_frames - 1) / 2);
    k_direct_fused<<<(unsigned)((units + 3) / 4), 128, 0, s_direct>>>(D, c.L, max_frames);
  }
  n++;
  if (D.nranks <= 1 || D.rank == 0) {   // factor-parallel mode: IMU factors and the prior belong to rank 0
    const size_t tsm = std::max((size_t)IMU_G * REC_IMU * sizeof(double), (size_t)(max_prior_n + 2) * sizeof(int));
    { const int ws = window_split(D.B), ni = ws > 1 ? 3 : 1, np = ws > 1 ? 4 : 1; k_window_tail<<<dim3(D.B, ni + np), TT, tsm, s_tail>>>(D, max_prior_n, ni); }
    n++;
  }
  if (fk) join_to(fk, st, 0);
  const size_t smem = build3_smem(max_frames, any_ex, max_prior_n);
  static size_t raised = 0;
  if (smem > raised) { cudaFuncSetAttribute(k_window_system, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); raised = smem; }
  k_window_system<<<dim3(D.B, window_split(D.B)), WT, smem, st>>>(D, c.S, max_prior_n);
  n++;
  if (fk) { join_to(fk, st, 1); join_to(fk, st, 2); }
  return n;
}

// Record path, sweep and build stage as ONE dependency graph over the four streams (no join / fork between the two stages):
//   main   k_proj<1> -> k_core_points -> [lines] -> k_window_system
//   aux 0  k_line_vp<1> -> k_core_lines
//   aux 1  IMU sweep -> [prior] -> k_window_tail
//   aux 2  prior sweep -> [proj, lines] -> k_direct_fused
// For a handful of windows the kernels are ~10 us each and the stage barrier cost as much as a kernel: one window 1.61 ->
// 1.54 ms per solve, sixteen 1.75 -> 1.69 ms (constant extrinsic only; the caller keeps the two-stage form for profiling).
// Merging the back-substitution and candidate-cost stages the same way was measured too and bought nothing.
int launch_record_linearisation(const Dev &D, const Params &P, char *base, const Build3Layout &lay, int max_frames, int max_prior_n,
                                cudaStream_t st, const Fork *fk) {
  Build3Ctx c; make_ctx(base, lay, c);
  int n = 0;
  double *cost0 = D.acc + ACC_COST0;
  cudaStream_t s0 = fk->aux[0], s1 = fk->aux[1], s2 = fk->aux[2];
  fork_from(fk, st, 3);
  n += launch_proj(D, P, true, false, 1, 0, D.rec_proj, nullptr, cost0, ACC_STRIDE, st);
  n += launch_line_vp(D, P, true, 1, 0, D.rec_line, D.rec_vp, cost0, ACC_STRIDE, s0);
  n += launch_imu(D, P, true, 1, 0, D.rec_imu, nullptr, cost0, ACC_STRIDE, s1);
  n += launch_prior(D, max_prior_n, true, 1, 0, D.rec_prior, cost0, ACC_STRIDE, s2);
  // a wait refers to the recording of an event that precedes it: the auxiliary streams have already been told to wait for
  // the first recording of `fork`, so the event can be recorded again
  cudaEventRecord(fk->fork, st);        // projection records written
  cudaEventRecord(fk->join[0], s0);     // line / VP records written
  cudaEventRecord(fk->join[2], s2);     // prior residual written
  cudaStreamWaitEvent(s1, fk->join[2], 0);
  cudaStreamWaitEvent(s2, fk->fork, 0);
  cudaStreamWaitEvent(s2, fk->join[0], 0);
  if (D.nP) { k_core_points<<<cdiv3(D.nP, 128 / LPP), 128, 0, st>>>(D, P, c.S); n++; }
  if (D.nL) { k_core_lines<<<cdiv3(D.nL, 128 / LPL), 128, 0, s0>>>(D, P, c.S); n++; }
  {
    const long long units = (long long)D.B * (max_frames * SEGS_D + max_frames * (max_frames - 1) / 2);
    k_direct_fused<<<(unsigned)((units + 3) / 4), 128, 0, s2>>>(D, c.L, max_frames);
    n++;
  }
  if (D.nranks <= 1 || D.rank == 0) {
    const size_t tsm = std::max((size_t)IMU_G * REC_IMU * sizeof(double), (size_t)(max_prior_n + 2) * sizeof(int));
    const int ws = window_split(D.B), ni = ws > 1 ? 3 : 1, np = ws > 1 ? 4 : 1;
    k_window_tail<<<dim3(D.B, ni + np), TT, tsm, s1>>>(D, max_prior_n, ni);
    n++;
  }
  join_to(fk, st, 0);
  const size_t smem = build3_smem(max_frames, false, max_prior_n);
  static size_t raised = 0;
  if (smem > raised) { cudaFuncSetAttribute(k_window_system, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); raised = smem; }
  k_window_system<<<dim3(D.B, window_split(D.B)), WT, smem, st>>>(D, c.S, max_prior_n);
  n++;
  join_to(fk, st, 1); join_to(fk, st, 2);
  return n;
}

// Linearisation stage of the fused path: IMU + prior sweeps (the only factor records that still exist), point and line
// linearisation with the factors evaluated in registers (uvs_lin.cu), IMU / prior tail, rank update.  Three strands:
// main = points -> rank update, aux 0 = lines, aux 1 = IMU sweep -> prior sweep -> tail.
int launch_build3_fused(const Dev &D, const Params &P, char *base, const Build3Layout &lay, int max_frames, int max_lines, int max_prior_n,
                        cudaStream_t st, const Fork *fk) {
  Build3Ctx c; make_ctx(base, lay, c);
  int n = 0;
  cudaStream_t s_lines = fk ? fk->aux[0] : st, s_tail = fk ? fk->aux[1] : st, s_prior = fk ? fk->aux[2] : st;
  double *cost0 = D.acc + ACC_COST0;
  if (fk) fork_from(fk, st, 3);
  n += launch_imu(D, P, true, 1, 0, D.rec_imu, nullptr, cost0, ACC_STRIDE, s_tail);
  n += launch_prior(D, max_prior_n, true, 1, 0, D.rec_prior, cost0, ACC_STRIDE, s_prior);   // beside the IMU sweep; the tail needs both
  if (fk) { cudaEventRecord(fk->join[2], s_prior); cudaStreamWaitEvent(s_tail, fk->join[2], 0); }
  n += launch_lin_points(D, P, base, lay, st);
  n += launch_lin_lines(D, P, base, lay, max_frames, max_lines, s_lines);
  if (D.nranks <= 1 || D.rank == 0) {   // factor-parallel mode: IMU factors and the prior belong to rank 0
    const size_t tsm = std::max((size_t)IMU_G * REC_IMU * sizeof(double), (size_t)(max_prior_n + 2) * sizeof(int));
    { const int ws = window_split(D.B), ni = ws > 1 ? 3 : 1, np = ws > 1 ? 4 : 1; k_window_tail<<<dim3(D.B, ni + np), TT, tsm, s_tail>>>(D, max_prior_n, ni); }
    n++;
  }
  if (fk) join_to(fk, st, 0);
  const size_t smem = build3_smem(max_frames, false, max_prior_n);
  static size_t raised = 0;
  if (smem > raised) { cudaFuncSetAttribute(k_window_system, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); raised = smem; }
  k_window_system<<<dim3(D.B, window_split(D.B)), WT, smem, st>>>(D, c.S, max_prior_n);
  n++;
  if (fk) join_to(fk, st, 1);
  return n;
}

int launch_back3(const Dev &D, char *base, const Build3Layout &lay, cudaStream_t st, const Fork *fk, bool unscaled_pts) {
  Build3Ctx c; make_ctx(base, lay, c);
  c.S.unscaled_pts = unscaled_pts ? 1 : 0;   // the fused point kernel leaves its columns unscaled (uvs_stash.cuh)
  int n = 0;
  if (fk) fork_from(fk, st, 1);
  if (D.nP) { k_back_points<<<cdiv3(D.nP, 128 / LPP), 128, 0, st>>>(D, c.S); n++; }
  if (D.nL) { k_back_lines<<<cdiv3(D.nL, 128 / LPL), 128, 0, fk ? fk->aux[0] : st>>>(D, c.S); n++; }
  if (fk) join_to(fk, st, 0);
  return n;
}

}  // namespace uvs
