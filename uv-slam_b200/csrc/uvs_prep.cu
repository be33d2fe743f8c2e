// uvs_prep.cu — once-per-upload device preparation and the API export kernels.
//
//   k_prep_*      derive global index records from the caller's per-window arrays, validate them
//   k_imu_info    sqrt_info = LLT(cov^-1).matrixL()^T per IMU factor (imu_factor.h:64) — the reference
//                 recomputes it in every Evaluate(); the constants are frozen during a solve
//                 (imu_factor.h:52-58), so it is computed once here
//   k_prior_H     J0^T J0 of every window's prior (constant during a solve)
//   k_split / k_export_imu   record arrays -> the [n][nres] / [n][jac] arrays of the C ABI
#include <algorithm>

#include "uvs_device.cuh"
#include "uvs_kernels.h"

namespace uvs {

// window of item `i` given the offset table off[0..B]
__device__ __forceinline__ int find_window(const int *__restrict__ off, int B, int i) {
  int lo = 0, hi = B;   // off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void flag_error(const Dev &D, int code) { atomicMax(D.err, code); }

__global__ void k_prep_proj(Dev D) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= D.nProj) return;
  const int w = find_window(D.proj_off, D.B, f);
  const int F = D.frame_off[w + 1] - D.frame_off[w], np = D.point_off[w + 1] - D.point_off[w];
  const int fi = D.proj_fi[f], fj = D.proj_fj[f], pt = D.proj_pt[f];
  if (fi < 0 || fi >= F || fj < 0 || fj >= F || pt < 0 || pt >= np || fi == fj) { flag_error(D, 1); return; }
  const int gp = D.point_off[w] + pt;
  D.proj_idx[f] = make_int4(D.frame_off[w] + fi, D.frame_off[w] + fj, gp, w);
  const bool first = (f == D.proj_off[w]) || (D.proj_pt[f - 1] != pt);
  const bool last = (f + 1 == D.proj_off[w + 1]) || (D.proj_pt[f + 1] != pt);
  if (first) {
    if (atomicExch(D.pt_begin + gp, f) != RANGE_UNSET) flag_error(D, 2);  // factors of a point not contiguous
  }
  if (last) atomicExch(D.pt_end + gp, f + 1);
  if (!first && D.proj_fi[f - 1] != fi) flag_error(D, 3);   // one anchor frame per point
  D.pt_win[gp] = w;
}

__global__ void k_prep_line(Dev D) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= D.nLobs) return;
  const int w = find_window(D.lobs_off, D.B, f);
  const int F = D.frame_off[w + 1] - D.frame_off[w], nl = D.line_off[w + 1] - D.line_off[w];
  const int fj = D.line_frame[f], lk = D.line_idx[f];
  if (fj < 0 || fj >= F || lk < 0 || lk >= nl) { flag_error(D, 4); return; }
  const int gl = D.line_off[w] + lk;
  D.line_idx4[f] = make_int4(D.frame_off[w] + fj, gl, w, -1);
  const bool first = (f == D.lobs_off[w]) || (D.line_idx[f - 1] != lk);
  const bool last = (f + 1 == D.lobs_off[w + 1]) || (D.line_idx[f + 1] != lk);
  if (first) {
    if (atomicExch(D.ln_begin + gl, f) != RANGE_UNSET) flag_error(D, 5);
  }
  if (last) atomicExch(D.ln_end + gl, f + 1);
  D.ln_win[gl] = w;
}

// runs after k_prep_line: pair every VP observation with the line observation of the same (frame, line)
__global__ void k_prep_vp(Dev D) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= D.nVobs) return;
  const int w = find_window(D.vobs_off, D.B, f);
  const int F = D.frame_off[w + 1] - D.frame_off[w], nl = D.line_off[w + 1] - D.line_off[w];
  const int fj = D.vp_frame[f], lk = D.vp_line[f];
  if (fj < 0 || fj >= F || lk < 0 || lk >= nl) { flag_error(D, 6); return; }
  const int gl = D.line_off[w] + lk;
  int lobs = -1;
  const int b = D.ln_begin[gl], e = D.ln_end[gl];
  if (b != RANGE_UNSET)
    for (int k = b; k < e; k++) if (D.line_frame[k] == fj) { lobs = k; break; }
  if (lobs < 0) { flag_error(D, 7); return; }   // VP factor without its line factor (estimator.cpp:916-925 adds both)
  D.vp_idx4[f] = make_int4(D.frame_off[w] + fj, gl, w, lobs);
  if (atomicExch(&D.line_idx4[lobs].w, f) != -1) flag_error(D, 8);  // two VP factors on one observation
}

__global__ void k_prep_fix_ranges(Dev D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D.nP && D.pt_begin[i] == RANGE_UNSET) { D.pt_begin[i] = 0; D.pt_end[i] = 0; D.pt_win[i] = find_window(D.point_off, D.B, i); }
  if (i < D.nL && D.ln_begin[i] == RANGE_UNSET) { D.ln_begin[i] = 0; D.ln_end[i] = 0; D.ln_win[i] = find_window(D.line_off, D.B, i); }
}

// IMU index records + sqrt_info.  The inverse (LU, partial pivoting) and the Cholesky factor use
// un-fused multiplies/adds so that the result does not depend on FMA contraction.
constexpr int II_WPC = 4;   // warps (= IMU factors) per CTA of k_imu_info
// One WARP per factor, the 15 x 15 matrices in shared memory (row stride 15: odd, conflict-free): lane i owns row i of the
// LU elimination and of the Cholesky factor, lane c column c of the inverse.  Every matrix entry sees exactly the
// operations, in the order, of the plain sequential algorithm (un-fused multiplies / subtractions), so the result is the
// one the single-thread version produced - that one kept two 225-double arrays in local memory and took 159 us for the ten
// factors of one window (299 us for a batch of 11 840).
__global__ void __launch_bounds__(32 * II_WPC) k_imu_info(Dev D) {
  __shared__ double s_lu[II_WPC][232], s_inv[II_WPC][232];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * II_WPC + warp;
  if (f >= D.nImu) return;
  const int w = find_window(D.imu_off, D.B, f);
  const int F = D.frame_off[w + 1] - D.frame_off[w];
  const int fi = D.imu_frame[f];
  if (fi < 0 || fi + 1 >= F) { if (lane == 0) flag_error(D, 9); return; }
  if (lane == 0) D.imu_idx[f] = make_int2(D.frame_off[w] + fi, w);
  const double *cov = D.imu_cov + 225 * (size_t)f;
  double *lu = s_lu[warp], *inv = s_inv[warp];
  for (int e = lane; e < 225; e += 32) lu[e] = cov[e];
  __syncwarp();
  const unsigned full = 0xffffffffu;
  const bool row = lane < 15;
  int piv[15];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 15; k++) {
    // partial pivoting: first row i >= k with the largest |lu[i][k]|
    double v = (row && lane >= k) ? fabs(lu[lane * 15 + k]) : -1.0;
    int idx = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double v2 = __shfl_xor_sync(full, v, o);
      const int i2 = __shfl_xor_sync(full, idx, o);
      if (v2 > v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
    }
    piv[k] = idx;
    if (v == 0.0) { ok = false; break; }
    if (idx != k && row) { const double t = lu[k * 15 + lane]; lu[k * 15 + lane] = lu[idx * 15 + lane]; lu[idx * 15 + lane] = t; }
    __syncwarp();
    if (row && lane > k) {
      const double l = lu[lane * 15 + k] / lu[k * 15 + k];
      lu[lane * 15 + k] = l;
      for (int j = k + 1; j < 15; j++) lu[lane * 15 + j] = __dsub_rn(lu[lane * 15 + j], __dmul_rn(l, lu[k * 15 + j]));
    }
    __syncwarp();
  }
  if (ok && row) {   // column `lane` of the inverse
    double x[15];
#pragma unroll
    for (int i = 0; i < 15; i++) x[i] = i == lane ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 15; k++) {
      const int p = piv[k];
      if (p != k) {   // x[k] <-> x[p], p > k (register array: resolved by selects)
        const double xk = x[k];
        double xp = 0.0;
#pragma unroll
        for (int i = 0; i < 15; i++) if (i == p) xp = x[i];
        x[k] = xp;
#pragma unroll
        for (int i = 0; i < 15; i++) if (i == p) x[i] = xk;
      }
    }
#pragma unroll
    for (int i = 0; i < 15; i++) { double sacc = x[i]; for (int j = 0; j < i; j++) sacc = __dsub_rn(sacc, __dmul_rn(lu[i * 15 + j], x[j])); x[i] = sacc; }
#pragma unroll
    for (int i = 14; i >= 0; i--) { double sacc = x[i]; for (int j = i + 1; j < 15; j++) sacc = __dsub_rn(sacc, __dmul_rn(lu[i * 15 + j], x[j])); x[i] = sacc / lu[i * 15 + i]; }
#pragma unroll
    for (int i = 0; i < 15; i++) inv[i * 15 + lane] = x[i];
  }
  __syncwarp();
  // lower Cholesky of inv (reads the lower triangle), stored transposed -> upper sqrt_info
  double *L = lu;
  for (int e = lane; e < 225; e += 32) L[e] = 0.0;
  __syncwarp();
  for (int j = 0; j < 15 && ok; j++) {
    double d = inv[j * 15 + j];
    for (int k = 0; k < j; k++) d = __dsub_rn(d, __dmul_rn(L[j * 15 + k], L[j * 15 + k]));
    if (!(d > 0.0)) { ok = false; break; }
    d = sqrt(d);
    __syncwarp();
    if (lane == j) L[j * 15 + j] = d;
    if (row && lane > j) {
      double sacc = inv[lane * 15 + j];
      for (int k = 0; k < j; k++) sacc = __dsub_rn(sacc, __dmul_rn(L[lane * 15 + k], L[j * 15 + k]));
      L[lane * 15 + j] = sacc / d;
    }
    __syncwarp();
  }
  if (!ok) { if (lane == 0) flag_error(D, 10); return; }
  double *out = D.imu_sqrt_info + 225 * (size_t)f;
  for (int e = lane; e < 225; e += 32) { const int i = e / 15, j = e - 15 * i; out[e] = L[j * 15 + i]; }
}

// prior_H = J0^T J0, one CTA per window
// J0^T J0 of a window's prior: a thread owns a 4 x 4 tile (eight loads per sixteen multiply-adds, J0 through L1), the
// tiles of a window are spread over gridDim.y CTAs (a handful of windows: several CTAs per window)
__global__ void __launch_bounds__(256) k_prior_H(Dev D) {
  const int w = blockIdx.x;
  const int n = D.prior_off[w + 1] - D.prior_off[w];
  if (n <= 0) return;
  const double *J0 = D.prior_J + D.priorJ_off[w];
  double *H = D.prior_H + D.priorJ_off[w];
  const int nt = (n + 3) / 4;
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < nt * nt; t += gridDim.y * blockDim.x) {
    const int p0 = 4 * (t / nt), q0 = 4 * (t % nt);
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
    for (int i = 0; i < n; i++) {
      const double *r = J0 + (size_t)i * n;
      double pa[4], qa[4];
#pragma unroll
      for (int a = 0; a < 4; a++) { pa[a] = p0 + a < n ? __ldg(r + p0 + a) : 0.0; qa[a] = q0 + a < n ? __ldg(r + q0 + a) : 0.0; }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[a][c] += pa[a] * qa[c];
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int c = 0; c < 4; c++) if (p0 + a < n && q0 + c < n) H[(size_t)(p0 + a) * n + q0 + c] = acc[a][c];
  }
}

// rec[n][REC] = [r(NR) | J(REC-NR)]  ->  r_out[n][NR], J_out[n][REC-NR] (either may be null)
__global__ void k_split(const double *__restrict__ rec, long long n, int REC, int NR, double *__restrict__ r_out,
                        double *__restrict__ J_out) {
  const long long total = n * REC;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long f = e / REC;
    const int c = (int)(e - f * REC);
    if (c < NR) { if (r_out) r_out[f * NR + c] = rec[e]; }
    else if (J_out) J_out[f * (REC - NR) + (c - NR)] = rec[e];
  }
}

// IMU record [r(15) | J 15x30 row-major] -> r_out[n][15], J_out[n][blocks concatenated]
// pose width PW = 6 (local layout) or 7 (Ceres layout, 7th column zero): blocks 15xPW, 15x9, 15xPW, 15x9.
__global__ void k_export_imu(const double *__restrict__ rec, int n, int PW, double *__restrict__ r_out,
                             double *__restrict__ J_out) {
  const int f = blockIdx.x;
  if (f >= n) return;
  const double *R = rec + (size_t)f * REC_IMU;
  const int jd = 15 * (2 * PW + 18);
  if (r_out) for (int i = threadIdx.x; i < 15; i += blockDim.x) r_out[15 * (size_t)f + i] = R[i];
  if (!J_out) return;
  double *J = J_out + (size_t)f * jd;
  const int bw[4] = {PW, 9, PW, 9}, lw[4] = {6, 9, 6, 9}, lo[4] = {0, 6, 15, 21};
  int base = 0;
  for (int b = 0; b < 4; b++) {
    for (int e = threadIdx.x; e < 15 * bw[b]; e += blockDim.x) {
      const int i = e / bw[b], c = e - i * bw[b];
      J[base + e] = c < lw[b] ? R[15 + i * 30 + lo[b] + c] : 0.0;
    }
    base += 15 * bw[b];
  }
}

// prior Jacobian export: J0 column blocks -> per block n x width row-major, concatenated
__global__ void k_export_prior(Dev D, int PW, double *__restrict__ J_out, const long long *__restrict__ out_off) {
  const int w = blockIdx.x;
  const int n = D.prior_off[w + 1] - D.prior_off[w];
  if (n <= 0) return;
  const double *J0 = D.prior_J + D.priorJ_off[w];
  double *J = J_out + out_off[w];
  long long base = 0;
  for (int b = D.pblk_off[w]; b < D.pblk_off[w + 1]; b++) {
    const int kind = D.pblk_kind[b], col = D.pblk_col[b];
    const int ls = (kind == 0 || kind == 2) ? 6 : (kind == 1 ? 9 : 1);
    const int width = (kind == 0 || kind == 2) ? PW : ls;
    for (int e = threadIdx.x; e < n * width; e += blockDim.x) {
      const int i = e / width, c = e - i * width;
      J[base + e] = c < ls ? J0[(size_t)i * n + col + c] : 0.0;
    }
    base += (long long)n * width;
  }
}

// current iterate (buffer cur[w]) -> flat arrays laid out like the upload.  In the factor-parallel
// multi-GPU mode each rank contributes only what it owns (the caller sums over ranks).
__global__ void k_gather_state(Dev D, double *__restrict__ pose, double *__restrict__ sb, double *__restrict__ ex,
                               double *__restrict__ td, double *__restrict__ inv, double *__restrict__ ortho) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool lead = D.nranks <= 1 || D.rank == 0;
  if (i < D.nF && lead) {
    const int w = find_window(D.frame_off, D.B, i), c = D.cur[w];
    for (int k = 0; k < 7; k++) pose[7 * (size_t)i + k] = D.pose[c][7 * (size_t)i + k];
    for (int k = 0; k < 9; k++) sb[9 * (size_t)i + k] = D.sb[c][9 * (size_t)i + k];
  }
  if (i < D.B && lead) {
    const int c = D.cur[i];
    for (int k = 0; k < 7; k++) ex[7 * (size_t)i + k] = D.ex[c][7 * (size_t)i + k];
    td[i] = D.td[c][i];
  }
  if (i < D.nP && (D.nranks <= 1 || (i % D.nranks) == D.rank)) inv[i] = D.inv_depth[D.cur[D.pt_win[i]]][i];
  if (i < D.nL && (D.nranks <= 1 || (i % D.nranks) == D.rank)) {
    const int c = D.cur[D.ln_win[i]];
    for (int k = 0; k < 4; k++) ortho[4 * (size_t)i + k] = D.ortho[c][4 * (size_t)i + k];
  }
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int launch_gather_state(const Dev &D, double *out_pose, double *out_sb, double *out_ex, double *out_td, double *out_inv,
                        double *out_ortho, cudaStream_t st) {
  const int m = std::max(std::max(D.nF, D.B), std::max(D.nP, D.nL));
  k_gather_state<<<cdiv(m, 256), 256, 0, st>>>(D, out_pose, out_sb, out_ex, out_td, out_inv, out_ortho);
  return 1;
}

int launch_prep(const Dev &D, cudaStream_t st) {
  int n = 0;
  if (D.nProj) { k_prep_proj<<<cdiv(D.nProj, 256), 256, 0, st>>>(D); n++; }
  if (D.nLobs) { k_prep_line<<<cdiv(D.nLobs, 256), 256, 0, st>>>(D); n++; }
  if (D.nVobs) { k_prep_vp<<<cdiv(D.nVobs, 256), 256, 0, st>>>(D); n++; }
  const int m = D.nP > D.nL ? D.nP : D.nL;
  if (m) { k_prep_fix_ranges<<<cdiv(m, 256), 256, 0, st>>>(D); n++; }
  if (D.nImu) { k_imu_info<<<cdiv(D.nImu, II_WPC), 32 * II_WPC, 0, st>>>(D); n++; }
  if (D.nPriorR) { k_prior_H<<<dim3(D.B, D.B >= 74 ? 1 : 8), 256, 0, st>>>(D); n++; }
  return n;
}

int launch_split(const double *rec, long long n, int REC, int NR, double *r_out, double *J_out, cudaStream_t st) {
  if (n == 0) return 0;
  long long blocks = (n * REC + 255) / 256;
  if (blocks > 65535) blocks = 65535;
  k_split<<<(int)blocks, 256, 0, st>>>(rec, n, REC, NR, r_out, J_out);
  return 1;
}

int launch_export_imu(const double *rec, int n, int PW, double *r_out, double *J_out, cudaStream_t st) {
  if (n == 0) return 0;
  k_export_imu<<<n, 128, 0, st>>>(rec, n, PW, r_out, J_out);
  return 1;
}

int launch_export_prior(const Dev &D, int PW, double *J_out, const long long *out_off, cudaStream_t st) {
  if (D.nPriorR == 0) return 0;
  k_export_prior<<<D.B, 128, 0, st>>>(D, PW, J_out, out_off);
  return 1;
}

}  // namespace uvs
