// uvs_build2.cu — atomics-free normal equations + Schur complement for windows of up to 12
// six-wide camera blocks (11 poses + extrinsic): the production path for the reference's
// WINDOW_SIZE = 10 (vins_estimator/src/parameters.h:12).  Larger windows use uvs_build.cu.
//
// Same algebra as uvs_build.cu (Ceres SPARSE_SCHUR restated: block J^T J, landmark elimination,
// back-substitution), different data movement:
//   * one CTA per (window, slice); every warp walks landmarks of the slice;
//   * a landmark's records are one contiguous chunk: the warp stages it in shared memory with
//     coalesced 16-byte loads;
//   * all 32 lanes share the landmark's  H_cc - W (E + D^2)^-1 W^T  entries (6x6 blocks x block
//     pairs), and add them into a WARP-PRIVATE copy of the pose-pose system held in shared memory
//     (upper block triangle, 36 doubles per block pair).  Lanes of one pass hit distinct entries,
//     so plain read-modify-write suffices: no atomics, and the summation order is deterministic;
//   * the CTA sums the private copies and writes the window's reduced system once.
#include "uvs_device.cuh"
#include "uvs_kernels.h"

namespace uvs {

constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ double wsum2(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
  return v;
}
__device__ __forceinline__ double clamp3(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }
__device__ __forceinline__ void atomic_max_nn(double *addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}
// t -> (x <= y) with t = y(y+1)/2 + x
__device__ __forceinline__ void unrank_pair(int t, int &x, int &y) {
  int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (i * (i + 1) / 2 > t) i--;
  while ((i + 1) * (i + 2) / 2 <= t) i++;
  y = i; x = t - i * (i + 1) / 2;
}

struct Build2Smem {
  int nb;          // six-wide camera blocks of this window: F poses (+1 extrinsic)
  int npairs;      // nb (nb + 1) / 2
  int vstride;     // doubles per private V
  int gstride;     // doubles per private gradient / colsq array  (6 nb)
  int sstride;     // doubles of per-warp scratch
};

__host__ __device__ inline int build2_scratch_doubles(int max_frames) {
  // points: (F-1) records x 42 + w (6 (F+1)); lines: F x 33 staged rows + F x 24 Y
  const int pts = 42 * (max_frames - 1) + 6 * (max_frames + 1) + 8;
  const int lns = 33 * max_frames + 24 * max_frames + 40;
  return ((pts > lns ? pts : lns) + 1) & ~1;
}

__device__ __forceinline__ void add_block_entry(double *Vp, int A, int Bk, int p, int q, double v) {
  if (A <= Bk) Vp[(Bk * (Bk + 1) / 2 + A) * 36 + p * 6 + q] += v;
  else Vp[(A * (A + 1) / 2 + Bk) * 36 + q * 6 + p] += v;
}

// ------------------------------------------------------------------------------------------------
template <bool kBack>
__device__ void point_warp(const Dev &D, const Params &P, int w, int gp, int lane, double *Vp, double *gp_s, double *gf_s,
                           double *cs_s, double *scr, int F, bool ex, double radius, bool have_scale, const double *delta) {
  const int REC = D.estimate_td ? REC_PROJ_TD : REC_PROJ;   // td windows never reach this path (REC = 40)
  const int f0 = D.pt_begin[gp], n = D.pt_end[gp] - f0;
  if (n <= 0) {
    if (kBack && lane == 0) { const int cur = D.cur[w]; D.delta_pt[gp] = 0.0; D.inv_depth[cur ^ 1][gp] = D.inv_depth[cur][gp]; }
    return;
  }
  // stage the records (contiguous n x REC doubles) with 16-byte loads
  {
    const double2 *src = reinterpret_cast<const double2 *>(D.rec_proj + (size_t)f0 * REC);
    double2 *dst = reinterpret_cast<double2 *>(scr);
    for (int e = lane; e < n * REC / 2; e += 32) dst[e] = src[e];
  }
  int myblk = 0;   // lane f < n: frame block of the observing pose
  const int fo = D.frame_off[w];
  if (lane < n) myblk = D.proj_idx[f0 + lane].y - fo;
  const int anchor = D.proj_idx[f0].x - fo;
  __syncwarp();
  const double *R = scr;
  double *wv = scr + n * REC;   // w vector, 6 per local block
  const int m = n + 1 + (ex ? 1 : 0);
  int *blks = reinterpret_cast<int *>(wv + 6 * m);   // global block index of every local block
  if (lane < n) blks[1 + lane] = myblk;
  if (lane == 0) { blks[0] = anchor; if (ex) blks[n + 1] = F; }
  __syncwarp();
  // landmark scalars
  double jl0 = 0.0, jl1 = 0.0, r0 = 0.0, r1 = 0.0;
  if (lane < n) { jl0 = R[lane * REC + 38]; jl1 = R[lane * REC + 39]; r0 = R[lane * REC]; r1 = R[lane * REC + 1]; }
  const double colsq = wsum2(jl0 * jl0 + jl1 * jl1);
  const double gk = wsum2(jl0 * r0 + jl1 * r1);
  double sk;
  if (!have_scale) { sk = 1.0 / (1.0 + sqrt(colsq)); if (lane == 0) D.scale_pt[gp] = sk; }
  else sk = D.scale_pt[gp];
  const double Et = sk * sk * colsq;
  const double hinv = 1.0 / (Et + clamp3(Et, P.min_lm_diag, P.max_lm_diag) / radius);
  const double gkt = sk * gk;
  // Jacobian element of factor f for local block kind: 0 anchor (Ji), 1 own (Jj), 2 extrinsic (Jex)
  auto J = [&](int f, int kind, int row, int c) -> double { return R[f * REC + 2 + kind * 12 + row * 6 + c]; };

  if (kBack) {
    // u_f = A_f delta_cam (2 rows per factor), t = g~ + s_k sum Jl^T u, delta_k = -s_k hinv t
    double u0 = 0.0, u1 = 0.0;
    if (lane < n) {
      const double *da = delta + 15 * anchor, *dj = delta + 15 * myblk;
#pragma unroll
      for (int c = 0; c < 6; c++) {
        u0 += J(lane, 0, 0, c) * da[c] + J(lane, 1, 0, c) * dj[c];
        u1 += J(lane, 0, 1, c) * da[c] + J(lane, 1, 1, c) * dj[c];
      }
      if (ex) {
        const double *de = delta + 15 * F;
#pragma unroll
        for (int c = 0; c < 6; c++) { u0 += J(lane, 2, 0, c) * de[c]; u1 += J(lane, 2, 1, c) * de[c]; }
      }
    }
    const double t = sk * (gk + wsum2(jl0 * u0 + jl1 * u1));
    const double dk = sk * (-hinv * t);
    const double jd0 = u0 + jl0 * dk, jd1 = u1 + jl1 * dk;
    const double mc = wsum2(lane < n ? jd0 * (r0 + 0.5 * jd0) + jd1 * (r1 + 0.5 * jd1) : 0.0);
    if (lane == 0) {
      const int cur = D.cur[w];
      const double lam = D.inv_depth[cur][gp];
      D.delta_pt[gp] = dk;
      D.inv_depth[cur ^ 1][gp] = lam + dk;
      gp_s[0] += -mc; gp_s[1] += dk * dk; gp_s[2] += lam * lam;   // per-warp partials: model change, |step|^2, |x|^2
    }
    __syncwarp();
    return;
  }

  if (lane == 0) atomic_max_nn(D.acc + (size_t)w * ACC_STRIDE + ACC_GMAX + D.rank, fabs(gk));
  // w vector and gradient / column norms: one lane per (local block, column)
  for (int e = lane; e < 6 * m; e += 32) {
    const int x = e / 6, c = e - 6 * x;
    double wx = 0.0, g = 0.0, q = 0.0;
    if (x >= 1 && x <= n) {
      const int f = x - 1;
      const double a0 = J(f, 1, 0, c), a1 = J(f, 1, 1, c);
      wx = a0 * R[f * REC + 38] + a1 * R[f * REC + 39];
      g = a0 * R[f * REC] + a1 * R[f * REC + 1];
      q = a0 * a0 + a1 * a1;
    } else {
      const int kind = x == 0 ? 0 : 2;
      for (int f = 0; f < n; f++) {
        const double a0 = J(f, kind, 0, c), a1 = J(f, kind, 1, c);
        wx += a0 * R[f * REC + 38] + a1 * R[f * REC + 39];
        g += a0 * R[f * REC] + a1 * R[f * REC + 1];
        q += a0 * a0 + a1 * a1;
      }
    }
    wx *= sk;
    wv[e] = wx;
    const int gb = blks[x] * 6 + c;
    gf_s[gb] += g;
    gp_s[gb] += g - wx * hinv * gkt;
    cs_s[gb] += q;
  }
  __syncwarp();
  // all entries of H_cc - w w^T hinv over the landmark's block pairs, 36 per pair, lanes over entries
  const int npair = m * (m + 1) / 2;
  const int total = npair * 36;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    int x = 0, y = 0;
    const int tt = t < total ? t : total - 1;
    unrank_pair(tt / 36, x, y);
    const int e = tt % 36, p = e / 6, q = e - 6 * p;
    if (t >= total) continue;
    const int A = blks[x], Bk = blks[y];
    double v = 0.0;
    const bool xo = x >= 1 && x <= n, yo = y >= 1 && y <= n;
    if (xo && yo) {
      if (x == y) { const int f = x - 1; v = J(f, 1, 0, p) * J(f, 1, 0, q) + J(f, 1, 1, p) * J(f, 1, 1, q); }
    } else if (xo) {            // own x extrinsic
      const int f = x - 1;
      v = J(f, 1, 0, p) * J(f, 2, 0, q) + J(f, 1, 1, p) * J(f, 2, 1, q);
    } else if (yo) {            // anchor x own
      const int f = y - 1;
      v = J(f, 0, 0, p) * J(f, 1, 0, q) + J(f, 0, 1, p) * J(f, 1, 1, q);
    } else {                    // anchor/extrinsic x anchor/extrinsic: sum over all factors
      const int kx = x == 0 ? 0 : 2, ky = y == 0 ? 0 : 2;
      for (int f = 0; f < n; f++) v += J(f, kx, 0, p) * J(f, ky, 0, q) + J(f, kx, 1, p) * J(f, ky, 1, q);
    }
    v -= wv[6 * x + p] * wv[6 * y + q] * hinv;
    add_block_entry(Vp, A, Bk, p, q, v);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
template <bool kBack>
__device__ void line_warp(const Dev &D, const Params &P, int w, int gl, int lane, double *Vp, double *gp_s, double *gf_s,
                          double *cs_s, double *scr, int F, double radius, bool have_scale, const double *delta) {
  const int f0 = D.ln_begin[gl], n = D.ln_end[gl] - f0;
  if (n <= 0) {
    if (kBack && lane < 4) { const int cur = D.cur[w]; D.delta_ln[4 * (size_t)gl + lane] = 0.0; D.ortho[cur ^ 1][4 * (size_t)gl + lane] = D.ortho[cur][4 * (size_t)gl + lane]; }
    return;
  }
  // stage: per observation 33 doubles  [r(3) | Jp 3x6 | Jl 3x4], row 2 = the VP factor of the observation (or zeros)
  const int fo = D.frame_off[w];
  for (int e = lane; e < n * 22; e += 32) {
    const int f = e / 22, c = e - 22 * f;
    const double v = D.rec_line[(size_t)(f0 + f) * REC_LINE + c];
    // line record [r0 r1 | Jp row0 (6) row1 (6) | Jl row0 (4) row1 (4)]
    int dst;
    if (c < 2) dst = c;
    else if (c < 14) dst = 3 + (c - 2);
    else dst = 21 + (c - 14);
    scr[f * 33 + dst] = v;
  }
  int myblk = 0, vpi = -1;
  if (lane < n) { const int4 ix = D.line_idx4[f0 + lane]; myblk = ix.x - fo; vpi = ix.w; }
  for (int f = 0; f < n; f++) {
    const int vf = __shfl_sync(FULLM, vpi, f);
    if (lane < 11) {
      const double v = vf >= 0 ? D.rec_vp[(size_t)vf * REC_VP + lane] : 0.0;
      // vp record [r | Jp (6) | Jl (4)]
      const int dst = lane == 0 ? 2 : (lane < 7 ? 15 + (lane - 1) : 29 + (lane - 7));
      scr[f * 33 + dst] = v;
    }
  }
  __syncwarp();
  const double *R = scr;
  double *Y = scr + n * 33;        // Y_f = W_f L^-T  (6x4 per observation)
  double *tmp = Y + n * 24;        // E (10) + g (4)
  auto rr = [&](int f, int row) -> double { return R[f * 33 + row]; };
  auto Jp = [&](int f, int row, int c) -> double { return R[f * 33 + 3 + row * 6 + c]; };
  auto Jl = [&](int f, int row, int c) -> double { return R[f * 33 + 21 + row * 4 + c]; };
  if (lane < 14) {
    double acc = 0.0;
    if (lane < 10) {
      int p = 0, rem = lane;
      while (rem >= 4 - p) { rem -= 4 - p; p++; }
      const int q = p + rem;
      for (int f = 0; f < n; f++) acc += Jl(f, 0, p) * Jl(f, 0, q) + Jl(f, 1, p) * Jl(f, 1, q) + Jl(f, 2, p) * Jl(f, 2, q);
    } else {
      const int c = lane - 10;
      for (int f = 0; f < n; f++) acc += Jl(f, 0, c) * rr(f, 0) + Jl(f, 1, c) * rr(f, 1) + Jl(f, 2, c) * rr(f, 2);
    }
    tmp[lane] = acc;
  }
  __syncwarp();
  double E[16], g4[4], s[4];
  {
    int k = 0;
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int q = p; q < 4; q++) { E[4 * p + q] = tmp[k]; E[4 * q + p] = tmp[k]; k++; }
#pragma unroll
    for (int c = 0; c < 4; c++) g4[c] = tmp[10 + c];
  }
  if (!have_scale) {
#pragma unroll
    for (int p = 0; p < 4; p++) s[p] = 1.0 / (1.0 + sqrt(E[5 * p]));
    if (lane < 4) D.scale_ln[4 * (size_t)gl + lane] = s[lane];
  } else {
#pragma unroll
    for (int p = 0; p < 4; p++) s[p] = D.scale_ln[4 * (size_t)gl + p];
  }
  double M[16];
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int q = 0; q < 4; q++) M[4 * p + q] = s[p] * s[q] * E[4 * p + q];
#pragma unroll
  for (int p = 0; p < 4; p++) M[5 * p] += clamp3(M[5 * p], P.min_lm_diag, P.max_lm_diag) / radius;
  double Lc[16], Li[16];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double dj = M[5 * j];
#pragma unroll
    for (int k = 0; k < j; k++) dj -= Lc[4 * j + k] * Lc[4 * j + k];
    if (!(dj > 0.0)) ok = false;
    dj = sqrt(dj);
    Lc[5 * j] = dj;
#pragma unroll
    for (int i = j + 1; i < 4; i++) {
      double sacc = M[4 * i + j];
#pragma unroll
      for (int k = 0; k < j; k++) sacc -= Lc[4 * i + k] * Lc[4 * j + k];
      Lc[4 * i + j] = sacc / dj;
    }
  }
#pragma unroll
  for (int col = 0; col < 4; col++)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < col) { Li[4 * i + col] = 0.0; continue; }
      double sacc = (i == col) ? 1.0 : 0.0;
#pragma unroll
      for (int k = col; k < i; k++) sacc -= Lc[4 * i + k] * Li[4 * k + col];
      Li[4 * i + col] = sacc / Lc[5 * i];
    }
  if (!ok) {
    if (!kBack && lane == 0) atomicAdd(D.acc + (size_t)w * ACC_STRIDE + ACC_FAIL, 1.0);
    return;
  }
  // z = L^-1 (D_s g)
  double z[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    double sacc = 0.0;
#pragma unroll
    for (int k = 0; k <= c; k++) sacc += Li[4 * c + k] * s[k] * g4[k];
    z[c] = sacc;
  }

  if (kBack) {
    double u[3] = {0.0, 0.0, 0.0};
    if (lane < n) {
      const double *dj = delta + 15 * myblk;
#pragma unroll
      for (int c = 0; c < 6; c++) { u[0] += Jp(lane, 0, c) * dj[c]; u[1] += Jp(lane, 1, c) * dj[c]; u[2] += Jp(lane, 2, c) * dj[c]; }
    }
    double t[4], dk[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const double part = lane < n ? Jl(lane, 0, c) * u[0] + Jl(lane, 1, c) * u[1] + Jl(lane, 2, c) * u[2] : 0.0;
      t[c] = s[c] * (g4[c] + wsum2(part));
    }
    // delta_k = -D_s M^-1 t,  M^-1 = L^-T L^-1
    double v1[4];
#pragma unroll
    for (int c = 0; c < 4; c++) { double sacc = 0.0;
#pragma unroll
      for (int k = 0; k <= c; k++) sacc += Li[4 * c + k] * t[k];
      v1[c] = sacc; }
#pragma unroll
    for (int c = 0; c < 4; c++) { double sacc = 0.0;
#pragma unroll
      for (int k = c; k < 4; k++) sacc += Li[4 * k + c] * v1[k];
      dk[c] = -s[c] * sacc; }
    double mc = 0.0;
    if (lane < n) {
#pragma unroll
      for (int row = 0; row < 3; row++) {
        const double jd = u[row] + Jl(lane, row, 0) * dk[0] + Jl(lane, row, 1) * dk[1] + Jl(lane, row, 2) * dk[2] + Jl(lane, row, 3) * dk[3];
        mc += jd * (rr(lane, row) + 0.5 * jd);
      }
    }
    mc = wsum2(mc);
    if (lane == 0) {
      const int cur = D.cur[w];
      double s2 = 0.0, x2 = 0.0;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double x = D.ortho[cur][4 * (size_t)gl + c];
        D.delta_ln[4 * (size_t)gl + c] = dk[c];
        D.ortho[cur ^ 1][4 * (size_t)gl + c] = x + dk[c];
        s2 += dk[c] * dk[c]; x2 += x * x;
      }
      gp_s[0] += -mc; gp_s[1] += s2; gp_s[2] += x2;
    }
    __syncwarp();
    return;
  }

  if (lane == 0) atomic_max_nn(D.acc + (size_t)w * ACC_STRIDE + ACC_GMAX + D.rank,
                               fmax(fmax(fabs(g4[0]), fabs(g4[1])), fmax(fabs(g4[2]), fabs(g4[3]))));
  // Y_f[p][.] = (Jp^T Jl D_s)[p][.] L^-T, gradient and column norms: one lane per (observation, pose column)
  for (int e = lane; e < 6 * n; e += 32) {
    const int f = e / 6, p = e - 6 * f;
    double W4[4];
#pragma unroll
    for (int c = 0; c < 4; c++) W4[c] = s[c] * (Jp(f, 0, p) * Jl(f, 0, c) + Jp(f, 1, p) * Jl(f, 1, c) + Jp(f, 2, p) * Jl(f, 2, c));
    double y4[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k <= c; k++) sacc += W4[k] * Li[4 * c + k];
      y4[c] = sacc;
      Y[f * 24 + p * 4 + c] = sacc;
    }
    const double g = Jp(f, 0, p) * rr(f, 0) + Jp(f, 1, p) * rr(f, 1) + Jp(f, 2, p) * rr(f, 2);
    const int gb = (D.line_idx4[f0 + f].x - fo) * 6 + p;
    gf_s[gb] += g;
    gp_s[gb] += g - (y4[0] * z[0] + y4[1] * z[1] + y4[2] * z[2] + y4[3] * z[3]);
    cs_s[gb] += Jp(f, 0, p) * Jp(f, 0, p) + Jp(f, 1, p) * Jp(f, 1, p) + Jp(f, 2, p) * Jp(f, 2, p);
  }
  __syncwarp();
  const int npair = n * (n + 1) / 2, total = npair * 36;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    const int tt = t < total ? t : total - 1;
    int x, y;
    unrank_pair(tt / 36, x, y);
    const int e = tt % 36, p = e / 6, q = e - 6 * p;
    const int A = __shfl_sync(FULLM, myblk, x), Bk = __shfl_sync(FULLM, myblk, y);
    if (t >= total) continue;
    const double *ya = Y + x * 24 + p * 4, *yb = Y + y * 24 + q * 4;
    double v = -(ya[0] * yb[0] + ya[1] * yb[1] + ya[2] * yb[2] + ya[3] * yb[3]);
    if (x == y) v += Jp(x, 0, p) * Jp(x, 0, q) + Jp(x, 1, p) * Jp(x, 1, q) + Jp(x, 2, p) * Jp(x, 2, q);
    add_block_entry(Vp, A, Bk, p, q, v);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// grid = B x G CTAs (G slices per window), blockDim = 32 x NW
template <bool kBack>
__global__ void k_window_landmarks(Dev D, Params P, int G, int max_frames) {
  extern __shared__ double sm[];
  const int w = blockIdx.x / G, g = blockIdx.x - w * G;
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  const int NW = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *acc = D.acc + (size_t)w * ACC_STRIDE;
  if (kBack && acc[ACC_FAIL] != 0.0) return;
  const int F = D.frame_off[w + 1] - D.frame_off[w];
  const bool ex = (D.win_flags[w] & WF_EXTRINSIC) != 0;
  const int nb = F + (ex ? 1 : 0);
  const int npairs = nb * (nb + 1) / 2;
  const int vstride = kBack ? 0 : npairs * 36, gstride = kBack ? 4 : 6 * nb;
  const int sstride = build2_scratch_doubles(max_frames);
  // layout: [NW x V] [NW x (gS, gfull, colsq)] [NW x scratch]
  double *Vall = sm;
  double *Gall = Vall + (size_t)NW * vstride;
  double *Sall = Gall + (size_t)NW * 3 * gstride;
  double *Vp = Vall + (size_t)warp * vstride;
  double *gp_s = Gall + (size_t)warp * 3 * gstride, *gf_s = gp_s + gstride, *cs_s = gf_s + gstride;
  double *scr = Sall + (size_t)warp * sstride;
  for (int e = lane; e < vstride; e += 32) Vp[e] = 0.0;
  for (int e = lane; e < 3 * gstride; e += 32) gp_s[e] = 0.0;
  __syncwarp();
  const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
  const double radius = D.ctl[w].radius;
  const bool have_scale = D.ctl[w].have_scale != 0;
  const double *delta = D.delta_cam + co;
  const int p0 = D.point_off[w], np = D.point_off[w + 1] - p0;
  const int l0 = D.line_off[w], nl = D.line_off[w + 1] - l0;
  // lines first (heavier), then points, round-robin over the warps of all slices
  for (int u = g * NW + warp; u < np + nl; u += G * NW) {
    if (u < nl) {
      const int gl = l0 + u;
      if (D.nranks > 1 && (gl % D.nranks) != D.rank) continue;
      line_warp<kBack>(D, P, w, gl, lane, Vp, gp_s, gf_s, cs_s, scr, F, radius, have_scale, delta);
    } else {
      const int gp = p0 + (u - nl);
      if (D.nranks > 1 && (gp % D.nranks) != D.rank) continue;
      point_warp<kBack>(D, P, w, gp, lane, Vp, gp_s, gf_s, cs_s, scr, F, ex, radius, have_scale, delta);
    }
  }
  __syncthreads();
  if (kBack) {
    if (threadIdx.x < 3) {
      double sacc = 0.0;
      for (int k = 0; k < NW; k++) sacc += Gall[(size_t)k * 3 * gstride + threadIdx.x];
      atomicAdd(acc + (threadIdx.x == 0 ? ACC_MODEL : (threadIdx.x == 1 ? ACC_STEP2 : ACC_XNORM2)), sacc);
    }
    return;
  }
  // sum the private copies; write (G == 1) or add (G > 1) the window's reduced system
  double *S = D.Smat + D.S_off[w];
  for (int e = threadIdx.x; e < vstride; e += blockDim.x) {
    double sacc = 0.0;
    for (int k = 0; k < NW; k++) sacc += Vall[(size_t)k * vstride + e];
    const int pr = e / 36, r = e - 36 * pr, p = r / 6, q = r - 6 * p;
    int A, Bk;
    unrank_pair(pr, A, Bk);
    if (A == Bk && p > q) continue;   // lower half of a diagonal block (duplicate of its mirror)
    const int row = (A < F ? 15 * A : 15 * F) + p, col = (Bk < F ? 15 * Bk : 15 * F) + q;
    if (G == 1) S[(size_t)row * d + col] += sacc;
    else atomicAdd(S + (size_t)row * d + col, sacc);
  }
  for (int e = threadIdx.x; e < 3 * gstride; e += blockDim.x) {
    const int which = e / gstride, c = e - which * gstride;
    double sacc = 0.0;
    for (int k = 0; k < NW; k++) sacc += Gall[(size_t)k * 3 * gstride + e];
    const int blk = c / 6, cc = c - 6 * blk;
    const int idx = co + (blk < F ? 15 * blk : 15 * F) + cc;
    double *dst = which == 0 ? D.gS : (which == 1 ? D.gfull : D.colsq_cam);
    if (G == 1) dst[idx] += sacc; else atomicAdd(dst + idx, sacc);
  }
}

size_t build2_smem_bytes(int max_frames, bool any_ex, int NW, bool back) {
  const int nb = max_frames + (any_ex ? 1 : 0);
  const size_t v = back ? 0 : (size_t)nb * (nb + 1) / 2 * 36, gs = back ? 4 : 6 * nb;
  return (size_t)NW * (v + 3 * gs + build2_scratch_doubles(max_frames)) * sizeof(double);
}

int launch_build2(const Dev &D, const Params &P, int G, int NW, int max_frames, bool any_ex, bool back, cudaStream_t st) {
  const size_t smem = build2_smem_bytes(max_frames, any_ex, NW, back);
  static size_t raised[2] = {0, 0};
  if (smem > raised[back ? 1 : 0]) {
    if (back) cudaFuncSetAttribute(k_window_landmarks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(k_window_landmarks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    raised[back ? 1 : 0] = smem;
  }
  if (back) k_window_landmarks<true><<<D.B * G, 32 * NW, smem, st>>>(D, P, G, max_frames);
  else k_window_landmarks<false><<<D.B * G, 32 * NW, smem, st>>>(D, P, G, max_frames);
  return 1;
}

}  // namespace uvs
