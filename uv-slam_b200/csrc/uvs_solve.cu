// uvs_solve.cu — reduced-camera-system solve and Levenberg-Marquardt step control, all on device.
//
// Restates what ceres::Solve does with the options of Estimator::optimization()
// (vins_estimator/src/estimator.cpp:982-994: SPARSE_SCHUR, LEVENBERG_MARQUARDT, defaults otherwise)
// after the landmark blocks have been eliminated (uvs_build.cu):
//   k_solve_init  trust-region state of every window
//   k_chol        Jacobi scaling (fixed from the first Jacobian), LM diagonal, dense Cholesky of the
//                 reduced camera system, camera step, x+ = Plus(x, delta) for poses / speed-biases /
//                 extrinsic / td (pose_local_parameterization.cpp:3-19), model cost change of the
//                 camera-only factors
//   k_chol_chain  the same for windows whose speed-bias blocks form a chain (the reference's shape): the 9-column blocks
//                 are eliminated one by one, a 66-column dense system remains; four windows per SM
//   k_step        step quality, accept / reject, radius update, termination tests, iteration log
// One CTA per window; windows of a batch advance in lock-step but accept / reject independently.
#include <algorithm>
#include <cstdlib>
#include <cstdio>

#include "uvs_device.cuh"
#include "uvs_math.cuh"
#include "uvs_kernels.h"

namespace uvs {

#ifndef UVS_CHOL_THREADS
#define UVS_CHOL_THREADS 512
#endif
constexpr int CT = UVS_CHOL_THREADS;

__device__ __forceinline__ double clampd2(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// clears the window's reduced system: only the upper triangle of S is ever written or read (k_window_system, k_direct*,
// k_window_tail add into it, k_chol reads it), so only that half is cleared - a warp per row, from the diagonal on
__device__ __forceinline__ void zero_window_system(const Dev &D, int w) {
  const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
  double *S = D.Smat + D.S_off[w];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < d; r += nw) {
    double *row = S + (size_t)r * d;
    for (int c = (r & ~3) + lane; c < d; c += 32) row[c] = 0.0;   // from the 32-byte sector that holds the diagonal
  }
  for (int e = threadIdx.x; e < d; e += blockDim.x) { D.gS[co + e] = 0.0; D.gfull[co + e] = 0.0; D.colsq_cam[co + e] = 0.0; }
}

__global__ void __launch_bounds__(CT) k_solve_init(Dev D, Params P) {
  const int w = blockIdx.x;
  if (threadIdx.x == 0) {
    WinCtl &c = D.ctl[w];
    c.cost = 0.0; c.radius = P.initial_radius; c.decrease_factor = 2.0; c.x_norm = 0.0;
    c.state = WS_ACTIVE | WS_NEED_JAC | WS_STEP_OK;
    c.iter = 0; c.n_success = 0; c.n_invalid = 0; c.termination = UVS_TERM_NO_CONVERGENCE; c.status = UVS_OK;
    c.have_scale = 0;
  }
  for (int e = threadIdx.x; e < ACC_STRIDE; e += blockDim.x) D.acc[(size_t)w * ACC_STRIDE + e] = 0.0;
  unsigned char *sm = reinterpret_cast<unsigned char *>(D.summary + w);
  for (int e = threadIdx.x; e < (int)sizeof(UvsSummary); e += blockDim.x) sm[e] = 0;
  zero_window_system(D, w);
}

__device__ __forceinline__ double block_max(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = red[0];
  for (int k = 1; k < (int)blockDim.x / 32; k++) m = fmax(m, red[k]);
  __syncthreads();
  return m;
}
__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = 0.0;
  for (int k = 0; k < (int)blockDim.x / 32; k++) m += red[k];
  __syncthreads();
  return m;
}

constexpr int NB = 8;   // block size of the shared-memory Cholesky
constexpr int MAX_PRIOR_COLS = 512;   // prior dimension bound (15 x 32 frames + extrinsic + td = 487)

// shared-memory layout of the blocked Cholesky (see chol_window): offsets in doubles
struct CholLayout { int K, vr, dbase, vbase, total; };
__host__ __device__ __forceinline__ CholLayout chol_layout(int d) {
  CholLayout L;
  L.K = (d + NB - 1) / NB;
  L.vr = d - NB * (L.K - 1);                                   // rows of the last block row
  L.dbase = L.K > 1 ? 32 * (L.K - 1) * (L.K - 2) + 2 * (L.K - 1) * 4 * L.vr : 0;   // after the off-diagonal fragments
  L.vbase = L.dbase + 36 * L.K;                                // after the packed diagonal blocks
  L.total = L.vbase + 2 * L.K * NB;
  return L;
}

// FP64 tensor-core MMA  D(8x8) += A(8x4) B(4x8)  (mma.sync m8n8k4: lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) + {0,1}])
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// linear index of the row-major lower triangle -> (I, J), J <= I
__device__ __forceinline__ void unrank_lower(int t, int &I, int &J) {
  int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (i * (i + 1) / 2 > t) i--;
  while ((i + 1) * (i + 2) / 2 <= t) i++;
  I = i; J = t - i * (i + 1) / 2;
}

// Blocked right-looking Cholesky, forward and back substitution of a system that already sits in shared memory in
// tensor-core fragment order (layout Lo at `A`; rhs in bz, solution returned in bz).  Called by every thread of the CTA.
// Returns false when a pivot is not positive (s_flag set).
__device__ bool blocked_chol_solve(double *A, const CholLayout &Lo, int nthr, unsigned short *s_pair, int &s_flag) {
  const int tid = threadIdx.x;
  const int K = Lo.K;
  double *Dg = A + Lo.dbase;                          // diagonal blocks
  double *bz = A + Lo.vbase;                          // rhs / z / y   [8K]
  double *invd_all = bz + (size_t)K * NB;             // reciprocal diagonal of L, all rows [8K]
  const int d = NB * (K - 1) + Lo.vr;
  const int warp = tid >> 5, lane = tid & 31;
  auto rows_of = [&](int I) { return I == K - 1 ? Lo.vr : NB; };
  auto frag = [&](int I, int f) { return A + 32 * I * (I - 1) + f * 4 * rows_of(I); };   // fragment f of block row I
    for (int t = tid; t < K * (K + 1) / 2; t += nthr) { int I, J; unrank_lower(t, I, J); s_pair[t] = (unsigned short)(I << 8 | J); }
    __syncthreads();
    // factor of a diagonal block with one warp, in registers: lane r < 8 owns row r; the pivot chain is
    // rsqrt -> scale -> rank-1 update, operands exchanged with shuffles.  Missing rows act as identity rows.
    auto potrf_block = [&](int k) {
      double *invd = invd_all + k * NB;
      double a[NB];
      const int r = lane & 7;
      const bool have = r < rows_of(k);
      double *row = Dg + 36 * k + r * (r + 1) / 2;
#pragma unroll
      for (int c = 0; c < NB; c++) a[c] = have ? (c <= r ? row[c] : 0.0) : (c == r ? 1.0 : 0.0);
      int bad = 0;
#pragma unroll
      for (int p = 0; p < NB; p++) {
        const double piv = __shfl_sync(0xffffffffu, a[p], p);
        if (!(piv > 0.0) || !isfinite(piv)) bad = 1;
        const double inv = rsqrt(piv);
        a[p] = r == p ? piv * inv : a[p] * inv;          // rows r < p are finished (their a[p] is unused)
        if (r == p && lane < NB) invd[p] = inv;
#pragma unroll
        for (int q = p + 1; q < NB; q++) {
          const double lq = __shfl_sync(0xffffffffu, a[p], q);
          a[q] -= a[p] * lq;                             // only meaningful for r >= q
        }
      }
      __syncwarp();   // lanes 8..31 mirror rows 0..7 (r = lane & 7): their reads above come before the owners' write-back
      if (lane < NB && have) {
#pragma unroll
        for (int c = 0; c < NB; c++) if (c <= r) row[c] = a[c];
      }
      if (lane == 0 && bad) s_flag = 1;
    };
    // C -= L_I L_J^T for one 8x8 block on the tensor cores: lane l holds L[l>>2][l&3] of both operands (the B operand
    // of m8n8k4 is column-major, i.e. L_J itself), and C[l>>2][2(l&3) + {0,1}]
    const int fr = lane >> 2, fc = lane & 3;
    auto block_update = [&](int I, int J, int k) {
      const int ra = min(fr, rows_of(I) - 1), rb = min(fr, rows_of(J) - 1);   // rows that do not exist: any finite value
      const double a0 = -frag(I, 2 * k)[4 * ra + fc], a1 = -frag(I, 2 * k + 1)[4 * ra + fc];
      const double b0 = frag(J, 2 * k)[4 * rb + fc], b1 = frag(J, 2 * k + 1)[4 * rb + fc];
      double c[2];
      if (J < I) {
        double2 *C = reinterpret_cast<double2 *>(frag(I, 2 * J + (fc >> 1)) + 4 * ra + 2 * (fc & 1));
        const double2 c2 = fr < rows_of(I) ? *C : make_double2(0.0, 0.0);   // rows that do not exist: their lanes own nothing
        c[0] = c2.x; c[1] = c2.y;
        dmma884(c, a0, b0); dmma884(c, a1, b1);
        if (fr < rows_of(I)) *C = make_double2(c[0], c[1]);
      } else {
        double *C = Dg + 36 * I + fr * (fr + 1) / 2 + 2 * fc;   // row fr, columns 2 fc, 2 fc + 1 (lower part only)
        const bool h0 = 2 * fc <= fr && fr < rows_of(I), h1 = 2 * fc + 1 <= fr && fr < rows_of(I);
        c[0] = h0 ? C[0] : 0.0; c[1] = h1 ? C[1] : 0.0;
        dmma884(c, a0, b0); dmma884(c, a1, b1);
        if (h0) C[0] = c[0];
        if (h1) C[1] = c[1];
      }
    };
    if (warp == 0) potrf_block(0);
    __syncthreads();
    for (int k = 0; k < K && !s_flag; k++) {
      const double *invd = invd_all + k * NB;
      const int vr = rows_of(k);                         // rows of block k that exist
      // panel: L_Ik = A_Ik L_kk^-T for the rows below, z_k = L_kk^-1 b_k
      const int nrows = max(0, d - NB * (k + 1));
      for (int t = tid; t <= nrows; t += nthr) {
        double x[NB];
        double2 *r0, *r1;   // columns 8k..8k+3 and 8k+4..8k+7 of the row
        if (t < nrows) {
          const int i = NB * (k + 1) + t, I = i >> 3, r = i & 7;
          r0 = reinterpret_cast<double2 *>(frag(I, 2 * k) + 4 * r); r1 = reinterpret_cast<double2 *>(frag(I, 2 * k + 1) + 4 * r);
        } else {
          r0 = reinterpret_cast<double2 *>(bz + k * NB); r1 = r0 + 2;
        }
        { const double2 t0 = r0[0], t1 = r0[1], t2 = r1[0], t3 = r1[1];
          x[0] = t0.x; x[1] = t0.y; x[2] = t1.x; x[3] = t1.y; x[4] = t2.x; x[5] = t2.y; x[6] = t3.x; x[7] = t3.y; }
#pragma unroll
        for (int p = 0; p < NB; p++) {
          if (p < vr) {
            const double *akk = Dg + 36 * k + p * (p + 1) / 2;   // row p of L_kk (broadcast reads)
            double v = x[p];
#pragma unroll
            for (int q = 0; q < p; q++) v -= x[q] * akk[q];
            x[p] = v * invd[p];
          }
        }
        r0[0] = make_double2(x[0], x[1]); r0[1] = make_double2(x[2], x[3]);
        r1[0] = make_double2(x[4], x[5]); r1[1] = make_double2(x[6], x[7]);
      }
      __syncthreads();
      // trailing update A_IJ -= L_Ik L_Jk^T, one 8x8 block per warp and step (two DMMAs), b_I -= L_Ik z_k.
      // Look-ahead: warp 0 updates the next diagonal block first and factors it while the other warps do the rest.
      const int nt = K - 1 - k;
      if (warp == 0) {
        if (nt > 0) {
          block_update(k + 1, k + 1, k);
          __syncwarp();
          potrf_block(k + 1);
        }
      } else {
        if (warp == 1) {
          // L_kk is not read again before the back-substitution: replace it by its inverse (lane c: column c),
          // which turns the 8-step dependent chain per block of the back-substitution into independent dot products
          double *Dk = Dg + 36 * k;
          double m[NB];
          const int c = lane & 7;
#pragma unroll
          for (int r = 0; r < NB; r++) {
            double sacc = 0.0;
#pragma unroll
            for (int q = 0; q < r; q++) if (r < vr) sacc += Dk[r * (r + 1) / 2 + q] * m[q];
            m[r] = r < c ? 0.0 : (r == c ? invd[r] : -sacc * invd[r]);
          }
          __syncwarp();
          if (lane < NB) {
#pragma unroll
            for (int r = 0; r < NB; r++) if (r >= c && r < vr) Dk[r * (r + 1) / 2 + c] = m[r];
          }
        }
        const int nblk = nt * (nt + 1) / 2;
        for (int blk = warp; blk < nblk; blk += nthr / 32 - 1) {   // blk 0 = (k+1, k+1): warp 0
          const int pr = s_pair[blk];
          block_update(k + 1 + (pr >> 8), k + 1 + (pr & 255), k);
        }
        for (int t = tid - 32; t < nrows; t += nthr - 32) {
          const int i = NB * (k + 1) + t, I = i >> 3, r = i & 7;
          const double2 *r0 = reinterpret_cast<const double2 *>(frag(I, 2 * k) + 4 * r), *r1 = reinterpret_cast<const double2 *>(frag(I, 2 * k + 1) + 4 * r);
          const double2 *z = reinterpret_cast<const double2 *>(bz + k * NB);
          double v = bz[i];
          v -= r0[0].x * z[0].x; v -= r0[0].y * z[0].y; v -= r0[1].x * z[1].x; v -= r0[1].y * z[1].y;
          v -= r1[0].x * z[2].x; v -= r1[0].y * z[2].y; v -= r1[1].x * z[3].x; v -= r1[1].y * z[3].y;
          bz[i] = v;
        }
      }
      __syncthreads();
    }
    if (s_flag) return false;
    // back-substitution L^T y = z by one warp, block by block from the bottom, "left-looking": for block k first
    // s = sum_{I>k} L_Ik^T y_I on the tensor cores ((y_I^T in row 0 of A) x (block of L as B), four independent
    // accumulator chains, nothing written back in between), then y_k = L_kk^-T (z_k - s) as eight independent dot
    // products (the diagonal blocks hold their inverses by now).
    if (warp == 0) {
      for (int k = K - 1; k >= 0; k--) {
        const int vr = rows_of(k);
        double c[4][2];
#pragma unroll
        for (int u = 0; u < 4; u++) c[u][0] = c[u][1] = 0.0;
        // one warp issues this whole chain: keep the instruction count per block low (pointer increments instead of
        // address arithmetic; the ragged last block row is peeled off)
        const double ymask = fr == 0 ? 1.0 : 0.0;
        {
          const int I = K - 1;   // last block row: rI rows
          if (I > k) {
            const int rI = rows_of(I);
            const double ya = fc < rI ? ymask * bz[I * NB + fc] : 0.0;
            const double yb = 4 + fc < rI ? ymask * bz[I * NB + 4 + fc] : 0.0;
            const double *f = frag(I, 2 * k + (fr >> 2)) + (fr & 3);
            dmma884(c[3], ya, f[4 * min(fc, rI - 1)]);
            dmma884(c[3], yb, f[4 * min(4 + fc, rI - 1)]);
          }
        }
        {
          // full block rows I = k+1 .. K-2: fragment (I, 2k + (fr >> 2)) sits at 32 I (I-1) + 32 (2k + (fr >> 2)), i.e. 64 I further per row
          int I = k + 1;
          const double *f = A + 32 * I * (I - 1) + 32 * (2 * k + (fr >> 2)) + (fr & 3) + 4 * fc;
          const double *yv = bz + I * NB + fc;
          for (; I + 4 <= K - 1; I += 4) {
            const double *f1 = f + 64 * I, *f2 = f1 + 64 * (I + 1), *f3 = f2 + 64 * (I + 2);
            // issue order: the two DMMAs of an accumulator are four apart, so none waits for its predecessor
            dmma884(c[0], ymask * yv[0], f[0]);    dmma884(c[1], ymask * yv[8], f1[0]);
            dmma884(c[2], ymask * yv[16], f2[0]);  dmma884(c[3], ymask * yv[24], f3[0]);
            dmma884(c[0], ymask * yv[4], f[16]);   dmma884(c[1], ymask * yv[12], f1[16]);
            dmma884(c[2], ymask * yv[20], f2[16]); dmma884(c[3], ymask * yv[28], f3[16]);
            f = f3 + 64 * (I + 3); yv += 32;
          }
          for (; I < K - 1; I++) {
            dmma884(c[0], ymask * yv[0], f[0]); dmma884(c[0], ymask * yv[4], f[16]);
            f += 64 * I; yv += NB;
          }
        }
        if (lane < 4) {   // row 0 of the accumulators: columns 2 lane, 2 lane + 1 of block k
          double2 *zc = reinterpret_cast<double2 *>(bz + k * NB + 2 * lane);
          double2 zz = *zc;
          zz.x -= (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
          zz.y -= (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
          *zc = zz;
        }
        __syncwarp();
        const double *Dk = Dg + 36 * k;
        const int p = lane & 7;
        double y = 0.0;
#pragma unroll
        for (int q = 0; q < NB; q++) if (q >= p && q < vr) y += Dk[q * (q + 1) / 2 + p] * bz[k * NB + q];
        __syncwarp();
        if (lane < NB) bz[k * NB + lane] = y;
        __syncwarp();
      }
    }
    __syncthreads();
    return true;
}


// ------------------------------------------------------------------------------------------------
// Structure-exploiting reduced solve ("chain" mode).  The speed-bias blocks B_f (9 columns each) of the reduced camera
// system couple only to their chain neighbours B_f-1, B_f+1 (IMU factors) and to the dense set D = {poses, extrinsic}
// (IMU factors, prior); they never see each other across more than one frame unless a prior says so (checked at
// upload).  Eliminating B_F-1, ..., B_0 in this order therefore needs, per block, a 9x9 Cholesky, the coupling blocks to
// B_f-1 (9x9) and to D (nd x 9), and a rank-9 update of the dense nd x nd part - after which D is a 66..79-column dense
// system for the blocked tensor-core Cholesky above.  Same arithmetic as a Cholesky of the whole matrix in that order,
// but 35..50 KB of shared memory per window instead of 112 KB (four windows per SM instead of two) and ~2.3x fewer
// FLOPs.  L_w (nd x 9 per block) goes to a global scratch for the back-substitution; everything else stays on chip.
constexpr int CH_LWG = 720;      // doubles of global scratch per chain block (L_w, nd x 9)
constexpr int LLS = 190;         // shared memory per chain block: L_c 9 x 10 (reciprocal diagonal) | L_x 9 x 10 | z_f 10
constexpr int CH_ROWS = 96;      // row threads (warps 1..3): one per row of the dense part, nd <= 96

struct ChainLayout { int o_LL, o_z, o_t, o_LwF, o_sc, o_lm, o_g, total; };
__host__ __device__ __forceinline__ ChainLayout chain_layout(int nd, int F) {
  const CholLayout Lo = chol_layout(nd);
  ChainLayout c;
  int o = (Lo.total + 1) & ~1;
  c.o_LL = o; o += LLS * F;                   // per block: C_f -> L_c | X_f -> L_x | b_f -> z_f, factored in place
  c.o_z = o; o += 10 * F;                     // z_f, later y_f
  c.o_t = o; o += 10 * F;
  c.o_LwF = o; o += 2 * Lo.K * 96;            // L_w of a block as tensor-core A fragments [block row][k-step 0..2][32], two buffers
  c.o_sc = o; o += (nd + 9 * F + 1) & ~1;     // Jacobi scale of every column
  c.o_lm = o; o += (nd + 9 * F + 1) & ~1;     // LM diagonal of every column
  c.o_g = o; o += (nd + 9 * F + 1) & ~1;      // scaled negative gradient (right-hand side) of every column
  c.total = o;
  return c;
}

// C(I,J) -= sum_s a_s b_s^T for one 8x8 block of the fragment-layout matrix (nk k-steps of 4; lane holds the operands)
__device__ __forceinline__ void frag_block_sub(double *A, double *Dg, int K, int vr, int I, int J, int lane, const double *a,
                                               const double *b, int nk) {
  const int fr = lane >> 2, fc = lane & 3;
  const int rows = I == K - 1 ? vr : NB;
  double c[2];
  if (J < I) {
    const int ra = min(fr, rows - 1);
    double2 *C = reinterpret_cast<double2 *>(A + 32 * I * (I - 1) + (2 * J + (fc >> 1)) * 4 * rows + 4 * ra + 2 * (fc & 1));
    const double2 c2 = fr < rows ? *C : make_double2(0.0, 0.0);   // rows that do not exist: their lanes own nothing
    c[0] = c2.x; c[1] = c2.y;
    for (int s = 0; s < nk; s++) dmma884(c, -a[s], b[s]);
    if (fr < rows) *C = make_double2(c[0], c[1]);
  } else {
    double *C = Dg + 36 * I + fr * (fr + 1) / 2 + 2 * fc;
    const bool h0 = 2 * fc <= fr && fr < rows, h1 = 2 * fc + 1 <= fr && fr < rows;
    c[0] = h0 ? C[0] : 0.0; c[1] = h1 ? C[1] : 0.0;
    for (int s = 0; s < nk; s++) dmma884(c, -a[s], b[s]);
    if (h0) C[0] = c[0];
    if (h1) C[1] = c[1];
  }
}

__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// Named barriers of the chain pipeline (bar.arrive / bar.sync: waiting warps sleep in hardware, nothing polls).  A barrier id
// serves every second block, and the producer of a phase only arrives once the consumers have left the previous phase of
// the same id (the back / empty barriers), so at most one phase per id is ever open.
constexpr int NB_READY = 1, NB_BACK = 3, NB_FULL = 5, NB_EMPTY = 7;   // + (block & 1)
__device__ __forceinline__ void nbar_arrive(int id, int count) {   // st.shared ; bar.arrive | bar.sync ; ld.shared is the documented producer / consumer pattern
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

struct ChainLayoutV1 { int o_C[2], o_W[2], o_X[2], o_bb, o_LL, o_z, o_LwF, o_t, o_sc, total; };
__host__ __device__ __forceinline__ ChainLayoutV1 chain_layout_v1(int nd, int F) {
  const CholLayout Lo = chol_layout(nd);
  ChainLayoutV1 c;
  int o = (Lo.total + 1) & ~1;
  c.o_C[0] = o; o += 82; c.o_C[1] = o; o += 82;
  const int wsz = (9 * nd + 1) & ~1;          // a W buffer: nd x 9
  c.o_W[0] = o; o += wsz; c.o_W[1] = o; o += wsz;
  c.o_X[0] = o; o += 82; c.o_X[1] = o; o += 82;
  c.o_bb = o; o += 9 * F + (F & 1);          // rhs of every chain block
  c.o_LL = o; o += 164 * F;                   // per block: L_c (81, lower, reciprocal diagonal) | L_x (81) | 2 pad
  c.o_z = o; o += 10 * F;                     // z_f, later y_f
  c.o_LwF = o; o += Lo.K * 96;                // L_w as tensor-core A fragments: [block row][k-step 0..2][32]
  c.o_t = o; o += 10 * F;
  c.o_sc = o; o += (nd + 9 * F + 1) & ~1;     // Jacobi scale of every column (four windows per SM leave room for this, not for more)
  c.total = o;
  return c;
}


// Barrier-phased variant of the chain solve (every warp takes part in every phase, two CTA barriers per block): slower for a
// single window than the pipeline below, faster when four windows share an SM - its waiting warps sleep at CTA barriers and
// leave the load / store unit to the one warp that runs the pivot chain.  Used for large batches (launch_chol_chain).
__device__ bool chain_solve_v1(double *smem, int nthr, unsigned short *s_pair, int &s_flag, const Params &P, double *Sg, int d, int F,
                            const double *scale, const double *colsq, const double *gS, double radius, double *yout, double *lwg) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
  const int nd = d - 9 * F;
  const CholLayout Lo = chol_layout(nd);
  const ChainLayoutV1 Ch = chain_layout_v1(nd, F);
  const int K = Lo.K;
  double *A = smem, *Dg = smem + Lo.dbase, *bz = smem + Lo.vbase, *invd_all = bz + K * NB;
  double *Cb[2] = {smem + Ch.o_C[0], smem + Ch.o_C[1]}, *Wb[2] = {smem + Ch.o_W[0], smem + Ch.o_W[1]};
  double *Xb[2] = {smem + Ch.o_X[0], smem + Ch.o_X[1]}, *bb = smem + Ch.o_bb, *LL = smem + Ch.o_LL, *zb = smem + Ch.o_z, *LwF = smem + Ch.o_LwF, *tb = smem + Ch.o_t;
  double *sc = smem + Ch.o_sc;   // the Jacobi scale, staged: it is read for every entry that is scaled
  auto sidx = [&](int q) { return q < 6 * F ? 15 * (q / 6) + q % 6 : 15 * F + (q - 6 * F); };   // dense index -> index in S
  auto cb = [&](int f) { return 15 * f + 6; };                                                    // first column of B_f in S
  auto rows_of = [&](int I) { return I == K - 1 ? Lo.vr : NB; };
  auto slot = [&](int i, int j) -> double * {   // element (i, j), j <= i, of the dense part (dense indices)
    const int I = i >> 3, r = i & 7;
    return (j >> 3) == I ? Dg + 36 * I + r * (r + 1) / 2 + (j & 7) : A + 32 * I * (I - 1) + (j >> 2) * 4 * rows_of(I) + 4 * r + (j & 3);
  };
  auto lm = [&](int s) { const double v = sc[s], h = v * v * colsq[s]; return clampd2(h, P.min_lm_diag, P.max_lm_diag) / radius; };
  auto sup = [&](int i, int j) { return i <= j ? Sg + (size_t)i * d + j : Sg + (size_t)j * d + i; };   // S is stored as its upper triangle
  // Copies and scaling of the chain blocks belong to warps 1.. (warp 0 runs the 9x9 factorisation meanwhile): thread
  // lt = tid - 32 owns column cc = lt % 9 of rows q0, q0 + R, ... of W, element (q0, cc) of C (q0 < 9) and of X (9 <= q0 < 18)
  const int lt = tid - 32, R = (nthr - 32) / 9;
  const bool part = lt >= 0 && lt < 9 * R;
  const int cc = part ? lt % 9 : 0, q0 = part ? lt / 9 : 0;
  // issue the copies of chain block fb into buffer `buf`: C_fb (lower), W_fb and, for fb > 0, the coupling X_fb (rows
  // B_fb-1, columns B_fb) - everything the factorisation of block fb needs
  auto load_block = [&](int fb, int buf) {
    if (!part) return;
    const int c0 = cb(fb), col = c0 + cc;
#pragma unroll 4
    for (int q = q0; q < nd; q += R) {
      const int sq = sidx(q);
      cp_async8(Wb[buf] + 9 * q + cc, sq <= col ? Sg + (size_t)sq * d + col : Sg + (size_t)col * d + sq);
    }
    if (q0 < 9) { if (cc <= q0) cp_async8(Cb[buf] + 9 * q0 + cc, Sg + (size_t)col * d + c0 + q0); }
    else if (q0 < 18 && fb > 0) cp_async8(Xb[buf] + 9 * (q0 - 9) + cc, Sg + (size_t)(cb(fb - 1) + q0 - 9) * d + col);
  };
  // every thread scales what it copied - and clears the copied entries of S: the system is consumed here, so the next
  // linearisation finds it zeroed without a clearing pass of its own (k_step keeps that pass for the other solvers)
  auto scale_block = [&](int fb, int buf) {
    if (!part) return;
    const int c0 = cb(fb), col = c0 + cc;
    const double scol = sc[col];
#pragma unroll 4
    for (int q = q0; q < nd; q += R) {
      const int sq = sidx(q);
      Wb[buf][9 * q + cc] *= sc[sq] * scol;
      if (sq <= col) Sg[(size_t)sq * d + col] = 0.0; else Sg[(size_t)col * d + sq] = 0.0;
    }
    if (q0 < 9) {
      if (cc <= q0) {
        double v = Cb[buf][9 * q0 + cc] * sc[c0 + q0] * scol; if (cc == q0) v += lm(col); Cb[buf][9 * q0 + cc] = v;
        Sg[(size_t)col * d + c0 + q0] = 0.0;
      }
    } else if (q0 < 18 && fb > 0) {
      Xb[buf][9 * (q0 - 9) + cc] *= sc[cb(fb - 1) + q0 - 9] * scol;
      Sg[(size_t)(cb(fb - 1) + q0 - 9) * d + col] = 0.0;
    }
  };

  for (int t = tid; t < K * (K + 1) / 2; t += nthr) { int I, J; unrank_lower(t, I, J); s_pair[t] = (unsigned short)(I << 8 | J); }
  // ---- load: dense part (fragment slots), first chain block, right-hand sides
  for (int qj = warp; qj < nd; qj += nw) {
    const int sj = sidx(qj);
    for (int qi = qj + lane; qi < nd; qi += 32) cp_async8(slot(qi, qj), Sg + (size_t)sj * d + sidx(qi));
  }
  load_block(F - 1, (F - 1) & 1);
  for (int e = tid; e < K * 96; e += nthr) LwF[e] = 0.0;
  for (int c = tid; c < d; c += nthr) sc[c] = scale[c];
  for (int c = tid; c < K * NB; c += nthr) { bz[c] = c < nd ? -scale[sidx(c)] * gS[sidx(c)] : 0.0; invd_all[c] = 1.0; }
  for (int e = tid; e < 9 * F; e += nthr) { const int f = e / 9, s = cb(f) + e - 9 * f; bb[e] = -scale[s] * gS[s]; }
  cp_async_wait();
  __syncthreads();
  for (int qj = warp; qj < nd; qj += nw) {
    const int sjj = sidx(qj);
    const double sj = sc[sjj];
    for (int qi = qj + lane; qi < nd; qi += 32) { double *a = slot(qi, qj); *a = sc[sidx(qi)] * sj * *a; Sg[(size_t)sjj * d + sidx(qi)] = 0.0; }
  }
  scale_block(F - 1, (F - 1) & 1);
  __syncthreads();
  for (int q = tid; q < nd; q += nthr) *slot(q, q) += lm(sidx(q));
  __syncthreads();
  // dense part -= L_w L_w^T on the tensor cores (three k-steps of four columns; columns 9..11 are zero), 8x8 blocks
  // first, first + step, ... of the lower triangle for this warp.  The update of block f + 1 runs on warps 1.. while
  // warp 0 factors block f (it only touches the dense part and the fragments written before the last barrier).
  auto dense_update = [&](int first, int step) {
    for (int blk = first; blk < K * (K + 1) / 2; blk += step) {
      const int pr = s_pair[blk], I = pr >> 8, J = pr & 255;
      double a[3], b[3];
#pragma unroll
      for (int s = 0; s < 3; s++) { a[s] = LwF[(I * 3 + s) * 32 + lane]; b[s] = LwF[(J * 3 + s) * 32 + lane]; }
      frag_block_sub(A, Dg, K, Lo.vr, I, J, lane, a, b, 3);
    }
  };
  // ---- chain elimination
  for (int f = F - 1; f >= 0 && !s_flag; f--) {
    const int cur = f & 1, nxt = cur ^ 1;
    double *C = Cb[cur], *W = Wb[cur], *Lc = LL + 164 * f, *Lx = Lc + 81, *z = zb + 10 * f;
    if (f > 0) load_block(f - 1, nxt);   // the next block lands while warp 0 factors this one
    if (warp == 0) {
      // Cholesky of the 9x9 block in registers, lane r < 9 owns row r (same pivot chain as the 8x8 blocks).  The
      // right-hand side b_f (lane 9) and the rows of the coupling block X_f (lanes 10..18) ride along as extra rows of
      // the factorisation: what the pivot loop leaves in them is z_f = L_c^-1 b_f and L_x = X_f L_c^-T - no explicit
      // inverse, no separate pass.
      const int r = lane;
      // one source / destination row per lane, no divergence: C row (entries c <= lane), b_f, X_f row, or nothing
      const double *src = lane < 9 ? C + 9 * lane : (lane == 9 ? bb + 9 * f : Xb[cur] + 9 * (lane - 10));
      const int cnt = lane < 9 ? lane + 1 : ((lane == 9 || (lane < 19 && f > 0)) ? 9 : 0);
      double *dst = lane < 9 ? Lc + 9 * lane : (lane == 9 ? z : Lx + 9 * (lane - 10));
      double a[9];
#pragma unroll
      for (int c = 0; c < 9; c++) a[c] = c < cnt ? src[c] : 0.0;
      int bad = 0;
      double myinv = 1.0;
#pragma unroll
      for (int p = 0; p < 9; p++) {
        const double piv = __shfl_sync(0xffffffffu, a[p], p);
        if (!(piv > 0.0) || !isfinite(piv)) bad = 1;
        const double inv = rsqrt(piv);
        a[p] = r == p ? piv * inv : a[p] * inv;
        if (r == p) myinv = inv;
#pragma unroll
        for (int q = p + 1; q < 9; q++) {
          const double lq = __shfl_sync(0xffffffffu, a[p], q);
          a[q] -= a[p] * lq;
        }
      }
      if (lane < 19) {   // L_c with the reciprocal diagonal | z_f | L_x
#pragma unroll
        for (int c = 0; c < 9; c++) dst[c] = lane < 9 ? (c < r ? a[c] : (c == r ? myinv : 0.0)) : a[c];
      }
      if (lane == 0 && bad) s_flag = 1;
    }
    else {
      if (f < F - 1) dense_update(warp - 1, nw - 1);
      if (f > 0) { cp_async_wait(); scale_block(f - 1, nxt); }
    }
    __syncthreads();
    if (s_flag) break;
    // L_w = W L_c^-T row by row (forward substitution); the same thread updates its row of the next block's W and the dense right-hand side
    for (int q = tid; q < nd; q += nthr) {
      double lw[9];
#pragma unroll
      for (int c = 0; c < 9; c++) {   // forward substitution with L_c (reciprocal diagonal)
        double v = W[9 * q + c];
#pragma unroll
        for (int k = 0; k < 9; k++) if (k < c) v -= lw[k] * Lc[9 * c + k];
        lw[c] = v * Lc[10 * c];
      }
      double *g = lwg + (size_t)f * CH_LWG + 9 * q;
      double bq = bz[q];
#pragma unroll
      for (int c = 0; c < 9; c++) {
        g[c] = lw[c];
        LwF[((q >> 3) * 3 + (c >> 2)) * 32 + 4 * (q & 7) + (c & 3)] = lw[c];
        bq -= lw[c] * z[c];
      }
      bz[q] = bq;
      if (f > 0) {
        double *Wn = Wb[nxt] + 9 * q;
#pragma unroll
        for (int c2 = 0; c2 < 9; c2++) {
          double v = Wn[c2];
#pragma unroll
          for (int k = 0; k < 9; k++) v -= lw[k] * Lx[9 * c2 + k];
          Wn[c2] = v;
        }
      }
    }
    if (f > 0 && tid >= nthr - 96 && tid < nthr - 96 + 81) {   // C_f-1 -= L_x L_x^T (lower), b_f-1 -= L_x z: another warp group
      const int e = tid - (nthr - 96), r = e / 9, c = e - 9 * r;
      if (c <= r) {
        double v = Cb[nxt][e];
#pragma unroll
        for (int k = 0; k < 9; k++) v -= Lx[9 * r + k] * Lx[9 * c + k];
        Cb[nxt][e] = v;
      }
      if (c == 0) {
        double v = bb[9 * (f - 1) + r];
#pragma unroll
        for (int k = 0; k < 9; k++) v -= Lx[9 * r + k] * z[k];
        bb[9 * (f - 1) + r] = v;
      }
    }
    __syncthreads();
  }
  if (s_flag) return false;
  dense_update(warp, nw);   // the update of block 0 has no factorisation to hide behind: all warps
  __syncthreads();

  // ---- dense part
  if (!blocked_chol_solve(A, Lo, nthr, s_pair, s_flag)) return false;

  // ---- back-substitution of the chain: t_f = z_f - L_w^T y_D for all blocks at once, then y_f = L_c^-T (t_f - L_x^T y_f-1)
  for (int base = 0; base < 9 * F; base += nthr / 2) {   // two threads per entry, each half of the dot product
    const int e = base + (tid >> 1), half = tid & 1;
    const int f = e / 9, c = e - 9 * f;
    double t = 0.0;
    if (e < 9 * F) {
      const double *g = lwg + (size_t)f * CH_LWG + c;
      const int qm = (nd + 1) >> 1;
      for (int q = half ? qm : 0; q < (half ? nd : qm); q++) t -= g[9 * q] * bz[q];
    }
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    if (e < 9 * F && half == 0) tb[10 * f + c] = zb[10 * f + c] + t;
  }
  __syncthreads();
  if (warp == 0) {
    const int r = lane < 9 ? lane : 8;
    for (int f = 0; f < F; f++) {
      const double *Lc = LL + 164 * f, *Lx = Lc + 81;
      double u = tb[10 * f + r];
      if (f > 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) u -= Lx[9 * k + r] * zb[10 * (f - 1) + k];
      }
      // y = L_c^-T u by backward substitution, lane r owns y_r: column r of L_c in registers
      double lcol[9];
#pragma unroll
      for (int k = 0; k < 9; k++) lcol[k] = Lc[9 * k + r];
      double y = 0.0;
#pragma unroll
      for (int k = 8; k >= 0; k--) {
        const double yk = __shfl_sync(0xffffffffu, u * lcol[k], k);   // lane k: u_k / L_kk
        if (r == k) y = yk;
        if (r < k) u -= lcol[k] * yk;
      }
      if (lane < 9) zb[10 * f + lane] = y;   // y_f replaces z_f
      __syncwarp();
    }
  }
  __syncthreads();
  for (int q = tid; q < nd; q += nthr) yout[sidx(q)] = bz[q];
  for (int e = tid; e < 9 * F; e += nthr) { const int f = e / 9; yout[cb(f) + e - 9 * f] = zb[10 * f + e - 9 * f]; }
  __syncthreads();
  return true;
}


// Solves (D_s S D_s + D^2) y = -D_s g for one window; y (in the order of S) -> yout (global).  All 256 threads of the CTA.
//
// The chain elimination is a three-stage pipeline of specialised warps that meet at named barriers only (no CTA barrier inside
// the loop over the blocks):
//   warp 0       factors block f: C_f minus the rank-9 term of block f+1, with b_f and the rows of X_f riding through the
//                9-pivot loop as extra rows (-> z_f, L_x); nothing it needs comes from the other warps, so it runs ahead
//   warps 1..3   one thread per row q of the dense part: L_w[q] = W_f[q] L_c^-T (W read straight from S, one block ahead),
//                W_f-1[q] -= L_w[q] L_x^T and the dense right-hand side, all in registers; L_w goes to the fragment buffer
//                of the block (two buffers) and to the global scratch of the back-substitution
//   warps 4..7   dense part -= L_w L_w^T on the FP64 tensor cores, as soon as a block's fragments are complete
__device__ bool chain_solve(double *smem, int nthr, unsigned short *s_pair, int &s_flag, const Params &P, double *Sg, int d, int F,
                            const double *scale, const double *colsq, const double *gS, double radius, double *yout, double *lwg) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
  const int nd = d - 9 * F;
  const CholLayout Lo = chol_layout(nd);
  const ChainLayout Ch = chain_layout(nd, F);
  const int K = Lo.K;
  double *A = smem, *Dg = smem + Lo.dbase, *bz = smem + Lo.vbase, *invd_all = bz + K * NB;
  double *LL = smem + Ch.o_LL, *zb = smem + Ch.o_z, *tb = smem + Ch.o_t, *LwF = smem + Ch.o_LwF, *sc = smem + Ch.o_sc;
  double *lmd = smem + Ch.o_lm, *rhs = smem + Ch.o_g;
  auto sidx = [&](int q) { return q < 6 * F ? 15 * (q / 6) + q % 6 : 15 * F + (q - 6 * F); };   // dense index -> index in S
  auto cb = [&](int f) { return 15 * f + 6; };                                                    // first column of B_f in S
  auto rows_of = [&](int I) { return I == K - 1 ? Lo.vr : NB; };
  auto slot = [&](int i, int j) -> double * {   // element (i, j), j <= i, of the dense part (dense indices)
    const int I = i >> 3, r = i & 7;
    return (j >> 3) == I ? Dg + 36 * I + r * (r + 1) / 2 + (j & 7) : A + 32 * I * (I - 1) + (j >> 2) * 4 * rows_of(I) + 4 * r + (j & 3);
  };
  // LM diagonal and right-hand side of every column are staged with the scale (one round trip for all three vectors):
  // read where they are needed, a global load per diagonal entry stalled its warp once per column
  auto lm = [&](int s) { return lmd[s]; };
  // entry e of the chain blocks' inputs: C_f (lower, 45) and, for f > 0, X_f (rows B_f-1, columns B_f, 81) -> slot in LL, the
  // two S indices (s1 <= s2: S is stored as its upper triangle)
  auto chain_entry = [&](int e, double *&dst, int &s1, int &s2) -> bool {
    const int f = e / 126, k = e - 126 * f;
    double *blk = LL + LLS * f;
    if (k < 45) { int r, c; unrank_lower(k, r, c); dst = blk + 10 * r + c; s1 = cb(f) + c; s2 = cb(f) + r; return true; }
    if (f == 0) return false;
    const int r = (k - 45) / 9, c = (k - 45) - 9 * r;
    dst = blk + 90 + 10 * r + c; s1 = cb(f - 1) + r; s2 = cb(f) + c;
    return true;
  };
#ifdef UVS_CHOL_TIMING
  long long tc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq0 = clock64(), tq1;
#define CH_T(i) do { tq1 = clock64(); tc[i] += tq1 - tq0; tq0 = tq1; } while (0)
  long long rt[6] = {0, 0, 0, 0, 0, 0}, r0 = 0;
#define CH_R0() do { r0 = clock64(); } while (0)
#define CH_R(i) do { const long long r1 = clock64(); rt[i] += r1 - r0; r0 = r1; } while (0)
#else
#define CH_R0() do { } while (0)
#define CH_R(i) do { } while (0)
#define CH_T(i) do { } while (0)
#endif

  // ---- load: everything the solve reads from S goes to shared memory (or, for W, to registers later) in one round trip
  for (int t = tid; t < K * (K + 1) / 2; t += nthr) { int I, J; unrank_lower(t, I, J); s_pair[t] = (unsigned short)(I << 8 | J); }
  for (int qj = warp; qj < nd; qj += nw) {
    const int sj = sidx(qj);
    for (int qi = qj + lane; qi < nd; qi += 32) cp_async8(slot(qi, qj), Sg + (size_t)sj * d + sidx(qi));
  }
  for (int e = tid; e < 126 * F; e += nthr) {
    double *dst; int s1, s2;
    if (chain_entry(e, dst, s1, s2)) cp_async8(dst, Sg + (size_t)s1 * d + s2);
  }
  for (int c = tid; c < d; c += nthr) {
    const double v = scale[c], h = v * v * colsq[c];
    sc[c] = v; lmd[c] = clampd2(h, P.min_lm_diag, P.max_lm_diag) / radius; rhs[c] = -v * gS[c];
  }
  for (int e = tid; e < 2 * K * 96; e += nthr) LwF[e] = 0.0;
  cp_async_wait();
  __syncthreads();
  // scaling, LM diagonal - and the copied entries of S are cleared: the system is consumed here, so the next
  // linearisation finds it zeroed without a clearing pass of its own (k_step keeps that pass for the other solvers)
  for (int qj = warp; qj < nd; qj += nw) {
    const int sjj = sidx(qj);
    const double sj = sc[sjj];
    for (int qi = qj + lane; qi < nd; qi += 32) {
      double *a = slot(qi, qj);
      double v = sc[sidx(qi)] * sj * *a;
      if (qi == qj) v += lm(sjj);
      *a = v;
      Sg[(size_t)sjj * d + sidx(qi)] = 0.0;
    }
  }
  for (int e = tid; e < 126 * F; e += nthr) {
    double *dst; int s1, s2;
    if (chain_entry(e, dst, s1, s2)) {
      double v = *dst * sc[s1] * sc[s2];
      if (s1 == s2) v += lm(s1);
      *dst = v;
      Sg[(size_t)s1 * d + s2] = 0.0;
    }
  }
  for (int e = tid; e < 9 * F; e += nthr) { const int f = e / 9, c = e - 9 * f; LL[LLS * f + 180 + c] = rhs[cb(f) + c]; }
  for (int c = tid; c < K * NB; c += nthr) { bz[c] = 0.0; invd_all[c] = 1.0; }
  // row threads: row q of W_F-1 (scaled) and the dense right-hand side, in registers
  const int q = tid - 32;
  const bool rowthr = q >= 0 && q < CH_ROWS, rowact = rowthr && q < nd;
  const int sq = rowact ? sidx(q) : 0;
  const double ssq = rowact ? sc[sq] : 0.0;
  double wn[9], bzq = 0.0;
  // raw row q of W_fb out of S (upper triangle: the row of S when q's column comes first, else the column), cleared behind
  auto load_w = [&](int fb, double (&dst)[9]) {
    const int c0 = cb(fb);
    if (sq <= c0) {
      double *src = Sg + (size_t)sq * d + c0;
#pragma unroll
      for (int c = 0; c < 9; c++) { dst[c] = src[c]; src[c] = 0.0; }
    } else {
      double *src = Sg + (size_t)c0 * d + sq;
#pragma unroll
      for (int c = 0; c < 9; c++) { dst[c] = src[(size_t)c * d]; src[(size_t)c * d] = 0.0; }
    }
  };
  if (rowact) {
    load_w(F - 1, wn);
#pragma unroll
    for (int c = 0; c < 9; c++) wn[c] *= ssq * sc[cb(F - 1) + c];
    bzq = rhs[sq];
  } else {
#pragma unroll
    for (int c = 0; c < 9; c++) wn[c] = 0.0;
  }
  __syncthreads();
  CH_T(0);

  // ---- chain elimination
  if (warp == 0) {
    for (int f = F - 1; f >= 0; f--) {
      double *Lc = LL + LLS * f, *Lx = Lc + 90;
      // one row per lane, no divergence: C row (entries c <= lane), b_f (lane 9), X_f row (lanes 10..18), or nothing
      double *row = lane < 9 ? Lc + 10 * lane : (lane == 9 ? Lx + 90 : Lx + 10 * (lane - 10));
      const int cnt = lane < 9 ? lane + 1 : ((lane == 9 || (lane < 19 && f > 0)) ? 9 : 0);
      const int r = lane;
      double a[9];
      CH_R0();
#pragma unroll
      for (int c = 0; c < 9; c++) a[c] = c < cnt ? row[c] : 0.0;
      if (f < F - 1) {
        // what the elimination of block f + 1 left for this one: C_f -= L_x L_x^T, b_f -= L_x z (rows 0..8 of U: L_x, row 9: z)
        const double *U = LL + LLS * (f + 1) + 90;
        const double *ur = U + 10 * min(lane, 9);
        double u[9];
#pragma unroll
        for (int k = 0; k < 9; k++) u[k] = lane <= 9 ? ur[k] : 0.0;
#pragma unroll
        for (int c = 0; c < 9; c++) {
          const double *lx = U + 10 * c;
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 9; k++) s += u[k] * lx[k];
          a[c] -= c < cnt ? s : 0.0;
        }
      }
      CH_R(0);
      // Cholesky of the 9x9 block in registers, lane r < 9 owns row r.  The right-hand side and the rows of the coupling block
      // ride along as extra rows: what the pivot loop leaves in them is z_f = L_c^-1 b_f and L_x = X_f L_c^-T
      int bad = 0;
      double myinv = 1.0;
#pragma unroll
      for (int p = 0; p < 9; p++) {
        const double piv = __shfl_sync(0xffffffffu, a[p], p);
        if (!(piv > 0.0) || !isfinite(piv)) bad = 1;
        const double inv = rsqrt(piv);
        a[p] = r == p ? piv * inv : a[p] * inv;
        if (r == p) myinv = inv;
#pragma unroll
        for (int qq = p + 1; qq < 9; qq++) {
          const double lq = __shfl_sync(0xffffffffu, a[p], qq);
          a[qq] -= a[p] * lq;
        }
      }
      CH_R(1);
      if (lane < 19) {   // L_c with the reciprocal diagonal | z_f | L_x, in place
#pragma unroll
        for (int c = 0; c < 9; c++) row[c] = lane < 9 ? (c < r ? a[c] : (c == r ? myinv : 0.0)) : a[c];
      }
      if (lane == 9) {
#pragma unroll
        for (int c = 0; c < 9; c++) zb[10 * f + c] = a[c];
      }
      if (lane == 0 && bad) s_flag = 1;
      CH_R(2);
      if (f <= F - 3) nbar_sync(NB_BACK + (f & 1), 32 + CH_ROWS);   // the row threads have left block f + 2 (same barrier id)
      nbar_arrive(NB_READY + (f & 1), 32 + CH_ROWS);
      CH_R(3);
    }
#ifdef UVS_CHOL_TIMING
    if (lane == 0 && blockIdx.x == 0) printf("  warp 0: update %lld pivots %lld stores %lld barriers %lld\n", rt[0], rt[1], rt[2], rt[3]);
#endif
  } else if (rowthr) {
    for (int f = F - 1; f >= 0; f--) {
      const double *Lc = LL + LLS * f, *Lx = Lc + 90, *z = Lx + 90;
      double nx[9];
      CH_R0();
      if (f > 0 && rowact) load_w(f - 1, nx);   // in flight while this block is processed
      CH_R(0);
      nbar_sync(NB_READY + (f & 1), 32 + CH_ROWS);
      CH_R(1);
      // L_w[q] = W_f[q] L_c^-T by forward substitution, in place
#pragma unroll
      for (int c = 0; c < 9; c++) {
        double v = wn[c];
#pragma unroll
        for (int k = 0; k < 9; k++) if (k < c) v -= wn[k] * Lc[10 * c + k];
        wn[c] = v * Lc[11 * c];
      }
      CH_R(2);
      if (f <= F - 3) nbar_sync(NB_EMPTY + (f & 1), CH_ROWS + 128);   // the fragment buffer of this block was last used two blocks ago
      CH_R(3);
      if (rowact) {
        double *g = lwg + (size_t)f * CH_LWG + 9 * q;
        double *fb = LwF + (f & 1) * K * 96 + (q >> 3) * 96 + 4 * (q & 7);
#pragma unroll
        for (int c = 0; c < 9; c++) {
          g[c] = wn[c];
          fb[(c >> 2) * 32 + (c & 3)] = wn[c];
          bzq -= wn[c] * z[c];
        }
      }
      nbar_arrive(NB_FULL + (f & 1), CH_ROWS + 128);
      CH_R(4);
      if (f > 0) {
        const int c1 = cb(f - 1);
#pragma unroll
        for (int c2 = 0; c2 < 9; c2++) {
          double v = rowact ? nx[c2] * ssq * sc[c1 + c2] : 0.0;
#pragma unroll
          for (int k = 0; k < 9; k++) v -= wn[k] * Lx[10 * c2 + k];
          nx[c2] = v;
        }
#pragma unroll
        for (int c = 0; c < 9; c++) wn[c] = nx[c];
      }
      if (f >= 2) nbar_arrive(NB_BACK + (f & 1), 32 + CH_ROWS);
      CH_R(5);
    }
#ifdef UVS_CHOL_TIMING
    if (q == 0 && blockIdx.x == 0) printf("  rows: issue loads %lld wait ready %lld forward %lld wait empty %lld write+arrive %lld next-W %lld\n", rt[0], rt[1], rt[2], rt[3], rt[4], rt[5]);
#endif
    if (rowact) bz[q] = bzq;
  } else {
    // dense part -= L_w L_w^T (three k-steps of four columns; columns 9..11 are zero), 8x8 blocks u, u + 4, ... of the lower triangle
    const int u = warp - 4;
    for (int f = F - 1; f >= 0; f--) {
      const double *buf = LwF + (f & 1) * K * 96;
      CH_R0();
      nbar_sync(NB_FULL + (f & 1), CH_ROWS + 128);
      CH_R(0);
      for (int blk = u; blk < K * (K + 1) / 2; blk += 4) {
        const int pr = s_pair[blk], I = pr >> 8, J = pr & 255;
        double a[3], b[3];
#pragma unroll
        for (int s = 0; s < 3; s++) { a[s] = buf[(I * 3 + s) * 32 + lane]; b[s] = buf[(J * 3 + s) * 32 + lane]; }
        frag_block_sub(A, Dg, K, Lo.vr, I, J, lane, a, b, 3);
      }
      if (f >= 2) nbar_arrive(NB_EMPTY + (f & 1), CH_ROWS + 128);
      CH_R(1);
    }
#ifdef UVS_CHOL_TIMING
    if (tid == 128 && blockIdx.x == 0) printf("  update: wait full %lld tiles %lld\n", rt[0], rt[1]);
#endif
  }
  __syncthreads();
  CH_T(1);
  if (s_flag) return false;

  // ---- dense part
  if (!blocked_chol_solve(A, Lo, nthr, s_pair, s_flag)) return false;
  CH_T(2);

  // ---- back-substitution of the chain: t_f = z_f - L_w^T y_D for all blocks at once, then y_f = L_c^-T (t_f - L_x^T y_f-1)
  for (int base = 0; base < 9 * F; base += nthr / 2) {   // two threads per entry, each half of the dot product
    const int e = base + (tid >> 1), half = tid & 1;
    const int f = e / 9, c = e - 9 * f;
    double t = 0.0;
    if (e < 9 * F) {
      const double *g = lwg + (size_t)f * CH_LWG + c;
      const int qm = (nd + 1) >> 1;
      for (int qq = half ? qm : 0; qq < (half ? nd : qm); qq++) t -= g[9 * qq] * bz[qq];
    }
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    if (e < 9 * F && half == 0) tb[10 * f + c] = zb[10 * f + c] + t;
  }
  __syncthreads();
  if (warp == 0) {
    const int r = lane < 9 ? lane : 8;
    for (int f = 0; f < F; f++) {
      const double *Lc = LL + LLS * f, *Lx = Lc + 90;
      double u = tb[10 * f + r];
      if (f > 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) u -= Lx[10 * k + r] * zb[10 * (f - 1) + k];
      }
      // y = L_c^-T u by backward substitution, lane r owns y_r: column r of L_c in registers
      double lcol[9];
#pragma unroll
      for (int k = 0; k < 9; k++) lcol[k] = Lc[10 * k + r];
      double y = 0.0;
#pragma unroll
      for (int k = 8; k >= 0; k--) {
        const double yk = __shfl_sync(0xffffffffu, u * lcol[k], k);   // lane k: u_k / L_kk
        if (r == k) y = yk;
        if (r < k) u -= lcol[k] * yk;
      }
      if (lane < 9) zb[10 * f + lane] = y;   // y_f replaces z_f
      __syncwarp();
    }
  }
  __syncthreads();
  for (int qq = tid; qq < nd; qq += nthr) yout[sidx(qq)] = bz[qq];
  for (int e = tid; e < 9 * F; e += nthr) { const int f = e / 9; yout[cb(f) + e - 9 * f] = zb[10 * f + e - 9 * f]; }
  __syncthreads();
  CH_T(3);
#ifdef UVS_CHOL_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("chain nd=%d F=%d cycles: load %lld | chain pipeline %lld | dense solve %lld | chain back-sub %lld\n", nd, F, tc[0], tc[1], tc[2], tc[3]);
#endif
  return true;
}

// developer aid: -DUVS_CHOL_TIMING prints the cycle count of every phase of window 0 (one line per launch)
#ifdef UVS_CHOL_TIMING
#define CHOL_TS(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) ts[i] = clock64(); } while (0)
#else
#define CHOL_TS(i) do { } while (0)
#endif

// kMode: 0 = dense blocked, factor in a per-window global scratch (large windows), 1 = dense blocked in shared memory, 2 = chain mode (barrier-phased),
//        3 = chain mode (warp-specialised pipeline)
template <int kMode>
__device__ void chol_window(const Dev &D, const Params &P, int w, double *smem, bool mc_identity) {
  __shared__ double red[CT / 32];   // CT = largest block size
  __shared__ int s_flag;
  constexpr bool kPacked = kMode == 1;
  __shared__ unsigned short s_pair[kMode == 1 ? 528 : (kMode >= 2 ? 64 : 1)];   // (I, J) of the t-th block of a lower triangle (K <= 32; chain mode K <= 10)
  const int nthr = blockDim.x;   // 256 when two windows fit one SM, else 512
#ifdef UVS_CHOL_TIMING
  long long ts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  CHOL_TS(0);
  int *s_cmap = reinterpret_cast<int *>(smem);   // aliases the factor: only used after the back-substitution
  double *vec_y = D.gS + D.cam_off[w];   // y (scaled step) parked in the consumed reduced-gradient buffer
  WinCtl &ctl = D.ctl[w];
  const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
  const int F = D.frame_off[w + 1] - D.frame_off[w];
  const int fl = D.win_flags[w];
  double *acc = D.acc + (size_t)w * ACC_STRIDE;
  const int tid = threadIdx.x;
  const double *gfull = D.gfull + co, *gS = D.gS + co, *colsq = D.colsq_cam + co;
  double *scale = D.scale_cam + co;
  double *Sg = D.Smat + D.S_off[w];

  // ---- gradient max norm of the current iterate, iteration-0 bookkeeping, gradient tolerance
  double gm = 0.0;
  for (int c = tid; c < d; c += nthr) gm = fmax(gm, fabs(gfull[c]));
  if (tid < MAX_RANKS) gm = fmax(gm, acc[ACC_GMAX + tid]);
  gm = block_max(gm, red);
  if (tid == 0) {
    UvsSummary &sm = D.summary[w];
    if (ctl.iter == 0) {
      ctl.cost = acc[ACC_COST0];
      sm.initial_cost = ctl.cost;
      sm.cost[0] = ctl.cost; sm.radius[0] = ctl.radius; sm.gradient_max_norm[0] = gm; sm.step_accepted[0] = 1;
      ctl.iter = 1;
    } else if ((ctl.state & WS_NEED_JAC) && ctl.iter - 1 < UVS_MAX_ITER_LOG) {
      sm.gradient_max_norm[ctl.iter - 1] = gm;
    }
    int stop = 0;
    if (!P.fixed_iterations && (ctl.state & WS_NEED_JAC) && gm <= P.gradient_tolerance) {
      ctl.termination = UVS_TERM_GRADIENT_TOL; ctl.state &= ~WS_ACTIVE; stop = 1;
    }
    if (!isfinite(ctl.cost)) { ctl.termination = UVS_TERM_FAILURE; ctl.status = UVS_ERR_NOT_FINITE; ctl.state &= ~WS_ACTIVE; stop = 1; }
    s_flag = stop;
  }
  __syncthreads();
  if (s_flag) return;

  // ---- Jacobi scaling, fixed from the first Jacobian: s = 1 / (1 + ||J[:,c]||)
  if (!ctl.have_scale) for (int c = tid; c < d; c += nthr) scale[c] = 1.0 / (1.0 + sqrt(colsq[c]));
  __syncthreads();

  double *vec;   // solution vector y (d doubles)
  const double radius = ctl.radius;
  CHOL_TS(1);
  if (kMode >= 2) {
    vec = vec_y;   // the solution is written straight to its global home
    CHOL_TS(2);
    double *lwg = D.chain_lw + (size_t)w * D.chain_lw_stride;
    const bool ok = kMode == 3 ? chain_solve(smem, nthr, s_pair, s_flag, P, Sg, d, F, scale, colsq, gS, radius, vec_y, lwg)
                               : chain_solve_v1(smem, nthr, s_pair, s_flag, P, Sg, d, F, scale, colsq, gS, radius, vec_y, lwg);
    if (!ok) {
      if (tid == 0) { acc[ACC_FAIL] += 1.0; ctl.state &= ~WS_STEP_OK; ctl.have_scale = 1; }
      zero_window_system(D, w);   // the failed solve consumed (and cleared) only part of the system
      return;
    }
    CHOL_TS(4);
    CHOL_TS(5);
  } else if (kPacked) {
    // ---- blocked path: the lower triangle in shared memory, 8x8 blocks, stored in FP64 tensor-core FRAGMENT order:
    //      block row I (rows 8I..8I+7), columns 0..8I-1, is a sequence of 8x4 fragments (32 consecutive doubles,
    //      element (r, c) at 4r + c = the lane that holds it in mma.sync m8n8k4), so every operand load of the
    //      trailing update is one contiguous, conflict-free 256-byte read; the diagonal blocks live apart as packed
    //      lower triangles (36 doubles); the last block row keeps only the rows that exist.  112 KB at d = 165, i.e.
    //      two windows per SM: while one sits in its pivot chain the other one's trailing update runs.
    const CholLayout Lo = chol_layout(d);
    const int K = Lo.K;
    double *A = smem;
    double *Dg = smem + Lo.dbase;                       // diagonal blocks
    double *bz = smem + Lo.vbase;                       // rhs / z / y   [8K]
    double *invd_all = bz + (size_t)K * NB;             // reciprocal diagonal of L, all rows [8K]
    vec = bz;
    const int warp = tid >> 5, lane = tid & 31;
    auto rows_of = [&](int I) { return I == K - 1 ? Lo.vr : NB; };
    auto frag = [&](int I, int f) { return A + 32 * I * (I - 1) + f * 4 * rows_of(I); };   // fragment f of block row I
    {
      // The upper triangle of Sg goes straight to its fragment slots with asynchronous 8-byte copies (LDGSTS): rows j by
      // warp, columns i >= j by lane (coalesced), all ~27 copies of a thread in flight at once - one memory round trip
      // for the whole matrix instead of one per row.  Scaling and the LM diagonal are applied in place afterwards.
      auto slot = [&](int i, int j) -> double * {   // element (i, j), j <= i, of the lower matrix
        const int I = i >> 3, r = i & 7;
        return (j >> 3) == I ? Dg + 36 * I + r * (r + 1) / 2 + (j & 7) : frag(I, j >> 2) + 4 * r + (j & 3);
      };
      for (int j = warp; j < d; j += nthr / 32) {
        const double *src = Sg + (size_t)j * d;
        for (int i = j + lane; i < d; i += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(slot(i, j))), "l"(src + i) : "memory");
      }
#ifdef UVS_CHOL_TIMING
      const long long tl0 = clock64();
#endif
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
      __syncthreads();
#ifdef UVS_CHOL_TIMING
      const long long tl1 = clock64();
      if (threadIdx.x == 0 && blockIdx.x == 0) printf("  load: issue %lld wait %lld", tl0 - ts[1], tl1 - tl0);
#endif
      for (int j = warp; j < d; j += nthr / 32) {
        const double sj = scale[j];
        for (int i = j + lane; i < d; i += 32) {
          double *a = slot(i, j);
          *a = scale[i] * sj * *a;
        }
      }
      // LM diagonal, one thread per column (kept out of the row loop: a load and a division on one lane would stall
      // the whole warp once per row); each diagonal slot was scaled by this very thread's warp-mates above -> barrier
      __syncthreads();
      for (int c = tid; c < d; c += nthr) {
        const double sc = scale[c], h = sc * sc * colsq[c];
        *slot(c, c) += clampd2(h, P.min_lm_diag, P.max_lm_diag) / radius;
      }
#ifdef UVS_CHOL_TIMING
      if (threadIdx.x == 0 && blockIdx.x == 0) printf(" scale %lld\n", clock64() - tl1);
#endif
    }
    for (int c = tid; c < K * NB; c += nthr) { bz[c] = c < d ? -scale[c] * gS[c] : 0.0; invd_all[c] = 1.0; }
    CHOL_TS(2);
    if (!blocked_chol_solve(A, Lo, nthr, s_pair, s_flag)) {
      if (tid == 0) { acc[ACC_FAIL] += 1.0; ctl.state &= ~WS_STEP_OK; ctl.have_scale = 1; }
      return;
    }
    CHOL_TS(4);
    CHOL_TS(5);
  } else {
    // ---- large windows (d > packed limit, e.g. the 31-frame stress window, d = 465): the same blocked tensor-core
    //      Cholesky, with the fragment-order factor in a per-window GLOBAL scratch (D.chol_frag) instead of shared memory.
    //      Every operand of the trailing update is still one contiguous 256-byte read per warp (now a coalesced global
    //      load; the current panel - d x 8 doubles - stays in L1), the C tiles are read-modify-written in L2 / HBM:
    //      ~d^3 / 24 x 16 bytes per factorisation.  Visibility between the phases is the CTA barrier (global accesses of a
    //      CTA are ordered by __syncthreads).  The pair table lives in dynamic shared memory behind the prior column map.
    const CholLayout Lo = chol_layout(d);
    const int K = Lo.K;
    double *A = D.chol_frag + (size_t)w * D.chol_frag_stride;
    double *Dg = A + Lo.dbase, *bz = A + Lo.vbase, *invd_all = bz + (size_t)K * NB;
    unsigned short *pairs = reinterpret_cast<unsigned short *>(smem) + MAX_PRIOR_COLS * 2;   // after the 2 KB column map
    vec = bz;
    const int warp = tid >> 5, lane = tid & 31;
    auto rows_of = [&](int I) { return I == K - 1 ? Lo.vr : NB; };
    auto slot = [&](int i, int j) -> double * {   // element (i, j), j <= i, of the lower matrix
      const int I = i >> 3, r = i & 7;
      return (j >> 3) == I ? Dg + 36 * I + r * (r + 1) / 2 + (j & 7) : A + (size_t)32 * I * (I - 1) + (j >> 2) * 4 * rows_of(I) + 4 * r + (j & 3);
    };
    for (int j = warp; j < d; j += nthr / 32) {
      const double *src = Sg + (size_t)j * d;
      const double sj = scale[j];
      for (int i = j + lane; i < d; i += 32) {
        double v = scale[i] * sj * src[i];
        if (i == j) { const double h = sj * sj * colsq[j]; v += clampd2(h, P.min_lm_diag, P.max_lm_diag) / radius; }
        *slot(i, j) = v;
      }
    }
    for (int c = tid; c < K * NB; c += nthr) { bz[c] = c < d ? -scale[c] * gS[c] : 0.0; invd_all[c] = 1.0; }
    __syncthreads();
    if (!blocked_chol_solve(A, Lo, nthr, pairs, s_flag)) {
      if (tid == 0) { acc[ACC_FAIL] += 1.0; ctl.state &= ~WS_STEP_OK; ctl.have_scale = 1; }
      return;
    }
  }

  // ---- delta = s .* y, candidate camera state, step / state norms
  const int cur = D.cur[w];
  const int fo = D.frame_off[w];
  double step2 = 0.0, x2 = 0.0;
  for (int c = tid; c < d; c += nthr) { const double yv = vec[c]; vec_y[c] = yv; D.delta_cam[co + c] = scale[c] * yv; }
  __syncthreads();
  const double *dl = D.delta_cam + co;
  const bool lead = D.nranks <= 1 || D.rank == 0;
  for (int f = tid; f < F; f += nthr) {
    const double *x = D.pose[cur] + 7 * (size_t)(fo + f);
    double *y = D.pose[cur ^ 1] + 7 * (size_t)(fo + f);
    const double *t = dl + 15 * f;
    for (int k = 0; k < 3; k++) y[k] = x[k] + t[k];
    q4 q = qmul(mkq(x[3], x[4], x[5], x[6]), mkq(t[3] / 2.0, t[4] / 2.0, t[5] / 2.0, 1.0));
    const double nq = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    y[3] = q.x / nq; y[4] = q.y / nq; y[5] = q.z / nq; y[6] = q.w / nq;
    for (int k = 0; k < 7; k++) { step2 += (y[k] - x[k]) * (y[k] - x[k]); x2 += x[k] * x[k]; }
    const double *xs = D.sb[cur] + 9 * (size_t)(fo + f);
    double *ys = D.sb[cur ^ 1] + 9 * (size_t)(fo + f);
    for (int k = 0; k < 9; k++) { ys[k] = xs[k] + t[6 + k]; step2 += t[6 + k] * t[6 + k]; x2 += xs[k] * xs[k]; }
  }
  if (tid == 0) {
    const double *x = D.ex[cur] + 7 * (size_t)w;
    double *y = D.ex[cur ^ 1] + 7 * (size_t)w;
    if (fl & WF_EXTRINSIC) {
      const double *t = dl + 15 * F;
      for (int k = 0; k < 3; k++) y[k] = x[k] + t[k];
      q4 q = qmul(mkq(x[3], x[4], x[5], x[6]), mkq(t[3] / 2.0, t[4] / 2.0, t[5] / 2.0, 1.0));
      const double nq = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
      y[3] = q.x / nq; y[4] = q.y / nq; y[5] = q.z / nq; y[6] = q.w / nq;
      for (int k = 0; k < 7; k++) { step2 += (y[k] - x[k]) * (y[k] - x[k]); x2 += x[k] * x[k]; }
    } else {
      for (int k = 0; k < 7; k++) y[k] = x[k];
    }
    const double tdx = D.td[cur][w];
    if (fl & WF_TD) {
      const double t = dl[15 * F + ((fl & WF_EXTRINSIC) ? 6 : 0)];
      D.td[cur ^ 1][w] = tdx + t; step2 += t * t; x2 += tdx * tdx;
    } else {
      D.td[cur ^ 1][w] = tdx;
    }
  }
  step2 = block_sum(step2, red);
  x2 = block_sum(x2, red);

  // ---- model cost change of the camera-only factors: sum (J delta) . (r + J delta / 2)
  {   // prior column -> camera offset (-1: constant block)
    const int n = D.prior_off[w + 1] - D.prior_off[w];
    for (int c = tid; c < n && c < MAX_PRIOR_COLS; c += nthr) s_cmap[c] = -1;
    __syncthreads();
    for (int b = D.pblk_off[w] + tid; b < D.pblk_off[w + 1]; b += nthr) {
      const int kind = D.pblk_kind[b], cam = D.pblk_cam[b], col = D.pblk_col[b];
      const int ls = (kind == 0 || kind == 2) ? 6 : (kind == 1 ? 9 : 1);
      if (cam >= 0) for (int c = 0; c < ls; c++) s_cmap[col + c] = cam + c;
    }
    __syncthreads();
  }
  double mc = 0.0;
  if (lead && mc_identity) {
    // -(g^T y + y^T H y / 2) = (y^T D^2 y - g^T y) / 2 for the solution of (H + D^2) y = -g: camera part
    for (int c = tid; c < d; c += nthr) {
      const double y = vec_y[c];
      const double h = scale[c] * scale[c] * colsq[c];
      mc += 0.5 * (clampd2(h, P.min_lm_diag, P.max_lm_diag) / radius * y * y - scale[c] * gfull[c] * y);
    }
    mc = -mc;   // accumulated below as -mc
  } else if (lead) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int f = D.imu_off[w] + warp; f < D.imu_off[w + 1]; f += nthr / 32) {
      const double *R = D.rec_imu + (size_t)f * REC_IMU;
      const int c0 = 15 * (D.imu_idx[f].x - fo);
      if (lane < 15) {
        double jd = 0.0;
        for (int c = 0; c < 30; c++) jd += R[15 + lane * 30 + c] * dl[c0 + c];
        mc += jd * (R[lane] + 0.5 * jd);
      }
    }
    const int n = D.prior_off[w + 1] - D.prior_off[w];
    if (n > 0) {
      const double *J0 = D.prior_J + D.priorJ_off[w];
      const double *r = D.rec_prior + D.prior_off[w];
      for (int i = warp; i < n; i += nthr / 32) {
        double jd = 0.0;
        for (int c = lane; c < n; c += 32) { const int cam = s_cmap[c]; if (cam >= 0) jd += J0[(size_t)i * n + c] * dl[cam]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) jd += __shfl_xor_sync(0xffffffffu, jd, o);
        if (lane == 0) mc += jd * (r[i] + 0.5 * jd);
      }
    }
  }
  mc = block_sum(mc, red);
#ifdef UVS_CHOL_TIMING
  if (kPacked && threadIdx.x == 0 && blockIdx.x == 0) {
    const long long t6 = clock64();
    printf("chol d=%d cycles: prologue %lld load %lld factor + substitutions %lld tail %lld total %lld\n", d, ts[1] - ts[0], ts[2] - ts[1], ts[4] - ts[2], t6 - ts[4], t6 - ts[0]);
  }
#endif
  if (tid == 0) {
    if (lead) { atomicAdd(acc + ACC_MODEL, -mc); atomicAdd(acc + ACC_STEP2, step2); atomicAdd(acc + ACC_XNORM2, x2); }
    ctl.state |= WS_STEP_OK;
    ctl.have_scale = 1;
  }
}

__global__ void __launch_bounds__(CT) k_chol(Dev D, Params P, int packed_limit, int mc_identity) {
  extern __shared__ double smem[];
  const int w = blockIdx.x;
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  if (D.acc[(size_t)w * ACC_STRIDE + ACC_FAIL] != 0.0) {   // a landmark block was not positive definite
    if (threadIdx.x == 0) { D.ctl[w].state &= ~WS_STEP_OK; D.ctl[w].have_scale = 1; if (D.ctl[w].iter == 0) { D.ctl[w].cost = D.acc[(size_t)w * ACC_STRIDE + ACC_COST0]; D.ctl[w].iter = 1; D.summary[w].initial_cost = D.ctl[w].cost; D.summary[w].cost[0] = D.ctl[w].cost; } }
    return;
  }
  const int d = D.cam_off[w + 1] - D.cam_off[w];
  if (d <= packed_limit) chol_window<1>(D, P, w, smem, mc_identity != 0);
  else chol_window<0>(D, P, w, smem, mc_identity != 0);
}

// chain mode: 256 threads, <= 64 registers so that four windows share an SM
template <int kMode>
__global__ void __launch_bounds__(256, 4) k_chol_chain(Dev D, Params P, int mc_identity) {
  extern __shared__ double smem[];
  const int w = blockIdx.x;
  if (!(D.ctl[w].state & WS_ACTIVE)) return;
  if (D.acc[(size_t)w * ACC_STRIDE + ACC_FAIL] != 0.0) {   // a landmark block was not positive definite
    if (threadIdx.x == 0) { D.ctl[w].state &= ~WS_STEP_OK; D.ctl[w].have_scale = 1; if (D.ctl[w].iter == 0) { D.ctl[w].cost = D.acc[(size_t)w * ACC_STRIDE + ACC_COST0]; D.ctl[w].iter = 1; D.summary[w].initial_cost = D.ctl[w].cost; D.summary[w].cost[0] = D.ctl[w].cost; } }
    zero_window_system(D, w);   // nobody consumes this system: clear it for the rebuild (k_step leaves the matrix to this kernel)
    return;
  }
  chol_window<kMode>(D, P, w, smem, mc_identity != 0);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CT) k_step(Dev D, Params P, int clear_system) {
  const int w = blockIdx.x;
  WinCtl &c = D.ctl[w];
  if (!(c.state & WS_ACTIVE)) return;
  double *acc = D.acc + (size_t)w * ACC_STRIDE;
  if (threadIdx.x == 0) {
    UvsSummary &sm = D.summary[w];
    const bool fixed = P.fixed_iterations != 0;
    const double mc = acc[ACC_MODEL], step_norm = sqrt(acc[ACC_STEP2]), cand_cost = acc[ACC_CAND_COST];
    const bool step_ok = (c.state & WS_STEP_OK) && acc[ACC_FAIL] == 0.0;
    const bool valid = step_ok && mc > 0.0 && isfinite(step_norm) && isfinite(mc);
    const int it = c.iter;   // index of the log entry written now
    double log_rel = 0.0, log_step = 0.0;
    int accepted = 0, done = 0;
    const double gprev = it - 1 < UVS_MAX_ITER_LOG ? sm.gradient_max_norm[it - 1] : 0.0;
    c.state &= ~WS_NEED_JAC;
    if (!valid) {   // invalid step: the LM strategy treats it as a rejected step
      c.radius /= c.decrease_factor; c.decrease_factor *= 2.0;
      if (++c.n_invalid >= 5 && !fixed) { c.termination = UVS_TERM_FAILURE; done = 1; }
      else if (!fixed && c.radius < P.min_radius) { c.termination = UVS_TERM_MIN_RADIUS; done = 1; }
    } else {
      c.n_invalid = 0;
      c.x_norm = sqrt(acc[ACC_XNORM2]);
      log_step = step_norm;
      const double cost_change = c.cost - cand_cost;
      if (!fixed && step_norm <= P.parameter_tolerance * (c.x_norm + P.parameter_tolerance)) {
        c.termination = UVS_TERM_PARAMETER_TOL; done = 1;
      } else if (!fixed && fabs(cost_change) <= P.function_tolerance * c.cost) {
        c.termination = UVS_TERM_FUNCTION_TOL; done = 1;
      } else {
        const double rel = cost_change / mc;
        log_rel = rel;
        accepted = isfinite(cand_cost) && rel > P.min_relative_decrease;
        if (accepted) {
          D.cur[w] ^= 1;
          c.cost = cand_cost;
          const double t = 2.0 * rel - 1.0;
          c.radius = fmin(P.max_radius, c.radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
          c.decrease_factor = 2.0;
          c.n_success++;
          c.state |= WS_NEED_JAC;
        } else {
          c.radius /= c.decrease_factor; c.decrease_factor *= 2.0;
        }
        if (!fixed && c.radius < P.min_radius) { c.termination = UVS_TERM_MIN_RADIUS; done = 1; }
      }
    }
    if (it < UVS_MAX_ITER_LOG) {
      sm.cost[it] = c.cost; sm.radius[it] = c.radius; sm.relative_decrease[it] = log_rel; sm.step_norm[it] = log_step;
      sm.gradient_max_norm[it] = accepted ? -1.0 : gprev;   // filled by the next linearisation when accepted
      sm.step_accepted[it] = accepted;
    }
    c.iter = it + 1;
    if (!done && c.iter - 1 >= P.max_num_iterations) { c.termination = UVS_TERM_NO_CONVERGENCE; done = 1; }
    if (done) c.state &= ~WS_ACTIVE;
    for (int e = 0; e < ACC_STRIDE; e++) acc[e] = 0.0;
  }
  // chain mode: k_chol_chain has cleared the matrix while consuming it; only the three vectors are left
  if (clear_system) zero_window_system(D, w);
  else {
    const int co = D.cam_off[w], d = D.cam_off[w + 1] - co;
    for (int e = threadIdx.x; e < d; e += blockDim.x) { D.gS[co + e] = 0.0; D.gfull[co + e] = 0.0; D.colsq_cam[co + e] = 0.0; }
  }
}

// final bookkeeping: summaries (device copy), number of active windows
__global__ void k_finish(Dev D) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= D.B) return;
  const WinCtl &c = D.ctl[w];
  UvsSummary &sm = D.summary[w];
  sm.num_iterations = c.iter;
  sm.num_successful_steps = c.n_success;
  sm.termination = c.termination;
  sm.status = isfinite(c.cost) ? c.status : UVS_ERR_NOT_FINITE;
  sm.final_cost = c.cost;
}

__global__ void k_count_active(Dev D, int *out) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= D.B) return;
  if (D.ctl[w].state & WS_ACTIVE) atomicAdd(out, 1);
}

// per-window cost at the current iterate for uvs_eval_cost: acc[ACC_CAND_COST] -> out
__global__ void k_copy_acc(Dev D, int slot, double *out, int zero) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= D.B) return;
  out[w] = D.acc[(size_t)w * ACC_STRIDE + slot];
  if (zero) D.acc[(size_t)w * ACC_STRIDE + slot] = 0.0;
}

static size_t chol_smem_bytes(int d) {
  const size_t K = (d + NB - 1) / NB;
  (void)K;
  return std::max((size_t)chol_layout(d).total * sizeof(double), (size_t)MAX_PRIOR_COLS * sizeof(int));
}
int chol_packed_limit(size_t max_smem) {
  int d = NB;
  while (d < 256 && chol_smem_bytes(d + 1) <= max_smem) d++;   // 256: the load phase keeps 8 x 32 entries of a row in flight
  return d;
}

int launch_solve_init(const Dev &D, const Params &P, cudaStream_t st) { k_solve_init<<<D.B, CT, 0, st>>>(D, P); return 1; }

size_t chol_chain_smem(int max_frames, bool any_ex) {
  const int nd = 6 * max_frames + (any_ex ? 6 : 0);
  return (size_t)std::max(chain_layout(nd, max_frames).total, chain_layout_v1(nd, max_frames).total) * sizeof(double);
}
int chol_chain_lw_doubles(int max_frames) { return max_frames * CH_LWG; }
long long chol_frag_doubles(int d) { return ((long long)chol_layout(d).total + 1) & ~1LL; }
// dynamic shared memory of the large-window mode: prior column map + pair table of the K (K + 1) / 2 lower blocks
static size_t chol_large_smem(int d) { const size_t K = (d + NB - 1) / NB; return (size_t)MAX_PRIOR_COLS * sizeof(int) + K * (K + 1) / 2 * sizeof(unsigned short) + 16; }

// Two variants of the chain solve.  Measured (us per LM iteration, B windows; gpurun_out/c*_scan): pipeline 92 / 95 / 107 /
// 178 / 392 against barrier-phased 107 / 109 / 108 / 119 / 161 at B = 1 / 16 / 64 / 148 / 592.  The pipeline has the shorter
// critical path, but its row threads read W and write L_w / the cleared entries of S with plain 8-byte global accesses
// between named barriers (a barrier waits for the warp's outstanding global accesses), so it degrades as soon as many
// windows share the memory system - even on different SMs; the barrier-phased variant prefetches with cp.async.
int launch_chol_chain(const Dev &D, const Params &P, int max_frames, bool any_ex, bool mc_identity, cudaStream_t st) {
  static const size_t pad = std::getenv("UVS_CHAIN_PAD_SMEM") ? (size_t)std::atoi(std::getenv("UVS_CHAIN_PAD_SMEM")) : 0;   // developer knob: fewer windows per SM
  static const int pipe_max = std::getenv("UVS_CHAIN_PIPE_MAX") ? std::atoi(std::getenv("UVS_CHAIN_PIPE_MAX")) : 32;    // largest batch that takes the pipeline
  const size_t smem = std::max(chol_chain_smem(max_frames, any_ex), (size_t)MAX_PRIOR_COLS * sizeof(int)) + pad;
  static size_t raised = 0;
  if (smem > raised) {
    cudaFuncSetAttribute(k_chol_chain<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_chol_chain<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_chol_chain<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_chol_chain<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    raised = smem;
  }
  // each variant is launched with its own layout's size: four windows of the barrier-phased one must keep fitting an SM
  const int nd = 6 * max_frames + (any_ex ? 6 : 0);
  const size_t smem_v1 = std::max((size_t)chain_layout_v1(nd, max_frames).total * sizeof(double), (size_t)MAX_PRIOR_COLS * sizeof(int)) + pad;
  if (D.B <= pipe_max) k_chol_chain<3><<<D.B, 256, smem, st>>>(D, P, mc_identity ? 1 : 0);
  else k_chol_chain<2><<<D.B, 256, smem_v1, st>>>(D, P, mc_identity ? 1 : 0);
  return 1;
}

int launch_chol(const Dev &D, const Params &P, int max_d, int packed_limit, bool mc_identity, cudaStream_t st) {
  if (max_d > packed_limit) {
    // a batch with a window too large for shared memory: every window takes the global-scratch mode (little shared memory,
    // several windows per SM)
    k_chol<<<D.B, CT, chol_large_smem(max_d), st>>>(D, P, 0, mc_identity ? 1 : 0);
    return 1;
  }
  const int dd = max_d;
  const size_t smem = chol_smem_bytes(dd);
  // two windows per SM when two factors fit its shared memory (d <= 165): 256 threads each, else one window with 512
  static int sm_smem = 0, static_smem = 0;
  if (!sm_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaFuncAttributes attr;
    if (cudaFuncGetAttributes(&attr, k_chol) == cudaSuccess) static_smem = (int)attr.sharedSizeBytes;
  }
  const bool two = 2 * (smem + static_smem + 1024) <= (size_t)sm_smem;
  static const int t2 = std::getenv("UVS_CHOL_T2") ? std::atoi(std::getenv("UVS_CHOL_T2")) : CT;   // tuning knob (64 registers per thread: two 512-thread CTAs fit)
  k_chol<<<D.B, two ? t2 : CT, smem, st>>>(D, P, packed_limit, mc_identity ? 1 : 0);
  return 1;
}
// largest dynamic shared-memory size k_chol can be launched with (opt-in limit minus its static arrays);
// raises the kernel's limit to that value.  Returns 0 on failure.
size_t chol_max_dynamic_smem(size_t optin_bytes) {
  cudaFuncAttributes attr;
  if (cudaFuncGetAttributes(&attr, k_chol) != cudaSuccess) return 0;
  if (optin_bytes <= attr.sharedSizeBytes + 512) return 0;
  const size_t dyn = optin_bytes - attr.sharedSizeBytes - 512;
  if (cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return 0;
  cudaFuncSetAttribute(k_chol, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return dyn;
}

int launch_step(const Dev &D, const Params &P, bool clear_system, cudaStream_t st) {
  if (clear_system) k_step<<<D.B, CT, 0, st>>>(D, P, 1);
  else k_step<<<D.B, 192, 0, st>>>(D, P, 0);
  return 1;
}
int launch_finish(const Dev &D, cudaStream_t st) { k_finish<<<(D.B + 127) / 128, 128, 0, st>>>(D); return 1; }
int launch_count_active(const Dev &D, int *out, cudaStream_t st) {
  cudaMemsetAsync(out, 0, sizeof(int), st);
  k_count_active<<<(D.B + 127) / 128, 128, 0, st>>>(D, out);
  return 1;
}
int launch_copy_acc(const Dev &D, int slot, double *out, int zero, cudaStream_t st) {
  k_copy_acc<<<(D.B + 127) / 128, 128, 0, st>>>(D, slot, out, zero);
  return 1;
}

}  // namespace uvs
