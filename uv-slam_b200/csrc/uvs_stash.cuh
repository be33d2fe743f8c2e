// uvs_stash.cuh — pieces shared by the landmark-path kernels (uvs_build3.cu: records from HBM; uvs_lin.cu: factors
// evaluated in registers): the Y stash, lane-group reductions, asynchronous copies, the FP64 tensor-core MMA wrapper.
#pragma once
#include "uvs_device.cuh"

namespace uvs {

__device__ __forceinline__ double clamp4(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }
__device__ __forceinline__ void atomic_max_nn3(double *addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ int pair_key(int a, int b) { return a <= b ? b * (b + 1) / 2 + a : a * (a + 1) / 2 + b; }
__device__ __forceinline__ void unrank_key(int t, int &a, int &b) {
  int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (i * (i + 1) / 2 > t) i--;
  while ((i + 1) * (i + 2) / 2 <= t) i++;
  b = i; a = t - i * (i + 1) / 2;
}

// warp-aggregated per-window accumulation (threads of a warp usually share the window)
__device__ __forceinline__ void add_win3(double *acc, int win, bool valid, double v0, double v1, double v2) {
  const unsigned full = 0xffffffffu;
  const int w0 = __shfl_sync(full, win, 0);
  const bool v00 = __shfl_sync(full, (int)valid, 0) != 0;
  const bool uniform = __all_sync(full, !valid || win == w0) && v00;
  if (uniform) {
    double a = valid ? v0 : 0.0, b = valid ? v1 : 0.0, c = valid ? v2 : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(full, a, o); b += __shfl_down_sync(full, b, o); c += __shfl_down_sync(full, c, o); }
    if ((threadIdx.x & 31) == 0) {
      double *p = acc + (size_t)w0 * ACC_STRIDE;
      atomicAdd(p + ACC_MODEL, a); atomicAdd(p + ACC_STEP2, b); atomicAdd(p + ACC_XNORM2, c);
    }
  } else if (valid) {
    double *p = acc + (size_t)win * ACC_STRIDE;
    atomicAdd(p + ACC_MODEL, v0); atomicAdd(p + ACC_STEP2, v1); atomicAdd(p + ACC_XNORM2, v2);
  }
}

// warp-aggregated accumulation of one per-window scalar
__device__ __forceinline__ void add_window_scalar(double *arr, int stride_doubles, int win, double v, bool valid) {
  const unsigned full = 0xffffffffu;
  const int w0 = __shfl_sync(full, win, 0);
  const bool uniform = __all_sync(full, !valid || win == w0) && __shfl_sync(full, (int)valid, 0);
  if (uniform) {
    double s = valid ? v : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(full, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(arr + (size_t)w0 * stride_doubles, s);
  } else if (valid) {
    atomicAdd(arr + (size_t)win * stride_doubles, v);
  }
}

// ------------------------------------------------------------------------------------------------
// stash (B3): dense landmark columns  Y[col][mp]:  entries [6 blk + k] = Y of camera block blk, [mp-2] = z,
//             col = colbase(w) + point  |  colbase(w) + np_w + 4 line + sub, colbase(w) = point_off[w] + 4 line_off[w]
//             (the window kernel streams whole chunks of columns with TMA bulk copies);
//             headers for the back-substitution: points ph[4] = sk, sh, D2, -;  lines lh[24] = s(4) D2(4) Linv(16)
//             entry [mp-1] of a column = its scale c: the rank update adds  -c y y^T.  Line columns and the point columns of
//             the record path hold scaled entries and c = 1 (set once at upload); the fused point kernel leaves the
//             entries unscaled (W_j = Jj^T Jl, W_i, g) and stores c = (sk sh)^2, so nothing is rewritten once the
//             landmark's sums are known (`unscaled_pts`).
struct Stash {
  double *Y, *ph, *lh;
  int mp;
  int unscaled_pts;
};
__device__ __forceinline__ long long colbase(const Dev &D, int w) { return (long long)D.point_off[w] + 4LL * D.line_off[w]; }

// sum over the 2^k lanes of a landmark's lane group (all lanes end with the same bits)
template <int kLanes>
__device__ __forceinline__ double group_sum(unsigned gmask, double v) {
#pragma unroll
  for (int o = 1; o < kLanes; o <<= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

constexpr int LPP = 4;   // lanes per point: one observation each (a C2 point has ~4), partial sums merged by shuffles
constexpr int LPL = 8;   // lanes per line  (a C2 line has ~7 observations)

// asynchronous global -> shared copies (LDGSTS): whole chunks of records are in flight at once without holding registers
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }


// FP64 tensor-core MMA  D(8x8) += A(8x4) B(4x8)  (mma.sync m8n8k4: lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) + {0,1}])
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

}  // namespace uvs
