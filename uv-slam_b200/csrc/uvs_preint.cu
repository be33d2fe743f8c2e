// uvs_preint.cu — IMU mid-point preintegration on the device (the step right before the hot path).
//
// Replaces IntegrationBase::{push_back, propagate, midPointIntegration, repropagate}
// (vins_estimator/src/factor/integration_base.h:30-158) for many keyframe intervals at once: one WARP per
// interval walks its IMU samples in order; the small vector algebra is computed by every lane, the 15x15
// products  jacobian <- F jacobian,  covariance <- F cov F^T + V N V^T  are spread over the lanes (shared memory).
// `repropagate` with new linearisation biases is the same call with different lin_ba / lin_bg.
#include <cstring>
#include "uvs_handle.h"
#include "uvs_kernels.h"
#include "uvs_math.cuh"

namespace uvs {

struct PreintArgs {
  int n;
  const int *off;                    // [n+1] sample ranges
  const double *dt, *acc, *gyr;      // [S], [S][3], [S][3]
  const double *acc0, *gyr0, *ba, *bg;   // [n][3]
  double noise[4];                   // acc_n, gyr_n, acc_w, gyr_w
  double *dp, *dq, *dv, *sum_dt, *jac, *cov;
};

__device__ __forceinline__ void put_blk(double *M, int ld, int r0, int c0, const m33 &B, double s) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) M[(r0 + i) * ld + c0 + j] = s * B.a[3 * i + j];
}

__global__ void __launch_bounds__(128) k_preintegrate(PreintArgs A) {
  __shared__ double sm[4][4 * 225 + 270];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 4 + warp;
  if (k >= A.n) return;
  double *jac = sm[warp], *cov = jac + 225, *F = cov + 225, *tmp = F + 225, *V = tmp + 225;
  for (int e = lane; e < 225; e += 32) { jac[e] = (e / 15 == e % 15) ? 1.0 : 0.0; cov[e] = 0.0; }
  const d3 ba = mk3(A.ba[3 * k], A.ba[3 * k + 1], A.ba[3 * k + 2]), bg = mk3(A.bg[3 * k], A.bg[3 * k + 1], A.bg[3 * k + 2]);
  d3 a0 = mk3(A.acc0[3 * k], A.acc0[3 * k + 1], A.acc0[3 * k + 2]), g0 = mk3(A.gyr0[3 * k], A.gyr0[3 * k + 1], A.gyr0[3 * k + 2]);
  d3 dp = mk3(0, 0, 0), dv = mk3(0, 0, 0);
  q4 dq = mkq(0, 0, 0, 1);
  double sum_dt = 0.0;
  const double nd[6] = {A.noise[0] * A.noise[0], A.noise[1] * A.noise[1], A.noise[0] * A.noise[0], A.noise[1] * A.noise[1],
                        A.noise[2] * A.noise[2], A.noise[3] * A.noise[3]};
  __syncwarp();
  for (int s = A.off[k]; s < A.off[k + 1]; s++) {
    const double dt = A.dt[s];
    const d3 a1 = mk3(A.acc[3 * s], A.acc[3 * s + 1], A.acc[3 * s + 2]), g1 = mk3(A.gyr[3 * s], A.gyr[3 * s + 1], A.gyr[3 * s + 2]);
    // midPointIntegration, integration_base.h:54-128
    const d3 un_acc_0 = qrot(dq, a0 - ba);
    const d3 un_gyr = 0.5 * (g0 + g1) - bg;
    const q4 rq = qmul(dq, mkq(un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2, 1.0));
    const d3 un_acc_1 = qrot(rq, a1 - ba);
    const d3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
    const d3 rp = dp + dt * dv + (0.5 * dt * dt) * un_acc;
    const d3 rv = dv + dt * un_acc;
    // F (15x15) and V (15x18)
    for (int e = lane; e < 225; e += 32) F[e] = 0.0;
    for (int e = lane; e < 270; e += 32) V[e] = 0.0;
    __syncwarp();
    if (lane == 0) {
      const m33 Rq = qmat(dq), Rr = qmat(rq);
      const m33 Sw = skew(un_gyr), Sa0 = skew(a0 - ba), Sa1 = skew(a1 - ba);
      m33 I; for (int e = 0; e < 9; e++) I.a[e] = (e % 4 == 0) ? 1.0 : 0.0;
      m33 ImW; for (int e = 0; e < 9; e++) ImW.a[e] = I.a[e] - Sw.a[e] * dt;
      const m33 RqA0 = mmul(Rq, Sa0), RrA1 = mmul(Rr, Sa1), RrA1W = mmul(RrA1, ImW);
      m33 t;
      put_blk(F, 15, 0, 0, I, 1.0);
      for (int e = 0; e < 9; e++) t.a[e] = RqA0.a[e] * (-0.25 * dt * dt) + RrA1W.a[e] * (-0.25 * dt * dt);
      put_blk(F, 15, 0, 3, t, 1.0);
      put_blk(F, 15, 0, 6, I, dt);
      for (int e = 0; e < 9; e++) t.a[e] = (Rq.a[e] + Rr.a[e]) * (-0.25 * dt * dt);
      put_blk(F, 15, 0, 9, t, 1.0);
      put_blk(F, 15, 0, 12, RrA1, -0.25 * dt * dt * -dt);
      put_blk(F, 15, 3, 3, ImW, 1.0);
      put_blk(F, 15, 3, 12, I, -1.0 * dt);
      for (int e = 0; e < 9; e++) t.a[e] = RqA0.a[e] * (-0.5 * dt) + RrA1W.a[e] * (-0.5 * dt);
      put_blk(F, 15, 6, 3, t, 1.0);
      put_blk(F, 15, 6, 6, I, 1.0);
      for (int e = 0; e < 9; e++) t.a[e] = (Rq.a[e] + Rr.a[e]) * (-0.5 * dt);
      put_blk(F, 15, 6, 9, t, 1.0);
      put_blk(F, 15, 6, 12, RrA1, -0.5 * dt * -dt);
      put_blk(F, 15, 9, 9, I, 1.0);
      put_blk(F, 15, 12, 12, I, 1.0);
      m33 nRrA1; for (int e = 0; e < 9; e++) nRrA1.a[e] = -RrA1.a[e];
      put_blk(V, 18, 0, 0, Rq, 0.25 * dt * dt);
      put_blk(V, 18, 0, 3, nRrA1, 0.25 * dt * dt * 0.5 * dt);
      put_blk(V, 18, 0, 6, Rr, 0.25 * dt * dt);
      put_blk(V, 18, 0, 9, nRrA1, 0.25 * dt * dt * 0.5 * dt);
      put_blk(V, 18, 3, 3, I, 0.5 * dt);
      put_blk(V, 18, 3, 9, I, 0.5 * dt);
      put_blk(V, 18, 6, 0, Rq, 0.5 * dt);
      put_blk(V, 18, 6, 3, nRrA1, 0.5 * dt * 0.5 * dt);
      put_blk(V, 18, 6, 6, Rr, 0.5 * dt);
      put_blk(V, 18, 6, 9, nRrA1, 0.5 * dt * 0.5 * dt);
      put_blk(V, 18, 9, 12, I, dt);
      put_blk(V, 18, 12, 15, I, dt);
    }
    __syncwarp();
    // jacobian = F * jacobian
    for (int e = lane; e < 225; e += 32) {
      const int i = e / 15, j = e - 15 * i;
      double a = 0.0;
      for (int m = 0; m < 15; m++) a += F[i * 15 + m] * jac[m * 15 + j];
      tmp[e] = a;
    }
    __syncwarp();
    for (int e = lane; e < 225; e += 32) jac[e] = tmp[e];
    __syncwarp();
    // covariance = F cov F^T + V N V^T
    for (int e = lane; e < 225; e += 32) {
      const int i = e / 15, j = e - 15 * i;
      double a = 0.0;
      for (int m = 0; m < 15; m++) a += F[i * 15 + m] * cov[m * 15 + j];
      tmp[e] = a;
    }
    __syncwarp();
    for (int e = lane; e < 225; e += 32) {
      const int i = e / 15, j = e - 15 * i;
      double a = 0.0;
      for (int m = 0; m < 15; m++) a += tmp[i * 15 + m] * F[j * 15 + m];
      double b = 0.0;
      for (int m = 0; m < 18; m++) b += V[i * 18 + m] * nd[m / 3] * V[j * 18 + m];
      cov[e] = a + b;
    }
    __syncwarp();
    dp = rp; dv = rv;
    const double nq = sqrt(rq.x * rq.x + rq.y * rq.y + rq.z * rq.z + rq.w * rq.w);
    dq = mkq(rq.x / nq, rq.y / nq, rq.z / nq, rq.w / nq);
    sum_dt += dt;
    a0 = a1; g0 = g1;
  }
  if (lane == 0) {
    A.dp[3 * k] = dp.x; A.dp[3 * k + 1] = dp.y; A.dp[3 * k + 2] = dp.z;
    A.dq[4 * k] = dq.x; A.dq[4 * k + 1] = dq.y; A.dq[4 * k + 2] = dq.z; A.dq[4 * k + 3] = dq.w;
    A.dv[3 * k] = dv.x; A.dv[3 * k + 1] = dv.y; A.dv[3 * k + 2] = dv.z;
    A.sum_dt[k] = sum_dt;
  }
  for (int e = lane; e < 225; e += 32) { A.jac[225 * (size_t)k + e] = jac[e]; A.cov[225 * (size_t)k + e] = cov[e]; }
}

}  // namespace uvs

using namespace uvs;

extern "C" int uvs_preintegrate(UvsHandle *h, int32_t n, const int32_t *sample_off, const double *dt, const double *acc, const double *gyr,
                                const double *acc0, const double *gyr0, const double *lin_ba, const double *lin_bg, const double *noise,
                                double *delta_p, double *delta_q, double *delta_v, double *sum_dt, double *jacobian, double *covariance) {
  if (!h || n <= 0 || !sample_off || !dt || !acc || !gyr || !acc0 || !gyr0 || !lin_ba || !lin_bg || !noise || !delta_p || !delta_q ||
      !delta_v || !sum_dt || !jacobian || !covariance)
    return handle_fail(h, UVS_ERR_INVALID_ARG, "uvs_preintegrate: bad arguments");
  if (cudaSetDevice(h->device) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "cudaSetDevice");
  const size_t S = (size_t)sample_off[n];
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t Dd = sizeof(double);
  // inputs then outputs in one scratch block
  size_t o = 0;
  const size_t o_off = o; o += al((n + 1) * sizeof(int));
  const size_t o_dt = o; o += al(S * Dd);
  const size_t o_acc = o; o += al(S * 3 * Dd);
  const size_t o_gyr = o; o += al(S * 3 * Dd);
  const size_t o_a0 = o; o += al((size_t)n * 3 * Dd);
  const size_t o_g0 = o; o += al((size_t)n * 3 * Dd);
  const size_t o_ba = o; o += al((size_t)n * 3 * Dd);
  const size_t o_bg = o; o += al((size_t)n * 3 * Dd);
  const size_t in_end = o;
  const size_t o_dp = o; o += al((size_t)n * 3 * Dd);
  const size_t o_dq = o; o += al((size_t)n * 4 * Dd);
  const size_t o_dv = o; o += al((size_t)n * 3 * Dd);
  const size_t o_sd = o; o += al((size_t)n * Dd);
  const size_t o_j = o; o += al((size_t)n * 225 * Dd);
  const size_t o_c = o; o += al((size_t)n * 225 * Dd);
  int rc = handle_ensure_scratch(h, o); if (rc) return rc;
  rc = handle_ensure_hscratch(h, o); if (rc) return rc;
  char *hs = h->hscratch.base, *ds = h->scratch.base;
  std::memcpy(hs + o_off, sample_off, (n + 1) * sizeof(int));
  std::memcpy(hs + o_dt, dt, S * Dd); std::memcpy(hs + o_acc, acc, S * 3 * Dd); std::memcpy(hs + o_gyr, gyr, S * 3 * Dd);
  std::memcpy(hs + o_a0, acc0, (size_t)n * 3 * Dd); std::memcpy(hs + o_g0, gyr0, (size_t)n * 3 * Dd);
  std::memcpy(hs + o_ba, lin_ba, (size_t)n * 3 * Dd); std::memcpy(hs + o_bg, lin_bg, (size_t)n * 3 * Dd);
  cudaStream_t st = h->stream;
  if (cudaMemcpyAsync(ds, hs, in_end, cudaMemcpyHostToDevice, st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_preintegrate: H2D");
  PreintArgs A;
  A.n = n; A.off = (const int *)(ds + o_off); A.dt = (const double *)(ds + o_dt); A.acc = (const double *)(ds + o_acc);
  A.gyr = (const double *)(ds + o_gyr); A.acc0 = (const double *)(ds + o_a0); A.gyr0 = (const double *)(ds + o_g0);
  A.ba = (const double *)(ds + o_ba); A.bg = (const double *)(ds + o_bg);
  for (int k = 0; k < 4; k++) A.noise[k] = noise[k];
  A.dp = (double *)(ds + o_dp); A.dq = (double *)(ds + o_dq); A.dv = (double *)(ds + o_dv); A.sum_dt = (double *)(ds + o_sd);
  A.jac = (double *)(ds + o_j); A.cov = (double *)(ds + o_c);
  k_preintegrate<<<(n + 3) / 4, 128, 0, st>>>(A);
  h->launches++;
  if (cudaGetLastError() != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_preintegrate: launch");
  if (cudaMemcpyAsync(hs + o_dp, ds + o_dp, o - o_dp, cudaMemcpyDeviceToHost, st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_preintegrate: D2H");
  if (cudaStreamSynchronize(st) != cudaSuccess) return handle_fail(h, UVS_ERR_CUDA, "uvs_preintegrate: sync");
  std::memcpy(delta_p, hs + o_dp, (size_t)n * 3 * Dd); std::memcpy(delta_q, hs + o_dq, (size_t)n * 4 * Dd);
  std::memcpy(delta_v, hs + o_dv, (size_t)n * 3 * Dd); std::memcpy(sum_dt, hs + o_sd, (size_t)n * Dd);
  std::memcpy(jacobian, hs + o_j, (size_t)n * 225 * Dd); std::memcpy(covariance, hs + o_c, (size_t)n * 225 * Dd);
  return UVS_OK;
}
