// Dependent-chain latencies on sm_100a (one warp): DFMA, DMMA m8n8k4, rsqrt(double), SHFL of a double, LDS.64.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_lat(double *out, long long *cyc, int n, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x] = 1.0 + 1e-9 * threadIdx.x; sm[threadIdx.x + 32] = 0.5;
  __syncthreads();
  double x = a, c0 = 0, c1 = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) x = fma(x, b, a);
  long long t1 = clock64();
  for (int i = 0; i < n; i++) dmma(c0, c1, a, c0 * 1e-30 + b);   // B depends on the previous result
  long long t2 = clock64();
  double y = 2.0 + x * 1e-30;
  for (int i = 0; i < n; i++) y = rsqrt(y) + 1.5;
  long long t3 = clock64();
  double z = y;
  for (int i = 0; i < n; i++) z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31) + 1e-9;
  long long t4 = clock64();
  double w = z; int idx = threadIdx.x;
  for (int i = 0; i < n; i++) { w += sm[idx]; idx = (idx + (w > 1e300 ? 1 : 0)) & 63; }
  long long t5 = clock64();
  double c2 = 0, c3 = 0;
  for (int i = 0; i < n; i++) dmma(c2, c3, a, b);   // accumulator chain only
  long long t6 = clock64();
  out[threadIdx.x] = x + c0 + c1 + y + z + w + c2 + c3;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
}

int main() {
  double *out; long long *cyc, h[6];
  cudaMalloc(&out, 32 * sizeof(double)); cudaMalloc(&cyc, sizeof h);
  const int n = 4096;
  k_lat<<<1, 32>>>(out, cyc, n, 1.0000001, 0.999999);
  k_lat<<<1, 32>>>(out, cyc, n, 1.0000001, 0.999999);
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const char *name[6] = {"DFMA dependent", "DMMA (B <- D) dependent", "rsqrt(double)+DADD", "SHFL(double)+DADD", "LDS.64+DADD", "DMMA accumulator chain"};
  for (int k = 0; k < 6; k++) printf("%-28s %.1f cycles\n", name[k], (double)h[k] / n);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
