// FP64 throughput probe on sm_100a: DFMA (vector pipe) vs DMMA (mma.sync m8n8k4 f64 tensor path).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double *out, int iters, double a, double b) {
  double c[NACC][2];
  for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double c[NACC];
  for (int i = 0; i < NACC; i++) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// correctness of the fragment layout: C = A(8x4) * B(4x8)
__global__ void k_check(const double *A, const double *B, double *C) {
  const int l = threadIdx.x;
  double c0 = 0, c1 = 0;
  dmma(c0, c1, A[(l / 4) * 4 + l % 4], B[(l % 4) * 8 + l / 4]);
  C[(l / 4) * 8 + 2 * (l % 4)] = c0; C[(l / 4) * 8 + 2 * (l % 4) + 1] = c1;
}

int main() {
  double *out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    float ms;
    k_dmma<8><<<148, warps * 32>>>(out, 100, 1.0, 1e-9);
    cudaEventRecord(e0); k_dmma<8><<<148, warps * 32>>>(out, iters, 1.0, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA warps/SM=%2d: %.2f TFLOP/s\n", warps, 2.0 * 256 * 8 * (double)iters * warps * 148 / (ms * 1e-3) / 1e12);
    k_dfma<16><<<148, warps * 32>>>(out, 100, 1.0, 1e-9);
    cudaEventRecord(e0); k_dfma<16><<<148, warps * 32>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA warps/SM=%2d: %.2f TFLOP/s\n", warps, 2.0 * 32 * 16 * (double)iters * warps * 148 / (ms * 1e-3) / 1e12);
  }
  double hA[32], hB[32], hC[64], *dA, *dB, *dC;
  for (int i = 0; i < 32; i++) { hA[i] = i + 1; hB[i] = 0.5 * i - 3; }
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  k_check<<<1, 32>>>(dA, dB, dC); cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { double r = 0; for (int k = 0; k < 4; k++) r += hA[i * 4 + k] * hB[k * 8 + j]; maxerr = fmax(maxerr, fabs(r - hC[i * 8 + j])); }
  printf("layout check max err %.3g (%s)\n", maxerr, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
