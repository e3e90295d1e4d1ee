// Micro-benchmark: does the fp64 pipe of sm_100a issue a DFMA of a half-masked warp in half the time?
// One warp per SM sub-partition (4 warps per block, one block per SM), 8 independent DFMA chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters, double b, double c, int active_lanes, int stride) {
  const int lane = threadIdx.x & 31;
  if ((lane % stride) != 0 || (lane / stride) >= active_lanes) return;
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
// dependent chain: latency of one DFMA
__global__ void lat(double* out, int iters, double b, double c) {
  double a0 = threadIdx.x;
  for (int i = 0; i < iters; ++i) { a0 = fma(a0, b, c); a0 = fma(a0, b, c); a0 = fma(a0, b, c); a0 = fma(a0, b, c); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 200000;
  struct { int warps, lanes, stride; } cfg[] = {{4, 32, 1}, {4, 16, 1}, {4, 16, 2}, {4, 8, 1}, {4, 8, 4}, {4, 4, 1}, {8, 32, 1}, {8, 16, 1}, {16, 32, 1}};
  for (auto c : cfg) {
    k<<<148, c.warps * 32>>>(out, 1000, 1.0000001, 1e-9, c.lanes, c.stride);
    cudaEventRecord(e0);
    k<<<148, c.warps * 32>>>(out, iters, 1.0000001, 1e-9, c.lanes, c.stride);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * clk * 1e3 / (iters * 8.0 * (c.warps / 4.0));
    printf("warps/SM=%d active lanes=%d stride=%d: %.3f ms, %.2f cycles per warp-DFMA per sub-partition (clock %d kHz)\n", c.warps, c.lanes, c.stride, ms, cyc, clk);
  }
  lat<<<148, 128>>>(out, 1000, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  lat<<<148, 128>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("dependent DFMA latency: %.2f cycles\n", ms * 1e-3 * clk * 1e3 / (iters * 4.0));
  return 0;
}
