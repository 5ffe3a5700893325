// FP64 latency / throughput probe for the epilogue design (dependent chain vs independent chains,
// 1 or 4 warps per SM sub-partition).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
void run(int threads, int iters) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  k<CHAINS><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<CHAINS><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / ((double)iters * CHAINS);
  double tf = 2.0 * 148.0 * threads * (double)iters * CHAINS / (ms * 1e-3) / 1e12;
  printf("chains %d threads/SM %4d: %.2f cycles per DFMA per warp (issue), %.1f cycles per dependent step, %.2f TFLOP/s\n",
         CHAINS, threads, per, (double)h / iters, tf);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  const int it = 20000;
  run<1>(32, it); run<1>(128, it); run<1>(1024, it);
  run<4>(128, it); run<8>(128, it); run<8>(256, it); run<8>(1024, it);
  return 0;
}
