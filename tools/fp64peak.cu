// fp64peak.cu — sustained DFMA throughput per SM on this device (the fp64 pipe roofline of the march kernel)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *out; cudaMalloc(&out, 8 * 148 * 1024 * 8);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int warps : {4, 8, 16, 32}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<8><<<pr.multiProcessorCount, warps * 32>>>(out, 0.999, 1e-3, 100);
    cudaEventRecord(e0);
    k<8><<<pr.multiProcessorCount, warps * 32>>>(out, 0.999, 1e-3, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma_total = (double)pr.multiProcessorCount * warps * 32 * 8 * iters;
    printf("warps/SM=%2d: %.2f TFLOP/s fp64 (2 flop/FMA), %.1f DFMA lanes/clk/SM at %d MHz nominal\n", warps,
           2 * fma_total / (ms * 1e-3) / 1e12, fma_total / (ms * 1e-3) / pr.multiProcessorCount / (clk * 1e3), clk / 1000);
  }
  return 0;
}
