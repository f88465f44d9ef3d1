// microbench.cu — dependent-issue latencies that bound one contour step on sm_100a:
// DFMA chain, 64-bit SHFL chain, LDS round trip, block barrier.  nvcc -arch=sm_100a -o mb microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b) {
  __shared__ double sm[256];
  double x = threadIdx.x * 1e-3;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b));
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b));
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // SHFL(64-bit) + DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) { double y = __shfl_down_sync(0xffffffffu, x, 1); x = fma(y, a, x); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // STS + LDS round trip chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) { sm[threadIdx.x] = x; __syncwarp(); x = sm[threadIdx.x ^ 1] + b; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // barrier chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) { __syncthreads(); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 4 independent DFMA chains (throughput per warp)
  double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(y0) : "d"(a), "d"(b)); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(y1) : "d"(a), "d"(b)); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(y2) : "d"(a), "d"(b)); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(y3) : "d"(a), "d"(b)); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  out[threadIdx.x] = x + y0 + y1 + y2 + y3;
}
int main() {
  double *out; long long *cyc, h[6];
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 6);
  for (int threads : {32, 128}) {
    k<<<1, threads>>>(out, cyc, 0.999, 1e-3); k<<<1, threads>>>(out, cyc, 0.999, 1e-3);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads=%d: DFMA dep %.1f cyc | DADD dep %.1f | SHFL64+DFMA %.1f | STS+LDS+DADD %.1f | BAR %.1f | 4xDFMA indep %.1f per 4\n", threads,
           h[0] / 256.0, h[1] / 256.0, h[2] / 128.0, h[3] / 128.0, h[4] / 128.0, h[5] / 256.0);
  }
  return 0;
}
