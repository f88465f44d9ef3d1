// tmem_bw.cu — can Blackwell's tensor memory (256 KB per SM, otherwise idle in an fp64 kernel) serve as a register-file
// extension for the march kernel's per-thread coefficients?  Measures, on sm_100a:
//   (1) tcgen05.ld throughput (32x32b.x16, 4 loads in flight per wait) in bytes per SM clock, for 1..4 CTAs of 128
//       threads per SM (each CTA allocates 128 columns = 64 fp64 per thread);
//   (2) tcgen05.st throughput, same shape;
//   (3) round-trip latency of a dependent ld -> use -> ld chain (x2 loads) and of ld + st (accumulator update);
//   (4) the same loads while the warp also runs a dependent DFMA chain (does TMEM traffic disturb the fp64 pipe?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCOLS 128

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst) {
  unsigned a = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t &a, uint32_t &b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: ld throughput, 1: st throughput, 2: ld latency chain, 3: ld+st accumulator chain, 4: ld throughput + DFMA chain
__global__ void __launch_bounds__(128) tmem_kernel(int mode, int iters, long long *cyc, unsigned *sink, int *smid) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&s_base);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base + ((uint32_t)(warp * 32) << 16);
  uint32_t r[4][16];
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int i = 0; i < 16; i++) r[b][i] = threadIdx.x * 64 + b * 16 + i;
  // initialise all 128 columns of this warp's lanes
  for (int c = 0; c < NCOLS; c += 16) tmem_st16(base + c, r[(c / 16) & 3]);
  wait_st();
  unsigned acc = 0;
  double x = threadIdx.x * 1e-3;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0 || mode == 4) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int half = 0; half < 2; half++) {
#pragma unroll
        for (int b = 0; b < 4; b++) tmem_ld16(base + half * 64 + b * 16, r[b]);
        if (mode == 4) {
#pragma unroll
          for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(0.999), "d"(1e-3));
        }
        wait_ld();
#pragma unroll
        for (int b = 0; b < 4; b++) acc ^= r[b][0] ^ r[b][7] ^ r[b][15];
      }
    }
  } else if (mode == 1) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int half = 0; half < 2; half++) {
#pragma unroll
        for (int b = 0; b < 4; b++) { r[b][0] += it; tmem_st16(base + half * 64 + b * 16, r[b]); }
        wait_st();
      }
    }
  } else if (mode == 2) {
    uint32_t col = 0;
    for (int it = 0; it < iters; it++) {
      uint32_t a, b;
      tmem_ld2(base + col, a, b);
      wait_ld();
      col = (a + b + it) & 0x7e;   // next address depends on the loaded value
      acc += a;
    }
  } else if (mode == 3) {
    for (int it = 0; it < iters; it++) {
      uint32_t a, b;
      tmem_ld2(base + 2 * (it & 31), a, b);
      wait_ld();
      double v = __hiloint2double(b, a);
      v = fma(v, 0.999, 1e-3);
      tmem_st2(base + 2 * (it & 31), __double2loint(v), __double2hiint(v));
      wait_st();
      acc += a;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) {
    cyc[blockIdx.x] = t1 - t0;
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    smid[blockIdx.x] = (int)s;
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + (unsigned)x;
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_base);
}

int main() {
  const int SMS = 148, MAXB = SMS * 4;
  long long *d_cyc, h_cyc[MAXB];
  unsigned *d_sink;
  int *d_smid, h_smid[MAXB];
  cudaMalloc(&d_cyc, sizeof(long long) * MAXB);
  cudaMalloc(&d_sink, sizeof(unsigned) * MAXB * 128);
  cudaMalloc(&d_smid, sizeof(int) * MAXB);
  const char *names[] = {"ld x16 (4 in flight)", "st x16 (4 in flight)", "ld x2 dependent chain", "ld x2 + DFMA + st x2 chain",
                         "ld x16 (4 in flight) + 16 dependent DFMA per batch"};
  for (int mode = 0; mode < 5; mode++) {
    for (int k = 1; k <= 4; k++) {
      if ((mode == 2 || mode == 3) && k > 1) continue;
      const int grid = (mode == 2 || mode == 3) ? 1 : SMS * k, iters = 2000;
      for (int rep = 0; rep < 2; rep++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        tmem_kernel<<<grid, 128>>>(mode, iters, d_cyc, d_sink, d_smid);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(err)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 0) continue;
        cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        cudaMemcpy(h_smid, d_smid, sizeof(int) * grid, cudaMemcpyDeviceToHost);
        long long mx = 0; double mean = 0;
        int per_sm[256] = {0}, maxper = 0;
        for (int b = 0; b < grid; b++) { mx = h_cyc[b] > mx ? h_cyc[b] : mx; mean += (double)h_cyc[b] / grid; per_sm[h_smid[b] & 255]++; }
        for (int s = 0; s < 256; s++) maxper = per_sm[s] > maxper ? per_sm[s] : maxper;
        if (mode == 2 || mode == 3) printf("%-52s : %.1f cycles per iteration\n", names[mode], mean / iters);
        else {
          const double bytes_per_cta = (double)iters * NCOLS * 4 * 128;   // every warp moves all 128 columns of its 32 lanes
          printf("%-52s : %d CTA/SM (max %d on one SM): %.1f B/clk/SM (mean CTA %.0f cyc, max %.0f), kernel %.3f ms -> %.1f TB/s chip\n",
                 names[mode], k, maxper, bytes_per_cta * k / mean, mean, (double)mx, ms, bytes_per_cta * grid / (ms * 1e-3) / 1e12);
        }
      }
    }
  }
  return 0;
}
