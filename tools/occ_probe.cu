// occupancy probe for the tensor-memory march kernel: what does the runtime think limits it?
#include <cstdio>
#include <cuda_runtime.h>
#include "../scft_b200/csrc/march1d_tmem.cuh"
using namespace scftb;
int main() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, (const void *)march_tm_kernel);
  printf("regs %d static smem %zu local %zu maxThreads %d maxDyn %d carveout %d\n", a.numRegs, a.sharedSizeBytes, a.localSizeBytes,
         a.maxThreadsPerBlock, a.maxDynamicSharedSizeBytes, a.preferredShmemCarveout);
  for (int dyn : {0, 8192, 30720}) {
    for (int carve : {-1, 100}) {
      cudaFuncSetAttribute((const void *)march_tm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
      cudaFuncSetAttribute((const void *)march_tm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      int occ = -1;
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)march_tm_kernel, 128, dyn);
      printf("dyn %d carve %d -> occ %d (%s)\n", dyn, carve, occ, cudaGetErrorString(e));
    }
  }
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("smem/SM %zu smem/block optin %zu regs/SM %d reserved smem/block %zu\n", pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin, pr.regsPerMultiprocessor, pr.reservedSharedMemPerBlock);
  return 0;
}
