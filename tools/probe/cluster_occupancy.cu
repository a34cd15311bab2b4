// Development probe: how many thread-block clusters of a given size / footprint are co-resident on this device.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void dummy(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; if (p) p[0] = s[0]; }
int main() {
  int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  printf("%s SMs=%d smem/SM=%zu smem/block optin=%zu regs/SM=%d\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin, pr.regsPerMultiprocessor);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int csz : {8, 12, 16}) for (int smem_kb : {36, 88, 118, 160, 200, 220}) for (int threads : {256, 416}) {
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(csz, 16); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem_kb * 1024;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, dummy, &cfg);
    printf("cluster=%2d smem=%3d KB threads=%d -> max active clusters %d (%s)\n", csz, smem_kb, threads, nc, cudaGetErrorString(e));
  }
  return 0;
}
