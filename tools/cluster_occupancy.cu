// How many thread-block clusters of a GEMM-sized CTA (384 threads, ~200 KB of shared memory, one CTA per SM) can be
// co-resident on this GPU, per cluster size: decides whether a 2x2 cluster of CTA pairs (4 CTAs) is worth building.
// nvcc -gencode arch=compute_100a,code=sm_100a tools/cluster_occupancy.cu -o build/cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dummy(int* p) {
  extern __shared__ int s[];
  if (p) p[0] = s[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = cs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_dummy, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d SMs busy (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
