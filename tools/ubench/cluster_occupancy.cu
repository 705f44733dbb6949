// How many thread-block clusters of a given size (with the shared memory / thread count of the GEMM kernels) can be
// co-resident on this GPU: nvcc -arch=sm_100a -o cluster_occupancy cluster_occupancy.cu && ./cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 227 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    for (int threads : {320, 576}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
      printf("cluster size %2d, %d threads, %d KB smem: max active clusters %d (%d SMs)%s\n", cs, threads, smem / 1024, n, n * cs,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
