// Prints cudaOccupancyMaxActiveClusters for a few (cluster size, dynamic smem, threads) combinations.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int smems[] = {48 * 1024, 100 * 1024, 180 * 1024, 222464};
  const int cs[] = {1, 2, 4, 8, 16};
  for (int c : cs)
    for (int sm : smems)
      for (int th : {320, 576}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(c * 64);
        cfg.blockDim = dim3(th);
        cfg.dynamicSmemBytes = sm;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension;
        a[0].val.clusterDim.x = c; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster %2d smem %6d threads %d -> max active clusters %d (%s)\n", c, sm, th, n, cudaGetErrorString(e));
      }
  return 0;
}
