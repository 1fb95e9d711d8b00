// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe tools/cluster_probe.cu && ./cluster_probe
// B200 result (round 1): cluster 1: 148 CTAs, 2: 74 clusters (148), 4: 33 (132), 8: 15 (120), 16: 7 (112).
// How many clusters of size c (one 200 KB CTA per SM) can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int c : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148 / c * c); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = c; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster %2d: max active clusters %d (%d CTAs) %s\n", c, n, n * c, cudaGetErrorString(e));
    }
    return 0;
}
