// Can a kernel be launched cooperatively (co-residency guaranteed) AND with a thread-block cluster dimension?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o coop_cluster_probe coop_cluster_probe.cu && ./coop_cluster_probe
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void probe(int* out, unsigned* counter, int rounds) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ int box;
    if (threadIdx.x == 0) box = 0;
    cl.sync();
    // every CTA adds its rank into the leader's shared memory (DSMEM), then a software barrier over the grid
    int* leader = cl.map_shared_rank(&box, 0);
    if (threadIdx.x == 0) atomicAdd(leader, (int)cl.block_rank() + 1);
    cl.sync();
    for (int r = 0; r < rounds; ++r) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            while (atomicAdd(counter, 0u) < (unsigned)gridDim.x * (r + 1)) { }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && cl.block_rank() == 0) out[blockIdx.x / cl.dim_blocks().x] = box;
}

int main() {
    int* out; unsigned* counter;
    const int cluster = 4, grid = 148 * 2;      // 296 CTAs = 74 clusters of 4
    cudaMalloc(&out, sizeof(int) * grid); cudaMalloc(&counter, 4); cudaMemset(counter, 0, 4);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    for (int nattr = 2; nattr >= 1; --nattr) {
        cfg.attrs = at; cfg.numAttrs = nattr;
        cudaMemset(counter, 0, 4);
        int clusters = 0;
        cudaError_t eo = cudaOccupancyMaxActiveClusters(&clusters, probe, &cfg);
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe, out, counter, 50);
        cudaError_t s = cudaDeviceSynchronize();
        int h[4] = {0};
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("attrs=%d (cluster%s): maxActiveClusters=%d (%s) launch=%s sync=%s box[0..3]=%d %d %d %d (want %d)\n", nattr,
               nattr == 2 ? "+cooperative" : " only", clusters, cudaGetErrorName(eo), cudaGetErrorName(e), cudaGetErrorName(s),
               h[0], h[1], h[2], h[3], cluster * (cluster + 1) / 2);
        cudaGetLastError();
    }
    return 0;
}
