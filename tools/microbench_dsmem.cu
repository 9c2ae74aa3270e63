// microbench_dsmem.cu -- random 8-byte gathers from DISTRIBUTED shared memory (thread-block cluster) on B200.
// Question: can a cluster-resident x window beat the L2 sector rate (~1 gather/clk/SM) for wide banded matrices?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mbd tools/microbench_dsmem.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t a) { a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16; return a; }

// window of WIN doubles per CTA; the cluster's window is CSIZE*WIN doubles; element e lives in CTA (e / WIN), offset e % WIN
template <int UNROLL>
__global__ void __launch_bounds__(1024) dsmem_gather(int iters, uint32_t win, uint32_t local_only, double* out) {
    extern __shared__ double sx[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t csize = cluster.num_blocks(), crank = cluster.block_rank();
    for (uint32_t i = threadIdx.x; i < win; i += blockDim.x) sx[i] = 1.0;
    cluster.sync();
    const uint32_t total = local_only ? win : win * csize;
    double acc = 0;
    uint32_t h = hash32((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 7u);
    for (int it = 0; it < iters; ++it) {
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            h = h * 1664525u + 1013904223u;
            const uint32_t e = hash32(h) % total;
            const uint32_t owner = local_only ? crank : e / win, off = e % win;
            const double* p = cluster.map_shared_rank(sx, owner);
            v[u] = p[off];
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u];
    }
    cluster.sync();
    if (acc == 1.2345e-300) out[0] = acc;
}

void run(int csize, uint32_t win, int local_only) {
    auto k = dsmem_gather<8>;
    const size_t smem = (size_t) win * 8;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    double* out; CK(cudaMalloc(&out, 8));
    cudaLaunchConfig_t cfg = {};
    const int nclusters = 148 / csize;
    cfg.gridDim = dim3(nclusters * csize); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int iters = 8;
    CK(cudaLaunchKernelEx(&cfg, k, iters, win, (uint32_t) local_only, out));
    CK(cudaDeviceSynchronize());
    iters = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, k, iters, win, (uint32_t) local_only, out));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double) cfg.gridDim.x * 1024 * iters * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("cluster=%d  window/CTA=%6.0f KB (cluster %7.0f KB) %s: %8.1f Ggather/s  %.3f gathers/clk/SM (%d CTAs)\n", csize, win * 8 / 1024.0,
           win * 8.0 * csize / 1024.0, local_only ? "LOCAL smem only " : "cluster-wide DSMEM", g / ms / 1e6, g / (ms * 1e-3) / cfg.gridDim.x / (clk * 1e3), cfg.gridDim.x);
    cudaFree(out);
}

int main() {
    for (int cs : {1, 2, 4, 8}) {
        run(cs, 20480, 1);   // 160 KB per CTA, local only (baseline)
        if (cs > 1) run(cs, 20480, 0);
    }
    return 0;
}
