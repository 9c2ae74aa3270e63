// microbench_gather.cu -- what bounds the x-gather of SpMV on B200?  (developer tool, not product)
// Random 8-byte gathers through different paths; indices are hashed on the fly (no index traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbg tools/microbench_gather.cu && /tmp/mbg
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t a) {
    a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16; return a;
}
__device__ __forceinline__ double ld_na(const double* p) {
    double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }

// MODE 0: __ldg   1: L1::no_allocate   2: ld.cg   3: tex1Dfetch<int2>   4: LDS (window in smem)  5: ldg, lanes pair-adjacent (2 per line)
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ x, cudaTextureObject_t tex, uint32_t mask, int iters, double* out) {
    extern __shared__ double sx[];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 4) {
        for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) sx[i] = x[i];
        __syncthreads();
    }
    double acc = 0;
    uint32_t h = hash32(tid * 2654435761u + 12345u);
    for (int it = 0; it < iters; ++it) {
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            h = h * 1664525u + 1013904223u;
            uint32_t idx = hash32(h) & mask;
            if (MODE == 5) idx = (idx & ~15u) | ((threadIdx.x & 15u));  // 16 lanes share one 128-B line
            if (MODE == 0 || MODE == 5) v[u] = __ldg(x + idx);
            else if (MODE == 1) v[u] = ld_na(x + idx);
            else if (MODE == 2) v[u] = ld_cg(x + idx);
            else if (MODE == 3) { int2 t = tex1Dfetch<int2>(tex, (int) idx); v[u] = __hiloint2double(t.y, t.x); }
            else v[u] = sx[idx];
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u];
    }
    if (acc == 1.2345e-300) out[tid] = acc;
}

template <int MODE>
void run(const char* name, const double* x, cudaTextureObject_t tex, uint32_t n, int carve, size_t smem, int blocks_per_sm) {
    constexpr int UNROLL = 8;
    auto k = gather_kernel<MODE, UNROLL>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    if (smem) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    double* out; CK(cudaMalloc(&out, 8));
    const int grid = 148 * blocks_per_sm, iters = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<grid, 256, smem>>>(x, tex, n - 1, 8, out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<<<grid, 256, smem>>>(x, tex, n - 1, iters, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double) grid * 256 * iters * UNROLL;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-44s n=%9u (%7.1f MB) carve=%3d blk/SM=%d : %8.1f Ggather/s  %.3f gathers/clk/SM (at %.0f MHz)  => SpMV-equivalent %.0f GB/s\n",
           name, n, n * 8.0 / 1e6, carve, blocks_per_sm, g / ms / 1e6, g / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1e3, g / ms / 1e6 * 12);
    cudaFree(out);
}

int main() {
    const uint32_t NMAX = 1u << 25;  // 256 MB
    double* x; CK(cudaMalloc(&x, (size_t) NMAX * 8)); CK(cudaMemset(x, 0, (size_t) NMAX * 8));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x;
    rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = (size_t) NMAX * 8;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    for (uint32_t n : {1u << 12, 1u << 16, 1u << 22, 1u << 25}) {
        run<0>("ldg  (L1 allocate)", x, tex, n, 0, 0, 8);
        run<0>("ldg  (L1 allocate), max-smem carveout", x, tex, n, 100, 0, 8);
        run<1>("ld.global.nc.L1::no_allocate", x, tex, n, 0, 0, 8);
        run<1>("ld.global.nc.L1::no_allocate, max-smem carve", x, tex, n, 100, 0, 8);
        run<2>("ld.global.cg", x, tex, n, 0, 0, 8);
        run<3>("tex1Dfetch<int2>", x, tex, n, 0, 0, 8);
        run<5>("ldg, 16 lanes per 128-B line", x, tex, n, 0, 0, 8);
        printf("\n");
    }
    run<4>("LDS random (32 KB window in smem)", x, tex, 1u << 12, 100, (1u << 12) * 8, 4);
    run<4>("LDS random (128 KB window in smem)", x, tex, 1u << 14, 100, (1u << 14) * 8, 1);
    return 0;
}
