// microbench_gather_stream.cu -- does a coalesced matrix stream cost the random x gather anything?  (developer tool, not product)
// Every thread does UNROLL random 8-byte gathers from an n-element vector per iteration and, next to them, reads S bytes of a
// coalesced stream per gather (8-byte values, 4-byte ids: what SELL / ELL kernels do).  Also: the same gathers with SHARE lanes
// falling into one 32-byte sector (is the L1-miss limit per lane, per sector or per request?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mbgs tools/microbench_gather_stream.cu && tools/mbgs
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t a) { a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16; return a; }

// STREAM: 0 none, 1 values only (8 B per gather), 2 values + ids (12 B per gather).  SHARE: lanes per 32-byte sector (1, 2, 4).
template <int STREAM, int SHARE, int UNROLL>
__global__ void __launch_bounds__(256) k(const double* __restrict__ x, const double* __restrict__ sv, const uint32_t* __restrict__ sj,
                                         uint32_t mask, int iters, size_t stream_elems, double* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31;
    double acc = 0;
    uint32_t h = hash32((SHARE > 1 ? tid / SHARE : tid) * 2654435761u + 12345u);
    size_t pos = tid;
    for (int it = 0; it < iters; ++it) {
        double v[UNROLL], a[UNROLL];
        uint32_t c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            a[u] = 1.0;
            c[u] = 0;
            if (STREAM >= 1) a[u] = __ldcs(sv + pos);
            if (STREAM >= 2) c[u] = __ldcs(sj + pos);
            pos += nthreads;
            if (pos >= stream_elems) pos -= stream_elems;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            h = h * 1664525u + 1013904223u;
            uint32_t idx = hash32(h) & mask;
            if (SHARE > 1) idx = (idx & ~(uint32_t) (SHARE - 1)) | (lane & (SHARE - 1));  // SHARE adjacent lanes, one sector (SHARE <= 4)
            v[u] = __ldg(x + ((idx + c[u]) & mask));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += a[u] * v[u];
    }
    if (acc == 1.2345e-300) out[tid] = acc;
}

template <int STREAM, int SHARE>
void run(const char* name, const double* x, const double* sv, const uint32_t* sj, uint32_t n, size_t stream_elems) {
    constexpr int UNROLL = 8;
    auto kk = k<STREAM, SHARE, UNROLL>;
    double* out; CK(cudaMalloc(&out, 8));
    const int grid = 148 * 8, iters = 128;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kk<<<grid, 256>>>(x, sv, sj, n - 1, 8, stream_elems, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kk<<<grid, 256>>>(x, sv, sj, n - 1, iters, stream_elems, out);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = best < ms ? best : ms;
    }
    const double g = (double) grid * 256 * iters * UNROLL;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double sb = STREAM == 0 ? 0 : STREAM == 1 ? 8 : 12;
    printf("%-52s n=%9u (%6.1f MB): %7.1f Ggather/s  %.3f gathers/clk/SM  stream %.0f B/gather = %6.0f GB/s  | 69.2 M gathers (cfg5, K=32 p=.15) would take %.0f us\n",
           name, n, n * 8.0 / 1e6, g / best / 1e6, g / (best * 1e-3) / 148 / (clk * 1e3), sb, g * sb / best / 1e6, 69.2e6 / (g / best / 1e3));
    fflush(stdout);
    cudaFree(out);
}

int main() {
    const uint32_t NMAX = 1u << 25;
    const size_t SE = 1ull << 28;  // 2 GB of values + 1 GB of ids: never L2-resident
    double *x, *sv; uint32_t* sj;
    CK(cudaMalloc(&x, (size_t) NMAX * 8)); CK(cudaMemset(x, 0, (size_t) NMAX * 8));
    CK(cudaMalloc(&sv, SE * 8)); CK(cudaMemset(sv, 0, SE * 8));
    CK(cudaMalloc(&sj, SE * 4)); CK(cudaMemset(sj, 0, SE * 4));
    for (uint32_t n : {1u << 22, 1u << 23}) {
        run<0, 1>("gathers only", x, sv, sj, n, SE);
        run<1, 1>("gathers + 8 B of coalesced values each", x, sv, sj, n, SE);
        run<2, 1>("gathers + 12 B of coalesced values and ids each", x, sv, sj, n, SE);
        run<0, 2>("gathers only, 2 lanes per 32-byte sector", x, sv, sj, n, SE);
        run<0, 4>("gathers only, 4 lanes per 32-byte sector", x, sv, sj, n, SE);
        run<2, 4>("gathers (4 lanes per sector) + 12 B stream", x, sv, sj, n, SE);
        printf("\n");
    }
    return 0;
}
