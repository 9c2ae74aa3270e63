// pipe_lab.cu -- stand-alone model of the host-buffer pipeline (x pieces up -> kernel chunk -> y chunk down) with every ingredient
// switchable, to find what keeps spmvb200_spmv_host below the bare duplex copy rate.  nvcc -O2 -arch=sm_100a -o pipe_lab pipe_lab.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void hog(const double4* a, double4* b, size_t n) {  // HBM-streaming stand-in for a kernel chunk
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void touch(const double* x, double* y, size_t r0, size_t r1) {  // cheap kernel: y = x on the chunk
    for (size_t i = r0 + blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < r1; i += (size_t) gridDim.x * blockDim.x) y[i] = x[i];
}

struct Cfg { int chunks; bool ramp; int mid; /*0 none,1 events only (timing),2 events only (no timing),3 touch kernel,4 hog kernel*/ bool stagger; int grid = 1184; };

int main() {
    const size_t n = 1ull << 25;
    double *hx, *hy, *dx, *dy;
    double4 *ga, *gb;
    const size_t hog_n = (1ull << 30) / 32 * 6;  // 6 GB read + 6 GB written over 16 chunks ~ the x-window kernel's traffic
    CK(cudaHostAlloc(&hx, n * 8, cudaHostAllocDefault));
    CK(cudaHostAlloc(&hy, n * 8, cudaHostAllocDefault));
    CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&dy, n * 8));
    CK(cudaMalloc(&ga, hog_n * 32)); CK(cudaMalloc(&gb, hog_n * 32));
    for (size_t i = 0; i < n; ++i) hx[i] = (double) i;
    cudaStream_t su, sc, sd;
    CK(cudaStreamCreateWithFlags(&su, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    std::vector<Cfg> cfgs = {
        {1, false, 0, false}, {16, false, 0, false}, {16, false, 0, true}, {16, false, 1, true}, {16, false, 2, true}, {16, false, 3, true}, {16, false, 4, true},
        {16, true, 0, true}, {16, true, 2, true}, {16, true, 3, true}, {16, true, 4, true}, {12, true, 4, true}, {8, false, 4, true}, {32, false, 4, true},
        {16, true, 4, true, 592}, {16, true, 4, true, 296}, {16, true, 4, true, 148}, {16, true, 4, true, 74}, {16, true, 4, true, 37}};
    for (const Cfg& c : cfgs) {
        std::vector<double> w;
        if (c.ramp && c.chunks >= 10) { for (int k = 0; k < 4; ++k) w.push_back(1 << k); for (int k = 0; k < c.chunks - 8; ++k) w.push_back(16); for (int k = 3; k >= 0; --k) w.push_back(1 << k); }
        else w.assign(c.chunks, 1.0);
        double tot = 0; for (double v : w) tot += v;
        std::vector<size_t> b(w.size() + 1, 0);
        double cum = 0;
        for (size_t k = 0; k < w.size(); ++k) { cum += w[k]; b[k + 1] = std::min(n, ((size_t) (n * cum / tot) + 8191) & ~(size_t) 8191); }
        b[w.size()] = n;
        const int nch = (int) w.size();
        std::vector<cudaEvent_t> xr(nch), ks(nch), ke(nch);
        for (int k = 0; k < nch; ++k) {
            CK(cudaEventCreateWithFlags(&xr[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ks[k], c.mid == 1 ? cudaEventDefault : cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ke[k], c.mid == 1 ? cudaEventDefault : cudaEventDisableTiming));
        }
        std::vector<float> ev_ms, host_ms;
        for (int rep = 0; rep < 12; ++rep) {
            CK(cudaDeviceSynchronize());
            auto t0 = std::chrono::steady_clock::now();
            CK(cudaEventRecord(e0, su));
            CK(cudaStreamWaitEvent(sd, e0, 0));
            for (int k = 0; k < nch; ++k) {
                CK(cudaMemcpyAsync(dx + b[k], hx + b[k], (b[k + 1] - b[k]) * 8, cudaMemcpyHostToDevice, su));
                CK(cudaEventRecord(xr[k], su));
            }
            for (int k = 0; k < nch; ++k) {
                if (c.mid) {
                    CK(cudaStreamWaitEvent(sc, xr[k], 0));
                    CK(cudaEventRecord(ks[k], sc));
                    if (c.mid == 3) touch<<<592, 256, 0, sc>>>(dx, dy, b[k], b[k + 1]);
                    if (c.mid == 4) { const size_t h0 = (size_t) ((double) hog_n * b[k] / n), h1 = (size_t) ((double) hog_n * b[k + 1] / n); hog<<<c.grid, 512, 0, sc>>>(ga + h0, gb + h0, h1 - h0); }
                    CK(cudaEventRecord(ke[k], sc));
                    CK(cudaStreamWaitEvent(sd, ke[k], 0));
                } else if (c.stagger) {
                    CK(cudaStreamWaitEvent(sd, xr[k], 0));
                }
                CK(cudaMemcpyAsync(hy + b[k], dy + b[k], (b[k + 1] - b[k]) * 8, cudaMemcpyDeviceToHost, sd));
            }
            CK(cudaEventRecord(e1, su));
            CK(cudaEventRecord(e2, sd));
            CK(cudaStreamSynchronize(sc)); CK(cudaStreamSynchronize(sd)); CK(cudaStreamSynchronize(su));
            host_ms.push_back(std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count());
            float a, bb;
            CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&bb, e0, e2));
            ev_ms.push_back(std::max(a, bb));
        }
        std::sort(ev_ms.begin(), ev_ms.end()); std::sort(host_ms.begin(), host_ms.end());
        printf("chunks %2d %s mid=%d stagger=%d grid=%4d : events min %.3f med %.3f | host min %.3f med %.3f ms\n", nch, c.ramp ? "ramp   " : "uniform", c.mid, (int) c.stagger, c.grid,
               ev_ms[0], ev_ms[6], host_ms[0], host_ms[6]);
        fflush(stdout);
        for (int k = 0; k < nch; ++k) { cudaEventDestroy(xr[k]); cudaEventDestroy(ks[k]); cudaEventDestroy(ke[k]); }
    }
    return 0;
}
