// hostmem_lab.cu -- duplex PCIe copy rate of 268 MB vectors by kind of host memory: cudaHostAlloc, malloc + cudaHostRegister (pages
// touched before / not), mmap + MADV_HUGEPAGE + cudaHostRegister, 2 MB-aligned malloc + madvise.  nvcc -O2 -arch=sm_100a
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static void run(const char* tag, double* hx, double* hy, double* dx, double* dy, size_t n) {
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    float up = 1e9f, down = 1e9f, both = 1e9f;
    for (int rep = 0; rep < 8; ++rep) {
        float a, b;
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(dx, hx, n * 8, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&a, e0, e1)); up = std::min(up, a);
        CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(hy, dy, n * 8, cudaMemcpyDeviceToHost, s1)); CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&a, e0, e1)); down = std::min(down, a);
        CK(cudaEventRecord(e0, s1)); CK(cudaStreamWaitEvent(s2, e0, 0));
        CK(cudaMemcpyAsync(dx, hx, n * 8, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(hy, dy, n * 8, cudaMemcpyDeviceToHost, s2));
        CK(cudaEventRecord(e1, s1)); CK(cudaEventRecord(e2, s2)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e0, e2)); both = std::min(both, std::max(a, b));
    }
    printf("%-46s up %.3f ms  down %.3f ms  duplex %.3f ms\n", tag, up, down, both);
    fflush(stdout);
    cudaStreamDestroy(s1); cudaStreamDestroy(s2);
}

int main() {
    const size_t n = 1ull << 25, bytes = n * 8;
    double *dx, *dy;
    CK(cudaMalloc(&dx, bytes)); CK(cudaMalloc(&dy, bytes));
    if (FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r")) { char b[128] = {0}; if (fgets(b, 127, f)) printf("THP enabled: %s", b); fclose(f); }
    {
        double *hx, *hy;
        CK(cudaHostAlloc(&hx, bytes, cudaHostAllocDefault)); CK(cudaHostAlloc(&hy, bytes, cudaHostAllocDefault));
        memset(hx, 1, bytes);
        run("cudaHostAlloc", hx, hy, dx, dy, n);
        cudaFreeHost(hx); cudaFreeHost(hy);
    }
    {
        double *hx = (double*) malloc(bytes), *hy = (double*) malloc(bytes);
        memset(hx, 1, bytes);  // x touched (a driver fills it), y untouched
        CK(cudaHostRegister(hx, bytes, cudaHostRegisterDefault)); CK(cudaHostRegister(hy, bytes, cudaHostRegisterDefault));
        run("malloc + register (x touched, y not)", hx, hy, dx, dy, n);
        cudaHostUnregister(hx); cudaHostUnregister(hy); free(hx); free(hy);
    }
    {
        double *hx = (double*) malloc(bytes), *hy = (double*) malloc(bytes);
        memset(hx, 1, bytes); memset(hy, 1, bytes);
        CK(cudaHostRegister(hx, bytes, cudaHostRegisterDefault)); CK(cudaHostRegister(hy, bytes, cudaHostRegisterDefault));
        run("malloc + register (both touched)", hx, hy, dx, dy, n);
        cudaHostUnregister(hx); cudaHostUnregister(hy); free(hx); free(hy);
    }
    {
        double* hx = (double*) mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        double* hy = (double*) mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        int r1 = madvise(hx, bytes, MADV_HUGEPAGE), r2 = madvise(hy, bytes, MADV_HUGEPAGE);
        memset(hx, 1, bytes); memset(hy, 1, bytes);
        CK(cudaHostRegister(hx, bytes, cudaHostRegisterDefault)); CK(cudaHostRegister(hy, bytes, cudaHostRegisterDefault));
        char tag[96]; snprintf(tag, 96, "mmap + MADV_HUGEPAGE(%d,%d) + touch + register", r1, r2);
        run(tag, hx, hy, dx, dy, n);
        cudaHostUnregister(hx); cudaHostUnregister(hy); munmap(hx, bytes); munmap(hy, bytes);
    }
    {
        double *hx = (double*) malloc(bytes), *hy = (double*) malloc(bytes);
        memset(hx, 1, bytes); memset(hy, 1, bytes);  // touched as 4 KB pages first, THEN advised (what a library can do to a caller's buffer)
        auto adv = [&](void* p) { uintptr_t a = ((uintptr_t) p + 4095) & ~(uintptr_t) 4095, e = ((uintptr_t) p + bytes) & ~(uintptr_t) 4095; return madvise((void*) a, e - a, MADV_HUGEPAGE); };
        int r1 = adv(hx), r2 = adv(hy);
        CK(cudaHostRegister(hx, bytes, cudaHostRegisterDefault)); CK(cudaHostRegister(hy, bytes, cudaHostRegisterDefault));
        char tag[96]; snprintf(tag, 96, "malloc + touch + late MADV_HUGEPAGE(%d,%d) + reg", r1, r2);
        run(tag, hx, hy, dx, dy, n);
        cudaHostUnregister(hx); cudaHostUnregister(hy); free(hx); free(hy);
    }
    {
        double *hx = (double*) malloc(bytes), *hy = (double*) malloc(bytes);
        memset(hx, 1, bytes);
        run("malloc, not registered (pageable)", hx, hy, dx, dy, n);
        free(hx); free(hy);
    }
    return 0;
}
