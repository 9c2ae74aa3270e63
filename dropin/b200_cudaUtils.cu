// dropin/b200_cudaUtils.cu -- drop-in for src/commons/cudaUtils.cu: the H2D helpers declared in
// src/include/cudaUtils.h:60-68 (spMatCpyCSR, spMatCpyELL, spMatCpyELLNNPitched), same signatures, same
// ownership rules (the caller cudaMalloc's the device struct and later frees the members through a host copy
// with cudaFreeSpmat, src/main.cu:195,215,280), same return convention (EXIT_SUCCESS / EXIT_FAILURE with the
// members freed), producing the REFERENCE's device layout so that the reference kernels or
// dropin/b200_SpMV_CUDA.cu can consume it.  Compile with the reference's headers on the include path.
// Defects of the reference that are not reproduced: element sizes come from the pointee
// types, RL is allocated and copied with its real length (for a transposed ELL struct that is MAX_ROW_NZ rows).
#include "cudaUtils.h"

namespace {
struct Upload {  // allocations of one upload; released together on failure
    void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    bool ok = true;
    const char* tag;
    explicit Upload(const char* t) : tag(t) {}
    template <typename T>
    T* flat(const T* host, size_t count) {
        T* d = nullptr;
        if (!ok) return d;
        if (cudaErr(cudaMalloc(&d, sizeof(T) * (count ? count : 1)), tag)) { ok = false; return nullptr; }
        ptr[n++] = d;
        if (count && cudaErr(cudaMemcpy(d, host, sizeof(T) * count, dirUp), tag)) ok = false;
        return d;
    }
    template <typename T>
    T* pitched(const T* host, size_t width, size_t height, size_t* pitch_elems) {
        T* d = nullptr;
        size_t pitch_bytes = 0;
        if (!ok) return d;
        if (cudaErr(cudaMallocPitch(&d, &pitch_bytes, sizeof(T) * (width ? width : 1), height ? height : 1), tag)) { ok = false; return nullptr; }
        ptr[n++] = d;
        if (width && height && cudaErr(cudaMemcpy2D(d, pitch_bytes, host, sizeof(T) * width, sizeof(T) * width, height, dirUp), tag)) ok = false;
        *pitch_elems = pitch_bytes / sizeof(T);
        return d;
    }
    int finish(spmat* dst, const spmat& local) {
        if (ok && cudaErr(cudaMemcpy(dst, &local, sizeof(local), dirUp), tag)) ok = false;
        if (!ok)
            for (int i = 0; i < n; ++i) cudaFree(ptr[i]);
        return ok ? EXIT_SUCCESS : EXIT_FAILURE;
    }
};

spmat header_of(const spmat* m) {
    spmat d;
    memset(&d, 0, sizeof(d));
    d.M = m->M;
    d.N = m->N;
    d.NZ = m->NZ;
    d.MAX_ROW_NZ = m->MAX_ROW_NZ;
    return d;
}
#ifdef ROWLENS
// ellTranspose swaps M <-> MAX_ROW_NZ and sets N = original M (src/commons/sparseUtils.c:168-171): the
// row-length vector then has MAX_ROW_NZ entries, not M.
size_t rl_entries(const spmat* m) { return (m->N == m->MAX_ROW_NZ && m->IRP == NULL) ? m->MAX_ROW_NZ : m->M; }
#endif
}  // namespace

int spMatCpyCSR(spmat* m, spmat* dst) {
    Upload up("spMatCpyCSR");
    spmat d = header_of(m);
    d.JA = up.flat(m->JA, m->NZ);
    d.AS = up.flat(m->AS, m->NZ);
    d.IRP = up.flat(m->IRP, m->M + 1);
#ifdef ROWLENS
    d.RL = up.flat(m->RL, m->M);
#endif
    return up.finish(dst, d);
}

int spMatCpyELL(spmat* m, spmat* dst) {
    Upload up("spMatCpyELL");
    spmat d = header_of(m);
    // @m is row-major M x MAX_ROW_NZ (after ellTranspose: K x rows) -> pitched 2-D allocations, pitch in elements
    d.JA = up.pitched(m->JA, m->MAX_ROW_NZ, m->M, &d.pitchJA);
    d.AS = up.pitched(m->AS, m->MAX_ROW_NZ, m->M, &d.pitchAS);
#ifdef ROWLENS
    d.RL = up.flat(m->RL, rl_entries(m));
#endif
    return up.finish(dst, d);
}

int spMatCpyELLNNPitched(spmat* m, spmat* dst) {
    Upload up("spMatCpyELLNNPitched");
    spmat d = header_of(m);
    d.JA = up.flat(m->JA, m->MAX_ROW_NZ * m->M);
    d.AS = up.flat(m->AS, m->MAX_ROW_NZ * m->M);
    d.pitchJA = d.pitchAS = m->MAX_ROW_NZ;
#ifdef ROWLENS
    d.RL = up.flat(m->RL, rl_entries(m));
#endif
    return up.finish(dst, d);
}
