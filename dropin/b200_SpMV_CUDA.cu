// dropin/b200_SpMV_CUDA.cu -- STRICT kernel-level drop-in for src/SpMV_CUDA.cu of SpMV_openMP_CUDA.
//
// Compile this file INSTEAD of the reference's src/SpMV_CUDA.cu, with the reference's own headers on the
// include path (nothing of the reference is copied here): it defines the five __global__ entry points
// declared in src/include/SpMV.h:119-128 with the exact signature
//     __global__ void NAME(spmat* m /*device struct*/, double* v, CONFIG cfg /*by value*/, double* outV)
// so the unchanged drivers (src/main.cu:233, test/SpMV_test.cu:112) and the tables SpmvCUDA_CSRFuncs /
// SpmvCUDA_ELLFuncs keep working, under WHATEVER launch geometry they pass: 1-D 256 x ceil(M/256), (32,32) x ceil(M/32) 1-D in x
// (the reference's warp kernels read blockIdx.y there and only ever compute rows 0..31), and the stale (32,32)
// shape the test harness leaves active for the 1-D ELL kernels.  Work is derived from a LINEAR thread id and
// grid-stride loops; a warp is 32 consecutive linear ids for every block shape the drivers use.
//
// Two tiers inside each kernel:
//   * FAST: when the struct was uploaded by dropin/b200_cudaUtils.cu it carries the narrow view of dropin/b200_view.h -- 32-bit ids,
//     SELL-32-sigma slices (CSR) or column-major storage with row lengths (ELL) -- and the kernel runs this engine's coalesced
//     algorithms on it: 12 B per non-zero instead of 16, every warp load a contiguous 256 B / 128 B run, row-length early exit.
//   * PLAIN: a struct uploaded by the reference's own spMatCpy* has no view; the kernel walks the 64-bit layout like the reference
//     does (bug-fixed: linear ids, all rows computed).
// Thread-per-row entry points add left to right with separate mul/add roundings => bit-identical to sgemvSerial
// (src/SpMV_CSR_OMP.c:229-250) for rows of at most 256 entries (longer CSR rows: a warp each, shuffle tree, within 1e-12 relative);
// warp entry points use an xor-shuffle tree (or the same serial-order kernel where that is the faster correct choice).
extern "C" {
#include "sparseMatrix.h"
#include "SpMV.h"
}
#include "cudaUtils.h"
#include "b200_view.h"

// Register budget: the drivers launch blocks of 256 and of 32 x 32 = 1024 threads.  B200_MIN_BLOCKS = 2 (two 1024-thread blocks per SM:
// 32 registers per thread, full occupancy; a few spills outside the hot loops) or 1 (64 registers, no spills).  These kernels live on
// loads in flight, so occupancy wins -- measured on B200 (profiles/r02d_dropin_regs.log): 27-point 128^3, CSR 0 129 vs 162 us, ELL 0 113 vs
// 161 us; 5-point 1024^2, CSR 0 26.9 vs 32.9 us.
#ifndef B200_MIN_BLOCKS
#define B200_MIN_BLOCKS 2
#endif
#define B200_BOUNDS __launch_bounds__(1024, B200_MIN_BLOCKS)

namespace {
__device__ __forceinline__ unsigned long long lin_tid() {
    const unsigned long long blk = blockIdx.x + (unsigned long long) gridDim.x * (blockIdx.y + (unsigned long long) gridDim.y * blockIdx.z);
    const unsigned tpb = blockDim.x * blockDim.y * blockDim.z;
    return blk * tpb + threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
}
__device__ __forceinline__ unsigned long long lin_threads() {
    return (unsigned long long) gridDim.x * gridDim.y * gridDim.z * (blockDim.x * blockDim.y * blockDim.z);
}
__device__ __forceinline__ bool whole_warps() { return ((blockDim.x * blockDim.y * blockDim.z) & 31u) == 0u; }
__device__ __forceinline__ double warp_tree(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 32);
    return v;
}
template <int LANES>
__device__ __forceinline__ double sub_tree(double v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 32);
    return v;
}
__device__ __forceinline__ const B200ViewCSR* csr_view(const spmat* m) {
    return m->pitchJA == (size_t) B200_VIEW_MAGIC ? reinterpret_cast<const B200ViewCSR*>(m->pitchAS) : nullptr;
}
__device__ __forceinline__ const B200ViewELL* ell_view(const spmat* m) {
    const B200ViewELL* v = reinterpret_cast<const B200ViewELL*>(m->IRP);
    return (v && v->magic == B200_VIEW_MAGIC) ? v : nullptr;
}

// SELL-32-sigma, one thread per sorted row, a warp per slice: 256 B of values + 128 B of ids per slot and warp, 4 slots in flight,
// loop bound = the slice's longest row (warp uniform), per-lane predicate at the row length; left-to-right sum.
__device__ __forceinline__ void sell_rows(const B200ViewCSR* __restrict__ vw, const double* __restrict__ x, double* __restrict__ y) {
    const unsigned* __restrict__ slice_ptr = vw->slice_ptr;
    const unsigned* __restrict__ perm = vw->perm;
    const unsigned* __restrict__ rl = vw->rl_sorted;
    const unsigned* __restrict__ sja = vw->sja;
    const double* __restrict__ sas = vw->sas;
    const unsigned Mpad = vw->Mpad;
    const unsigned stride = (unsigned) min(lin_threads(), 0x80000000ull);  // 32-bit index arithmetic (Mpad < 2^31): registers matter here
    if (lin_tid() >= Mpad) return;
    for (unsigned i = (unsigned) lin_tid(); i < Mpad; i += stride) {
        const unsigned len = __ldg(rl + i);
        const unsigned sp0 = __ldg(slice_ptr + (i >> 5)), sp1 = __ldg(slice_ptr + (i >> 5) + 1);
        const unsigned wmax = (sp1 - sp0) >> 5;
        const double* a = sas + sp0 + (i & 31);
        const unsigned* j = sja + sp0 + (i & 31);
        double acc = 0;
        unsigned k = 0;
        for (; k + 4 <= wmax; k += 4) {
            double v[4], xv[4];
            unsigned c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = k + u < len;
                v[u] = ok ? __ldcs(a + (k + u) * 32) : 0.0;
                c[u] = ok ? __ldcs(j + (k + u) * 32) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = (k + u < len) ? __ldg(x + c[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k + u < len) acc = __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
        }
        for (; k < wmax; ++k)
            if (k < len) acc = __dadd_rn(acc, __dmul_rn(__ldcs(a + k * 32), __ldg(x + __ldcs(j + k * 32))));
        const unsigned row = __ldg(perm + i);
        if (row != 0xffffffffu) y[row] = acc;
    }
}
// rows left out of the slices (longer than B200_LONG_ROW): a warp each, 128-bit value loads, shuffle tree
__device__ __forceinline__ void long_rows(const B200ViewCSR* __restrict__ vw, const double* __restrict__ as, const double* __restrict__ x,
                                          double* __restrict__ y) {
    const unsigned lane = (unsigned) (lin_tid() & 31);
    for (unsigned long long w = lin_tid() >> 5; w < vw->nlong; w += lin_threads() >> 5) {
        const unsigned row = vw->long_rows[w];
        const unsigned s = vw->irp32[row], e = vw->irp32[row + 1];
        double t = 0;
        for (unsigned i = (s & ~1u) + 2 * lane; i < e; i += 64) {
            const double2 v = __ldcs(reinterpret_cast<const double2*>(as + i));
            const uint2 c = __ldcs(reinterpret_cast<const uint2*>(vw->ja32 + i));
            if (i >= s) t = fma(v.x, __ldg(x + c.x), t);
            if (i + 1 < e) t = fma(v.y, __ldg(x + c.y), t);
        }
        t = warp_tree(t);
        if (lane == 0) y[row] = t;
    }
}
// narrow CSR, LANES lanes per row (2 non-zeros per lane and step), rows up to B200_LONG_ROW entries
template <int LANES>
__device__ __forceinline__ void vector_rows(const B200ViewCSR* __restrict__ vw, unsigned M, const double* __restrict__ as, const double* __restrict__ x,
                                            double* __restrict__ y) {
    const unsigned* __restrict__ irp = vw->irp32;
    const unsigned* __restrict__ ja = vw->ja32;
    constexpr unsigned RPW = 32 / LANES;  // rows per warp
    const unsigned lane = (unsigned) (lin_tid() % LANES), sub = (unsigned) ((lin_tid() & 31) / LANES);
    const unsigned long long nwarps = lin_threads() >> 5;
    // warp-uniform loop bound: a warp whose rows all lie beyond the matrix leaves at once (the (32,32) x ceil(M/32) geometry launches
    // one WARP per row, 32 / LANES times more sub-warps than rows)
    for (unsigned long long row0 = (lin_tid() >> 5) * RPW; row0 < M; row0 += nwarps * RPW) {
        const unsigned long long row = row0 + sub;
        double acc = 0;
        bool mine = row < M;
        if (mine) {
            const unsigned s = __ldg(irp + row), e = __ldg(irp + row + 1);
            mine = e - s <= B200_LONG_ROW;
            if (mine)
                for (unsigned i = (s & ~1u) + 2 * lane; i < e; i += 2 * LANES) {
                    const double2 v = __ldcs(reinterpret_cast<const double2*>(as + i));
                    const uint2 c = __ldcs(reinterpret_cast<const uint2*>(ja + i));
                    if (i >= s) acc = fma(v.x, __ldg(x + c.x), acc);
                    if (i + 1 < e) acc = fma(v.y, __ldg(x + c.y), acc);
                }
        }
        acc = sub_tree<LANES>(acc);
        if (lane == 0 && mine) y[row] = acc;
    }
}
// column-major narrow ELL, one thread per row, early exit at the warp's longest row, left-to-right sum
__device__ __forceinline__ void ell_cm_rows(const B200ViewELL* __restrict__ vw, const double* __restrict__ as, size_t as_pitch,
                                            const double* __restrict__ x, double* __restrict__ y) {
    const unsigned rows = vw->rows;
    const unsigned* __restrict__ ja = vw->ja32;
    const unsigned* __restrict__ rl = vw->rl32;
    const size_t jp = vw->pitch;
    const unsigned long long rround = ((unsigned long long) rows + 31) / 32 * 32;
    for (unsigned long long row = lin_tid(); row < rround; row += lin_threads()) {
        const bool live = row < rows;
        const unsigned len = live ? __ldg(rl + row) : 0u;
        const unsigned wmax = __reduce_max_sync(0xffffffffu, len);
        const double* a = as + row;
        const unsigned* j = ja + row;
        double acc = 0;
        unsigned k = 0;
        for (; k + 4 <= wmax; k += 4) {
            double v[4], xv[4];
            unsigned c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = k + u < len;
                v[u] = ok ? __ldcs(a + (size_t) (k + u) * as_pitch) : 0.0;
                c[u] = ok ? __ldcs(j + (size_t) (k + u) * jp) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = (k + u < len) ? __ldg(x + c[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k + u < len) acc = __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
        }
        for (; k < wmax; ++k)
            if (k < len) acc = __dadd_rn(acc, __dmul_rn(__ldcs(a + (size_t) k * as_pitch), __ldg(x + __ldcs(j + (size_t) k * jp))));
        if (live) y[row] = acc;
    }
}
}  // namespace

// one thread per row, any geometry
extern "C" __global__ void B200_BOUNDS cudaSpMVRowsCSR(spmat* m, double* v, CONFIG cfg, double* outV) {
    const B200ViewCSR* vw = csr_view(m);
    if (vw && whole_warps()) {
        sell_rows(vw, v, outV);
        if (vw->nlong) long_rows(vw, m->AS, v, outV);
        return;
    }
    const ulong M = m->M;
    const ulong* __restrict__ irp = m->IRP;
    const ulong* __restrict__ ja = m->JA;
    const double* __restrict__ as = m->AS;
    const double* __restrict__ x = v;
    for (ulong row = lin_tid(); row < M; row += lin_threads()) {
        const ulong s = __ldg(irp + row), e = __ldg(irp + row + 1);
        double acc = 0;
        for (ulong j = s; j < e; ++j) acc = __dadd_rn(acc, __dmul_rn(__ldg(as + j), __ldg(x + __ldg(ja + j))));
        outV[row] = acc;
    }
}

// one (sub-)warp per row
extern "C" __global__ void B200_BOUNDS cudaSpMVWarpPerRowCSR(spmat* m, double* v, CONFIG cfg, double* outV) {
    const B200ViewCSR* vw = csr_view(m);
    if (vw && whole_warps()) {
        const unsigned M = (unsigned) m->M;
        switch (vw->lanes) {
            case 2: vector_rows<2>(vw, M, m->AS, v, outV); break;
            case 4: vector_rows<4>(vw, M, m->AS, v, outV); break;
            case 8: vector_rows<8>(vw, M, m->AS, v, outV); break;
            case 16: vector_rows<16>(vw, M, m->AS, v, outV); break;
            default: vector_rows<32>(vw, M, m->AS, v, outV); break;
        }
        if (vw->nlong) long_rows(vw, m->AS, v, outV);
        return;
    }
    const ulong M = m->M;
    const ulong* __restrict__ irp = m->IRP;
    const ulong* __restrict__ ja = m->JA;
    const double* __restrict__ as = m->AS;
    const double* __restrict__ x = v;
    const unsigned lane = (unsigned) (lin_tid() & 31);
    for (ulong row = lin_tid() >> 5; row < M; row += lin_threads() >> 5) {
        const ulong s = __ldg(irp + row), e = __ldg(irp + row + 1);
        double acc = 0;
        for (ulong j = s + lane; j < e; j += 32) acc = fma(__ldg(as + j), __ldg(x + __ldg(ja + j)), acc);
        acc = warp_tree(acc);
        if (lane == 0) outV[row] = acc;
    }
}

// column-major ("transposed") ELL as built by ellTranspose (src/commons/sparseUtils.c:145-185): the struct holds
// M = slots per row (K), MAX_ROW_NZ = number of matrix rows, pitch in elements.  The plain tier visits all K slots like the
// reference kernel does (the transposed struct's RL is only partially uploaded by the reference); the fast tier
// reads 32-bit ids and stops at the row length.
extern "C" __global__ void B200_BOUNDS cudaSpMVRowsELL(spmat* m, double* v, CONFIG cfg, double* outV) {
    const B200ViewELL* vw = ell_view(m);
    if (vw && whole_warps()) {
        ell_cm_rows(vw, m->AS, m->pitchAS, v, outV);
        return;
    }
    const ulong rows = m->MAX_ROW_NZ, K = m->M, pA = m->pitchAS, pJ = m->pitchJA;
    const ulong* __restrict__ ja = m->JA;
    const double* __restrict__ as = m->AS;
    const double* __restrict__ x = v;
    for (ulong row = lin_tid(); row < rows; row += lin_threads()) {
        double acc = 0;
        for (ulong k = 0; k < K; ++k) acc = __dadd_rn(acc, __dmul_rn(__ldg(as + k * pA + row), __ldg(x + __ldg(ja + k * pJ + row))));
        outV[row] = acc;
    }
}

// row-major pitched ELL, one thread per row.  Fast tier: the view holds a column-major copy (ids and values), so the walk is
// coalesced instead of one 8-byte element per 32-byte sector.
extern "C" __global__ void B200_BOUNDS cudaSpMVRowsELLNNTransposed(spmat* m, double* v, CONFIG cfg, double* outV) {
    const B200ViewELL* vw = ell_view(m);
    if (vw && vw->as_cm && whole_warps()) {
        ell_cm_rows(vw, vw->as_cm, vw->pitch, v, outV);
        return;
    }
    const ulong M = m->M, K = m->MAX_ROW_NZ, pA = m->pitchAS, pJ = m->pitchJA;
    const ulong* __restrict__ ja = m->JA;
    const double* __restrict__ as = m->AS;
    const double* __restrict__ x = v;
#ifdef ROWLENS
    const ulong* __restrict__ rl = m->RL;
#endif
    for (ulong row = lin_tid(); row < M; row += lin_threads()) {
#ifdef ROWLENS
        const ulong len = __ldg(rl + row);
#else
        const ulong len = K;
#endif
        double acc = 0;
        for (ulong k = 0; k < len; ++k) acc = __dadd_rn(acc, __dmul_rn(__ldg(as + row * pA + k), __ldg(x + __ldg(ja + row * pJ + k))));
        outV[row] = acc;
    }
}

// row-major pitched ELL, one warp per row in the reference.  Fast tier: the same column-major walk as above -- for the short rows ELL
// is used for (K of a few tens) a warp per row leaves most lanes idle, a thread per row on coalesced storage does not.
extern "C" __global__ void B200_BOUNDS cudaSpMVWarpsPerRowELLNTrasposed(spmat* m, double* v, CONFIG cfg, double* outV) {
    const B200ViewELL* vw = ell_view(m);
    if (vw && vw->as_cm && whole_warps()) {
        ell_cm_rows(vw, vw->as_cm, vw->pitch, v, outV);
        return;
    }
    const ulong M = m->M, K = m->MAX_ROW_NZ, pA = m->pitchAS, pJ = m->pitchJA;
    const ulong* __restrict__ ja = m->JA;
    const double* __restrict__ as = m->AS;
    const double* __restrict__ x = v;
#ifdef ROWLENS
    const ulong* __restrict__ rl = m->RL;
#endif
    const unsigned lane = (unsigned) (lin_tid() & 31);
    for (ulong row = lin_tid() >> 5; row < M; row += lin_threads() >> 5) {
#ifdef ROWLENS
        const ulong len = __ldg(rl + row);
#else
        const ulong len = K;
#endif
        double acc = 0;
        for (ulong k = lane; k < len; k += 32) acc = fma(__ldg(as + row * pA + k), __ldg(x + __ldg(ja + row * pJ + k)), acc);
        acc = warp_tree(acc);
        if (lane == 0) outV[row] = acc;
    }
}
