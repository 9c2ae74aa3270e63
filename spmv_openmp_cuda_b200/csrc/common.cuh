// common.cuh -- shared helpers for the B200 SpMV engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

namespace spmvb200 {

// ---------------------------------------------------------------------------------------------
// error plumbing: 0 = OK, non-zero = failure, message on stderr (reference convention:
// ERRPRINT, src/include/macros.h:57-58) and kept for spmvb200_last_error().
// ---------------------------------------------------------------------------------------------
extern thread_local char g_err[512];
extern thread_local int g_quiet;  // set while probing an optional format: the message is kept but not printed

inline int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (!g_quiet) fprintf(stderr, "\33[31m\33[1m\33[44mspmv_b200: %s\33[0m\n", g_err);
    return 1;
}

#define CU_TRY(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return ::spmvb200::fail("%s -> %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// counter-based hashing (splitmix64 finaliser): every synthetic value is a pure function of
// (seed, indices) so host and device generators agree bit for bit.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash2(uint64_t seed, uint64_t a) { return mix64(mix64(seed) ^ a); }
__host__ __device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
    return mix64(hash2(seed, a) + 0xD1B54A32D192ED03ull * (b + 1));
}
// uniform in [0,1) with 53 bits
__host__ __device__ __forceinline__ double u01(uint64_t h) { return (double) (h >> 11) * (1.0 / 9007199254740992.0); }
// uniform in (-1,1)
__host__ __device__ __forceinline__ double usym(uint64_t h) { return 2.0 * u01(h) - 1.0; }
// uniform integer in [0,n) (n < 2^32) by multiply-shift
__host__ __device__ __forceinline__ uint64_t urange(uint64_t h, uint64_t n) {
#ifdef __CUDA_ARCH__
    return __umul64hi(h, n);
#else
    return (uint64_t) (((unsigned __int128) h * n) >> 64);
#endif
}

// ---------------------------------------------------------------------------------------------
// device-side PTX helpers (sm_100a)
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// L2 eviction policies for cp.async.bulk / ld cache hints
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// streaming (read-once) loads: bypass L1 allocation, evict-first in L2
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_stream(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ uint2 ld_stream(const uint2* p) { return __ldcs(p); }
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldcs(p); }
// x gather: read-only path, normal L1/L2 allocation (we WANT x to stay cached)
__device__ __forceinline__ double ld_x(const double* x, uint32_t c) { return __ldg(x + c); }

// L1 policy pair for gather-bound kernels on matrices with HOT columns (power-law graphs: the 7 296 hottest 32-byte sectors of x --
// what a 228 KB L1 holds -- serve 46 % of R-MAT's gathers): the matrix stream must not allocate in L1 at all, x lines are kept.
// POL = 0: the default pair above (ld.global.cs / ld.global.nc).  POL = 1: L1::no_allocate for the stream, L1::evict_last for x.
template <int POL>
__device__ __forceinline__ double ld_mat(const double* p) {
    if (POL == 0) return __ldcs(p);
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <int POL>
__device__ __forceinline__ uint32_t ld_mat(const uint32_t* p) {
    if (POL == 0) return __ldcs(p);
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <int POL>
__device__ __forceinline__ double2 ld_mat(const double2* p) {
    if (POL == 0) return __ldcs(p);
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
template <int POL>
__device__ __forceinline__ uint2 ld_mat(const uint2* p) {
    if (POL == 0) return __ldcs(p);
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
template <int POL>
__device__ __forceinline__ double ld_xp(const double* x, uint32_t c) {
    if (POL == 0) return __ldg(x + c);
    double v;
    asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(x + c));
    return v;
}

// Fused output delivery (multi-GPU x <- y iterations): besides y[row], the kernel that computes a row stores it to every
// destination that wants global row (row + row_offset) -- peer-mapped vectors of the other GPUs of the box, written over
// NVLink as posted stores while the SpMV is still running (the all-gather is the kernel's epilogue, not a collective after it).
// Fused NEIGHBOUR SYNCHRONISATION of the iterated step (spmvb200_shard_step): instead of a barrier kernel across all GPUs after every
// SpMV, the SpMV kernel itself synchronises -- and only its BOUNDARY CTAs take part: the CTAs whose rows are delivered to a peer or
// whose x windows reach outside the columns this rank owns (for a banded matrix: the row blocks within the band width of the slab's
// two edges; every other CTA touches nothing a peer writes or reads and starts / ends as usual).  A boundary CTA
//   (1) waits, before its first access, until each of the nsync neighbours has finished the PREVIOUS step (my_flags[peer] >=
//       wait_epoch: their deliveries into my x have landed, and they no longer read the buffer this step's deliveries overwrite);
//   (2) counts itself out after its last store (system-scope fence, then a ticket); the last of the nboundary boundary CTAs publishes
//       sig_epoch into every neighbour's flag array.
// One launch per step, no spinning or fencing in the other CTAs, and a rank only ever waits for the ranks it exchanges rows with.
// Measured at 2 GPUs on cfg4 (1.21 ms kernel): EVERY CTA polling and fencing +20 % (8192 one-per-SM CTAs x ~3 us of serial latency);
// boundary CTAs only, but each CTA working out whether it is one from the tile lists, +3.8 % (two more dependent loads at every CTA
// start); a precomputed per-CTA flag: see DESIGN.md; the separate barrier kernel: +2.3 %.
struct PushArgs {
    int n;
    double* dst[8];
    uint32_t lo[8], hi[8];
    uint32_t row_offset;
    int nsync;               // 0: no fused synchronisation
    uint32_t* my_flags;      // my flag array: cell p is written by rank p
    uint32_t* peer_cell[8];  // cell [my rank] of neighbour i's flag array (peer-mapped)
    uint8_t peer_rank[8];
    uint32_t wait_epoch, sig_epoch;
    uint32_t* ticket;        // device counter: boundary CTAs of this launch that have finished (reset by the last one)
    uint32_t nboundary;      // boundary CTAs of this launch (counted once per handle and launch shape)
    const uint8_t* cta_boundary;  // [grid] 1 = boundary CTA (precomputed: the kernel must not pay dependent loads to find out)
    uint32_t rb_rot;         // one-CTA-per-row-block launches: CTA b takes row block (b + rb_rot) mod grid, so that the boundary row
                             // blocks at the LOW end of the slab are scheduled last, like those at the high end -- every signal is
                             // then sent at the end of a kernel and first needed at the end of the next one: a whole kernel of slack
                             // between neighbours instead of none (the low-end CTAs would otherwise start first and stall on a
                             // neighbour that is a few microseconds behind)
    uint32_t own_lo, own_hi; // columns of x this rank owns: [own_lo, own_hi)
};
__device__ __forceinline__ void push_out(const PushArgs& p, uint32_t row, double v) {
    const uint32_t g = row + p.row_offset;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < p.n && g >= p.lo[i] && g < p.hi[i]) p.dst[i][g] = v;
}
__device__ __forceinline__ bool push_rows_wanted(const PushArgs& p, uint32_t row_lo, uint32_t row_hi) {  // local rows [row_lo, row_hi)
    const uint32_t a = row_lo + p.row_offset, b = row_hi + p.row_offset;
    bool hit = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) hit |= (i < p.n && a < p.hi[i] && b > p.lo[i]);
    return hit;
}
// one thread of a boundary CTA, before the CTA's first access to x or to a peer
__device__ __forceinline__ void push_sync_wait(const PushArgs& p) {
    uint64_t t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < p.nsync; ++i) {
        uint32_t seen, spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.my_flags + p.peer_rank[i]) : "memory");
            if ((++spins & 0xfffu) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) __trap();  // a neighbour that never arrives must not hang the GPU
            }
        } while ((int32_t) (seen - p.wait_epoch) < 0);
    }
}
// one thread of a boundary CTA, after every thread of the CTA has made its last store (CTA-wide barrier before the call)
__device__ __forceinline__ void push_sync_signal(const PushArgs& p) {
    __threadfence_system();
    const uint32_t done = atomicAdd(p.ticket, 1u);
    if (done == p.nboundary - 1u) {
        *p.ticket = 0u;
        __threadfence_system();
        for (int i = 0; i < p.nsync; ++i)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.peer_cell[i]), "r"(p.sig_epoch) : "memory");
    }
}

template <int LANES>
__device__ __forceinline__ double subwarp_sum(double v) {
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off, 32);
    return v;
}
#endif  // __CUDACC__

}  // namespace spmvb200
