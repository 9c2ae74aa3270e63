// kernels.cuh -- hand-written sm_100a SpMV kernels (fp64 values, 32-bit indices).
//
// Replaces the reference's five __global__ kernels (src/SpMV_CUDA.cu:33-135):
//   csr_stream_kernel   <- cudaSpMVRowsCSR            (thread-per-row semantics, bit-exact with sgemvSerial)
//                          and the new row-length-adaptive mode (ADAPTIVE=true)
//   csr_vector_kernel   <- cudaSpMVWarpPerRowCSR      (sub-warp per row, 128-bit loads, shuffle reduction)
//   ell_colmajor_kernel <- cudaSpMVRowsELL            (pitched column-major ELL, row-length early exit)
//   ell_rowmajor_kernel <- cudaSpMVRowsELLNNTransposed / cudaSpMVWarpsPerRowELLNTrasposed
//
// All are HBM-bound (2 flop per >= 12 bytes): the design goal is bytes in flight and no wasted
// sectors, not math throughput.  See DESIGN.md for the roofline of each.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace spmvb200 {

// ---------------------------------------------------------------------------------------------
// Row-block plan for the CSR stream kernel.
//   tile b covers rows [row0[b], row0[b+1]) and non-zeros [nnz0[b], nnz0[b+1]) -- always whole rows,
//   at most TILE non-zeros and TILE_ROWS rows -- unless bit 31 of row0 is set: then the tile is one
//   SEGMENT (<= TILE non-zeros) of the long row (row0 & 0x7fffffff); `aux` indexes its LongRec.
// The array has ntiles+1 entries (sentinel {M, NZ}).
// ---------------------------------------------------------------------------------------------
struct __align__(16) TileDesc {
    uint32_t row0, nnz0, aux, pad;
};
struct __align__(16) LongRec {
    uint32_t row, first_tile, ntiles, pad;
};
constexpr uint32_t SEG_FLAG = 0x80000000u;

constexpr int STREAM_TILE = 2048;        // non-zeros per tile (24 KB of values + column ids)
constexpr int STREAM_BLOCK = 128;        // threads per CTA
constexpr int STREAM_TILE_ROWS = 512;    // rows per tile (row-pointer slice in shared memory); 8 CTAs/SM fit
constexpr int STREAM_LONG_T = 64;         // ADAPTIVE: rows longer than this are reduced by a whole warp
constexpr int STREAM_PRE_T = 48;         // exact kind: a tile whose longest row exceeds this multiplies element-parallel first
constexpr int VEC_MID = 256;             // vector kernels / SELL copies: longer rows go to the per-row kernels (csr_midrow_*)

// ---------------------------------------------------------------------------------------------
// CSR "stream" kernel.  One CTA per tile:
//   1. thread 0 issues two TMA bulk copies (values, column ids) of the tile's 16-byte aligned
//      non-zero range into shared memory, completion on an mbarrier; L2 evict-first (read once);
//      the other threads meanwhile fetch the tile's row-pointer slice.
//   2. one thread per row walks its row in shared memory: value and column id come from the staged
//      tile, x from global memory (read-only path, L1/L2 resident), products are added left to
//      right with separate mul/add roundings -- the summation order of sgemvSerial
//      (src/SpMV_CSR_OMP.c:229-250), hence bit-identical results.  Adjacent lanes own adjacent rows,
//      so for stencil-like matrices a warp's x gather touches one or two 128-byte lines.
//      ADAPTIVE: rows longer than STREAM_LONG_T are instead handled by a whole warp (lane-strided
//      walk, shuffle tree), and even-length rows are walked from a per-lane rotated start so that
//      equal-length rows do not collide on shared-memory banks (order differs => tolerance, not
//      bit-exact).
//   Segment tiles (pieces of a row longer than TILE) block-reduce to one partial; the last
//   segment to finish (ticket counter) adds the partials in segment order => deterministic.
// The matrix stream never touches L1 or registers before it is consumed: bytes in flight are
// 8 resident CTAs x 24 KB per SM, independent of occupancy and register count.
// ---------------------------------------------------------------------------------------------
template <int TILE, int BLOCK, int TILE_ROWS, bool ADAPTIVE, int VARIANT>
__global__ void __launch_bounds__(BLOCK, 8)
csr_stream_kernel(const TileDesc* __restrict__ desc, const LongRec* __restrict__ longrec,
                  const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, const double* __restrict__ as,
                  const double* __restrict__ x, double* __restrict__ y, double* __restrict__ partial,
                  uint32_t* __restrict__ ticket, uint32_t tile_base, uint32_t pre_t) {
    constexpr int CAP = TILE + 8;
    constexpr int NWARPS = BLOCK / 32;
    constexpr int MAXLONG = TILE / STREAM_LONG_T;

    __shared__ __align__(128) double s_val[CAP];
    __shared__ __align__(16) uint32_t s_col[CAP];
    __shared__ uint32_t s_rp[TILE_ROWS + 1];
    __shared__ uint32_t s_long[ADAPTIVE ? MAXLONG : 1];
    __shared__ double s_red[NWARPS];
    __shared__ uint32_t s_nlong;
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t tid = threadIdx.x;
    const uint32_t b = blockIdx.x + tile_base;  // tile_base > 0: a row chunk of the pipelined host path
    const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(desc + b));
    const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(desc + b + 1));
    const bool seg = (d0.x & SEG_FLAG) != 0;
    const uint32_t r0 = d0.x & ~SEG_FLAG;
    const uint32_t n0 = d0.y, n1 = d1.y;
    const uint32_t nnz = n1 - n0;
    const uint32_t a0 = n0 & ~3u;               // 16-byte aligned start (4 x u32, 4 x f64 = 32 B)
    const uint32_t cnt = ((n1 + 3u) & ~3u) - a0;  // <= TILE + 6
    const uint32_t nrows = seg ? 0u : (d1.x & ~SEG_FLAG) - r0;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        if (nnz) {
            const uint64_t pol = policy_evict_first();
            mbar_arrive_expect_tx(&s_bar, cnt * 12u);
            bulk_g2s(s_val, as + a0, cnt * 8u, &s_bar, pol);
            bulk_g2s(s_col, ja + a0, cnt * 4u, &s_bar, pol);
        }
        s_nlong = 0;
    }
    // row pointers as positions inside the staged tile (which starts at non-zero a0)
    if (!seg)
        for (uint32_t i = tid; i <= nrows; i += BLOCK) s_rp[i] = __ldg(irp + r0 + i) - a0;
    __syncthreads();
    if (nnz) mbar_wait(&s_bar, 0);

    if (seg) {  // ---- one piece of a long row: block sum -> partial -> ordered combine by the last arriver
        double t = 0;
        const uint32_t lo = n0 - a0, hi = n1 - a0;
        for (uint32_t j = lo + tid; j < hi; j += BLOCK) t = fma(s_val[j], ld_x(x, s_col[j]), t);
        t = subwarp_sum<32>(t);
        if ((tid & 31) == 0) s_red[tid >> 5] = t;
        __syncthreads();
        if (tid == 0) {
            double tot = 0;
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) tot += s_red[w];
            const LongRec rec = longrec[d0.z];
            partial[b] = tot;
            __threadfence();
            const uint32_t done = atomicAdd(ticket + d0.z, 1u);
            if (done == rec.ntiles - 1) {
                __threadfence();
                double acc = 0;
                for (uint32_t k = 0; k < rec.ntiles; ++k) acc += __ldcg(partial + rec.first_tile + k);
                y[rec.row] = acc;
                ticket[d0.z] = 0;  // ready for the next launch
            }
        }
        return;
    }

    constexpr int RI = TILE_ROWS / BLOCK;
    static_assert(TILE_ROWS % BLOCK == 0, "tile rows must be a multiple of the block");
    if (!ADAPTIVE) {
        // ---- exact kind, tile with a long row: a thread walking e.g. 1500 non-zeros alone exposes one L2 gather after the other
        // while its 127 neighbours idle.  All threads first turn the staged values into PRODUCTS (element-parallel: 16 independent
        // gathers per thread), then the row's owner adds them left to right from shared memory: mul and add keep their separate
        // roundings and their order, so the result is still bit-identical to sgemvSerial.
        uint32_t mx = 0;
        for (uint32_t r = tid; r < nrows; r += BLOCK) mx = max(mx, s_rp[r + 1] - s_rp[r]);
        if (__syncthreads_or(mx > pre_t)) {
            const uint32_t lo = n0 - a0, hi = n1 - a0;
#pragma unroll 4
            for (uint32_t j = lo + tid; j < hi; j += BLOCK) s_val[j] = __dmul_rn(s_val[j], ld_x(x, s_col[j]));
            __syncthreads();
            for (uint32_t r = tid; r < nrows; r += BLOCK) {
                const uint32_t e = s_rp[r + 1];
                double acc = 0;
#pragma unroll 8
                for (uint32_t j = s_rp[r]; j < e; ++j) acc = __dadd_rn(acc, s_val[j]);
                y[r0 + r] = acc;
            }
            return;
        }
    }
    if (nrows <= (uint32_t) BLOCK) {
        // ---- at most one row per thread: walk it with 4 independent gathers in flight
        if (tid < nrows) {
            const uint32_t s = s_rp[tid], e = s_rp[tid + 1], len = e - s;
            double acc = 0;
            bool mine = true;
            if (ADAPTIVE) {
                if (len > (uint32_t) STREAM_LONG_T) {  // a whole warp takes this row below
                    s_long[atomicAdd(&s_nlong, 1u)] = tid;
                    mine = false;
                } else {
                    uint32_t j = (len & 1u) ? 0u : (tid & 31u) % (len ? len : 1u);  // rotated start for even lengths (bank conflicts)
#pragma unroll 4
                    for (uint32_t k = 0; k < len; ++k) {
                        acc = fma(s_val[s + j], ld_x(x, s_col[s + j]), acc);
                        if (++j == len) j = 0;
                    }
                }
            } else {
#pragma unroll 4
                for (uint32_t j = s; j < e; ++j) acc = __dadd_rn(acc, __dmul_rn(s_val[j], ld_x(x, s_col[j])));
            }
            if (mine) y[r0 + tid] = acc;
        }
    } else {
        // ---- short-row tile (hundreds of rows): RI rows per thread INTERLEAVED (rows tid, tid+BLOCK, ...).  Their
        // shared-memory reads and x gathers are independent, so RI gathers are in flight per thread instead of one
        // exposed gather latency per row.
        uint32_t rs[RI], rlen[RI], rot[RI];
        double acc[RI];
        uint32_t maxlen = 0;
#pragma unroll
        for (int q = 0; q < RI; ++q) {
            const uint32_t r = tid + q * BLOCK;
            acc[q] = 0;
            rs[q] = 0;
            rlen[q] = 0;
            rot[q] = 0;
            if (r < nrows) {
                rs[q] = s_rp[r];
                rlen[q] = s_rp[r + 1] - rs[q];
                if (ADAPTIVE) {
                    if (rlen[q] > (uint32_t) STREAM_LONG_T) {
                        s_long[atomicAdd(&s_nlong, 1u)] = r;
                        rlen[q] = 0;
                    } else if (!(rlen[q] & 1u) && rlen[q]) {
                        rot[q] = (tid & 31u) % rlen[q];
                    }
                }
            }
            maxlen = max(maxlen, rlen[q]);
        }
#pragma unroll 2
        for (uint32_t k = 0; k < maxlen; ++k) {
#pragma unroll
            for (int q = 0; q < RI; ++q) {
                if (k < rlen[q]) {
                    uint32_t j = k;
                    if (ADAPTIVE) {
                        j = k + rot[q];
                        if (j >= rlen[q]) j -= rlen[q];
                    }
                    const double v = s_val[rs[q] + j];
                    const double xv = ld_x(x, s_col[rs[q] + j]);
                    acc[q] = ADAPTIVE ? fma(v, xv, acc[q]) : __dadd_rn(acc[q], __dmul_rn(v, xv));
                }
            }
        }
#pragma unroll
        for (int q = 0; q < RI; ++q) {
            const uint32_t r = tid + q * BLOCK;
            if (r < nrows && !(ADAPTIVE && s_rp[r + 1] - s_rp[r] > (uint32_t) STREAM_LONG_T)) y[r0 + r] = acc[q];
        }
    }
    if (ADAPTIVE) {
        __syncthreads();
        const uint32_t nl = s_nlong, lane = tid & 31;
        for (uint32_t k = tid >> 5; k < nl; k += NWARPS) {
            const uint32_t r = s_long[k];
            const uint32_t s = s_rp[r], e = s_rp[r + 1];
            double acc = 0;
            for (uint32_t j = s + lane; j < e; j += 32) acc = fma(s_val[j], ld_x(x, s_col[j]), acc);
            acc = subwarp_sum<32>(acc);
            if (lane == 0) y[r0 + r] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CSR "vector" kernel: LANES (2..32, chosen from the mean row length) lanes per row, each lane
// loads two consecutive non-zeros per step with one 128-bit value load and one 64-bit index load
// (row starts are peeled to an even index so the loads stay aligned), shuffle-tree reduction.
// Replaces cudaSpMVWarpPerRowCSR (src/SpMV_CUDA.cu:52-73) and fixes its blockIdx.y defect
// (SURVEY.md §2.3-1): the row comes from a linear thread id.
// ---------------------------------------------------------------------------------------------
template <int LANES, int BLOCK, int POL = 0>
__global__ void __launch_bounds__(BLOCK)
csr_vector_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, const double* __restrict__ as,
                  const double* __restrict__ x, double* __restrict__ y, uint32_t row_begin, uint32_t M, uint32_t max_len) {
    const uint32_t gt = blockIdx.x * BLOCK + threadIdx.x;
    const uint32_t row = row_begin + gt / LANES, lane = gt % LANES;
    double acc = 0;
    bool mine = row < M;
    if (mine) {
        const uint32_t s = __ldg(irp + row), e = __ldg(irp + row + 1);
        mine = e - s <= max_len;  // longer rows belong to csr_longrow_kernel (same launch sequence)
        if (mine)
            for (uint32_t i = (s & ~1u) + 2 * lane; i < e; i += 2 * LANES) {
                const double2 v = ld_mat<POL>(reinterpret_cast<const double2*>(as + i));
                const uint2 c = ld_mat<POL>(reinterpret_cast<const uint2*>(ja + i));
                if (i >= s) acc = fma(v.x, ld_xp<POL>(x, c.x), acc);
                if (i + 1 < e) acc = fma(v.y, ld_xp<POL>(x, c.y), acc);
            }
    }
    acc = subwarp_sum<LANES>(acc);
    if (lane == 0 && mine) y[row] = acc;
}

// Same row algorithm, but each CTA owns a CONTIGUOUS, nnz-balanced span of rows (span_b) and walks it front to
// back with all its threads: with one or two big CTAs per SM the x window an SM touches slides slowly, so the
// gathers of banded / locally-coupled matrices hit the 228 KB L1 instead of going to L2 (whose sector rate is what
// bounds random gathers on B200).
template <int LANES, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
csr_vector_span_kernel(const uint32_t* __restrict__ span_b, const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja,
                       const double* __restrict__ as, const double* __restrict__ x, double* __restrict__ y, uint32_t max_len) {
    constexpr uint32_t ROWS = BLOCK / LANES;
    const uint32_t r_end = span_b[blockIdx.x + 1];
    const uint32_t sub = threadIdx.x / LANES, lane = threadIdx.x % LANES;
    for (uint32_t r0 = span_b[blockIdx.x]; r0 < r_end; r0 += ROWS) {
        const uint32_t row = r0 + sub;
        double acc = 0;
        bool mine = row < r_end;
        if (mine) {
            const uint32_t s = __ldg(irp + row), e = __ldg(irp + row + 1);
            mine = e - s <= max_len;
            if (mine)
                for (uint32_t i = (s & ~1u) + 2 * lane; i < e; i += 2 * LANES) {
                    const double2 v = ld_stream(reinterpret_cast<const double2*>(as + i));
                    const uint2 c = ld_stream(reinterpret_cast<const uint2*>(ja + i));
                    if (i >= s) acc = fma(v.x, ld_x(x, c.x), acc);
                    if (i + 1 < e) acc = fma(v.y, ld_x(x, c.y), acc);
                }
        }
        acc = subwarp_sum<LANES>(acc);
        if (lane == 0 && mine) y[row] = acc;
    }
}

// Rows longer than one tile, for the vector kernel: one CTA per <= TILE-non-zero segment (the plan's
// segment tiles), direct coalesced loads, block sum, ordered combine by the last segment to finish.
template <int BLOCK, int POL = 0>
__global__ void __launch_bounds__(BLOCK)
csr_longrow_kernel(const uint32_t* __restrict__ seg_tiles, const TileDesc* __restrict__ desc, const LongRec* __restrict__ longrec,
                   const uint32_t* __restrict__ ja, const double* __restrict__ as, const double* __restrict__ x,
                   double* __restrict__ y, double* __restrict__ partial, uint32_t* __restrict__ ticket) {
    __shared__ double s_red[BLOCK / 32];
    const uint32_t tid = threadIdx.x, b = seg_tiles[blockIdx.x];
    const uint32_t n0 = desc[b].nnz0, n1 = desc[b + 1].nnz0, rec_id = desc[b].aux;
    double t = 0;
    for (uint32_t j = n0 + tid; j < n1; j += BLOCK) t = fma(ld_mat<POL>(as + j), ld_xp<POL>(x, ld_mat<POL>(ja + j)), t);
    t = subwarp_sum<32>(t);
    if ((tid & 31) == 0) s_red[tid >> 5] = t;
    __syncthreads();
    if (tid == 0) {
        double tot = 0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; ++w) tot += s_red[w];
        const LongRec rec = longrec[rec_id];
        partial[b] = tot;
        __threadfence();
        const uint32_t done = atomicAdd(ticket + rec_id, 1u);
        if (done == rec.ntiles - 1) {
            __threadfence();
            double acc = 0;
            for (uint32_t k = 0; k < rec.ntiles; ++k) acc += __ldcg(partial + rec.first_tile + k);
            y[rec.row] = acc;
            ticket[rec_id] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Column-major pitched ELL, one thread per row: slot k of row r at k*pitch + r, so a warp reads 32
// consecutive values (256 B) and 32 consecutive column ids (128 B) per slot -- fully coalesced.
// Row-length early exit: the loop bound is the warp's longest row, shorter rows are predicated
// off, padding is never fetched by a lane (the reference kernel walks all K slots and gathers
// x[0] for padding, src/SpMV_CUDA.cu:83-93).  UNROLL independent slots are in flight per thread.
// Per-row sum is left to right with separate mul/add => bit-identical to sgemvSerial.
// ---------------------------------------------------------------------------------------------
// IDX16: the column ids are stored as 16-bit offsets from (row + base) -- possible whenever max(col - row) - min(col - row)
// < 2^16 over the whole handle (every stencil / banded matrix of BASELINE.json): 10 instead of 12 bytes per non-zero.
__device__ __forceinline__ uint32_t ld_stream(const uint16_t* p) { return (uint32_t) __ldcs(reinterpret_cast<const unsigned short*>(p)); }
template <int UNROLL, int BLOCK, bool IDX16>
__global__ void __launch_bounds__(BLOCK)
ell_colmajor_kernel(const double* __restrict__ as, const void* __restrict__ ja_any, const uint32_t* __restrict__ rl,
                    uint64_t pitch, uint32_t row_begin, uint32_t M, uint32_t K, int32_t base, const double* __restrict__ x, double* __restrict__ y,
                    const PushArgs push) {
    using idx_t = typename std::conditional<IDX16, uint16_t, uint32_t>::type;
    const uint32_t row = row_begin + blockIdx.x * BLOCK + threadIdx.x;  // rows [row_begin, M)
    const bool live = row < M;
    const uint32_t len = live ? (rl ? __ldg(rl + row) : K) : 0u;
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, len);
    const double* a = as + row;
    const idx_t* j = reinterpret_cast<const idx_t*>(ja_any) + row;
    const uint32_t cbase = IDX16 ? (uint32_t) ((int32_t) row + base) : 0u;
    double acc = 0;
    uint32_t k = 0;
    for (; k + UNROLL <= wmax; k += UNROLL) {
        double v[UNROLL];
        uint32_t c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const bool ok = k + u < len;
            v[u] = ok ? ld_stream(a + (uint64_t) (k + u) * pitch) : 0.0;
            c[u] = ok ? ld_stream(j + (uint64_t) (k + u) * pitch) + cbase : 0u;
        }
        double xv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) xv[u] = (k + u < len) ? ld_x(x, c[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (k + u < len) acc = __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
    }
    for (; k < wmax; ++k) {
        if (k < len) {
            const double v = ld_stream(a + (uint64_t) k * pitch);
            const uint32_t c = ld_stream(j + (uint64_t) k * pitch) + cbase;
            acc = __dadd_rn(acc, __dmul_rn(v, ld_x(x, c)));
        }
    }
    if (live) {
        y[row] = acc;
        if (push.n) push_out(push, row, acc);
    }
}
// Two adjacent rows per thread (IDX16 only): slot k of rows 2t, 2t+1 is one 16-byte value load and one 4-byte id load, so a thread
// has twice the bytes in flight per load instruction -- the 16-bit kernel is bound by bytes in flight (94 % occupancy, 82 KB per SM
// outstanding at most), not by anything else.  Same left-to-right sums, same early exit (per warp: the longest of its 64 rows).
// SPEC: the first UNROLL slots are fetched BEFORE the row lengths have arrived (every row of the rectangle has them in memory:
// padding is AS = 0, id = 0), so the length vector, the first values and the first ids travel together -- one dependent DRAM round
// trip less per thread, which is what an isolated launch of a small matrix is made of (cfg1: 84 MB, ~13 us of streaming under
// ~8 us of launch + latency chain).  The x gathers stay predicated on the length (a padding id is not a valid column).  The host
// asks for it only when the rectangle is nearly full (K*M <= 1.15 NZ) and K >= UNROLL.
template <int UNROLL, int BLOCK, bool SPEC = false>
__global__ void __launch_bounds__(BLOCK)
ell_colmajor_pair_kernel(const double* __restrict__ as, const uint16_t* __restrict__ ja16, const uint32_t* __restrict__ rl, uint64_t pitch,
                         uint32_t row_begin, uint32_t M, int32_t base, const double* __restrict__ x, double* __restrict__ y, const PushArgs push) {
    const uint32_t row = row_begin + 2u * (blockIdx.x * BLOCK + threadIdx.x);  // row_begin even, pitch a multiple of 64
    const bool live0 = row < M, live1 = row + 1 < M;
    uint32_t len0 = 0, len1 = 0;
    if (live1) {
        const uint2 l = __ldg(reinterpret_cast<const uint2*>(rl + row));
        len0 = l.x;
        len1 = l.y;
    } else if (live0) {
        len0 = __ldg(rl + row);
    }
    const double2* a = reinterpret_cast<const double2*>(as + row);
    const uint32_t* j = reinterpret_cast<const uint32_t*>(ja16 + row);
    const uint64_t p2 = pitch >> 1;
    const uint32_t cb0 = (uint32_t) ((int32_t) row + base), cb1 = cb0 + 1u;
    double acc0 = 0, acc1 = 0;
    uint32_t k = 0;
    if (SPEC) {
        double2 v[UNROLL];
        uint32_t c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {  // issued ahead of the first use of len0 / len1: no dependence on the length loads
            v[u] = live0 ? ld_stream(a + (uint64_t) u * p2) : make_double2(0.0, 0.0);
            c[u] = live0 ? ld_stream(j + (uint64_t) u * p2) : 0u;
        }
        double x0[UNROLL], x1[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            x0[u] = ((uint32_t) u < len0) ? ld_x(x, (c[u] & 0xffffu) + cb0) : 0.0;
            x1[u] = ((uint32_t) u < len1) ? ld_x(x, (c[u] >> 16) + cb1) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if ((uint32_t) u < len0) acc0 = __dadd_rn(acc0, __dmul_rn(v[u].x, x0[u]));
            if ((uint32_t) u < len1) acc1 = __dadd_rn(acc1, __dmul_rn(v[u].y, x1[u]));
        }
        k = UNROLL;
    }
    const uint32_t lmax = max(len0, len1);
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, lmax);
    for (; k + UNROLL <= wmax; k += UNROLL) {
        double2 v[UNROLL];
        uint32_t c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const bool ok = k + u < lmax;
            v[u] = ok ? ld_stream(a + (uint64_t) (k + u) * p2) : make_double2(0.0, 0.0);
            c[u] = ok ? ld_stream(j + (uint64_t) (k + u) * p2) : 0u;
        }
        double x0[UNROLL], x1[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            x0[u] = (k + u < len0) ? ld_x(x, (c[u] & 0xffffu) + cb0) : 0.0;
            x1[u] = (k + u < len1) ? ld_x(x, (c[u] >> 16) + cb1) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (k + u < len0) acc0 = __dadd_rn(acc0, __dmul_rn(v[u].x, x0[u]));
            if (k + u < len1) acc1 = __dadd_rn(acc1, __dmul_rn(v[u].y, x1[u]));
        }
    }
    for (; k < wmax; ++k) {
        if (k < lmax) {
            const double2 v = ld_stream(a + (uint64_t) k * p2);
            const uint32_t c = ld_stream(j + (uint64_t) k * p2);
            if (k < len0) acc0 = __dadd_rn(acc0, __dmul_rn(v.x, ld_x(x, (c & 0xffffu) + cb0)));
            if (k < len1) acc1 = __dadd_rn(acc1, __dmul_rn(v.y, ld_x(x, (c >> 16) + cb1)));
        }
    }
    // (storing the pair as one 16-byte word was measured 7 % SLOWER on cfg2: 101.0 vs 94.3 us, three runs each on one box)
    if (live0) {
        y[row] = acc0;
        if (push.n) push_out(push, row, acc0);
    }
    if (live1) {
        y[row + 1] = acc1;
        if (push.n) push_out(push, row + 1, acc1);
    }
}
// kinds whose kernels have no fused delivery: copy the finished y to the destinations that want it
__global__ void push_rows_kernel(const double* __restrict__ y, uint32_t M, const PushArgs push) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < M; r += stride) push_out(push, r, y[r]);
}
// barrier across the GPUs of a box: flags[q] is GPU q's flag array (peer-mapped on the others), epoch grows by one per use.
// Thread p tells GPU p "rank has arrived at epoch", then waits for GPU p's arrival.  System-scope release / acquire.
struct BarrierArgs {
    uint32_t* flags[8];
};
// A peer that never arrives (crashed process, a rank that skipped the call) must not hang the GPU: after 20 s the kernel traps and the
// stream reports a launch failure instead.
__global__ void peer_barrier_kernel(const BarrierArgs b, int n, int rank, uint32_t epoch) {
    const int p = threadIdx.x;
    if (p >= n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(b.flags[p] + rank), "r"(epoch) : "memory");
    uint32_t seen;
    uint64_t t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t spins = 0;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(b.flags[rank] + p) : "memory");
        if ((++spins & 0xfffu) == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 20000000000ull) __trap();
        }
    } while ((int32_t) (seen - epoch) < 0);
}
// range of (col - row) over the valid slots of a column-major ELL, then the 16-bit offsets themselves
__global__ void ell_delta_range_kernel(const uint32_t* __restrict__ ja, const uint32_t* __restrict__ rl, uint64_t pitch, uint32_t M,
                                       long long* __restrict__ mn, long long* __restrict__ mx) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    long long lo = (1ll << 62), hi = -(1ll << 62);
    if (r < M) {
        const uint32_t len = rl[r];
        for (uint32_t k = 0; k < len; ++k) {
            const long long d = (long long) ja[(uint64_t) k * pitch + r] - (long long) r;
            lo = min(lo, d);
            hi = max(hi, d);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(mn, lo);
        atomicMax(mx, hi);
    }
}
__global__ void ell_make_idx16_kernel(const uint32_t* __restrict__ ja, const uint32_t* __restrict__ rl, uint64_t pitch, uint32_t M, uint32_t K,
                                      int32_t base, uint16_t* __restrict__ ja16) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const uint32_t len = rl[r];
    for (uint32_t k = 0; k < K; ++k)
        ja16[(uint64_t) k * pitch + r] = k < len ? (uint16_t) ((long long) ja[(uint64_t) k * pitch + r] - (long long) r - base) : (uint16_t) 0;
}

// ---------------------------------------------------------------------------------------------
// SELL-32-sigma: rows are sorted by decreasing length inside windows of sigma rows (perm), cut into slices of 32
// consecutive sorted rows, each slice stored column-major (slot k of lane l at slice_ptr + 32k + l) and padded only
// to its own longest row.  One thread per row, a warp per slice: fully coalesced 256 B / 128 B loads per slot, loop
// bound = the slice length (warp uniform), per-lane predicate at the row length.  Left-to-right sum with separate
// mul/add => bit-identical to sgemvSerial.  y is written through the permutation (scatter inside one window).
// ---------------------------------------------------------------------------------------------
template <int UNROLL, int BLOCK, int POL = 0>
__global__ void __launch_bounds__(BLOCK)
sell_kernel(const uint32_t* __restrict__ slice_ptr, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ rl_sorted,
            const double* __restrict__ as, const uint32_t* __restrict__ ja, uint32_t Mpad, const double* __restrict__ x,
            double* __restrict__ y) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= Mpad) return;  // Mpad is a multiple of 32: whole warps leave together
    const uint32_t len = __ldg(rl_sorted + i);
    const uint32_t sp0 = __ldg(slice_ptr + (i >> 5)), sp1 = __ldg(slice_ptr + (i >> 5) + 1);
    const uint32_t wmax = (sp1 - sp0) >> 5;
    const double* a = as + sp0 + (i & 31);
    const uint32_t* j = ja + sp0 + (i & 31);
    double acc = 0;
    uint32_t k = 0;
    for (; k + UNROLL <= wmax; k += UNROLL) {
        double v[UNROLL];
        uint32_t c[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const bool ok = k + u < len;
            v[u] = ok ? ld_mat<POL>(a + (k + u) * 32) : 0.0;
            c[u] = ok ? ld_mat<POL>(j + (k + u) * 32) : 0u;
        }
        double xv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) xv[u] = (k + u < len) ? ld_xp<POL>(x, c[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (k + u < len) acc = __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
    }
    for (; k < wmax; ++k)
        if (k < len) acc = __dadd_rn(acc, __dmul_rn(ld_mat<POL>(a + k * 32), ld_xp<POL>(x, ld_mat<POL>(j + k * 32))));
    const uint32_t row = __ldg(perm + i);
    if (row != 0xffffffffu) y[row] = acc;
}

// SELL construction (device): sort keys, slice lengths, fill
// cap: rows longer than this are left EMPTY here (their y entry is written by other kernels right after: the adaptive mode's
// SELL + per-row-CTA hybrid for skewed matrices)
// source of a SELL build: a CSR handle (pitch = 0) or a column-major ELL handle (slot k of row r at k * pitch + r, lengths in rl)
struct RowSrc {
    const uint32_t* irp;
    const uint32_t* rl;
    const uint32_t* ja;
    const double* as;
    uint64_t pitch;
    __device__ __forceinline__ uint32_t len(uint32_t r) const { return pitch ? rl[r] : irp[r + 1] - irp[r]; }
    __device__ __forceinline__ uint64_t at(uint32_t r, uint32_t k) const { return pitch ? (uint64_t) k * pitch + r : (uint64_t) irp[r] + k; }
};
__global__ void sell_keys_kernel(const RowSrc src, uint32_t M, uint32_t Mpad, uint32_t sigma, uint32_t cap,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Mpad) return;
    uint32_t len = r < M ? src.len(r) : 0u;
    const bool capped = len > cap;  // not this copy's row: no entries here and NO write of y (other kernels own it, concurrently)
    if (capped) len = 0u;
    keys[r] = ((uint64_t) (r / sigma) << 32) | (uint64_t) (0xffffffffu - len);  // ascending sort = descending length per window
    vals[r] = (r < M && !capped) ? r : 0xffffffffu;
}
__global__ void sell_slices_kernel(const uint64_t* __restrict__ keys_sorted, uint32_t Mpad, uint32_t* __restrict__ rl_sorted,
                                   uint64_t* __restrict__ slice_slots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Mpad) return;
    const uint32_t len = 0xffffffffu - (uint32_t) (keys_sorted[i] & 0xffffffffull);
    rl_sorted[i] = len;
    if ((i & 31) == 0) slice_slots[i >> 5] = (uint64_t) len * 32;  // first row of a slice is its longest
}
__global__ void sell_fill_kernel(const RowSrc src, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ slice_ptr, uint32_t Mpad,
                                 uint32_t cap, uint32_t* __restrict__ sja, double* __restrict__ sas) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Mpad) return;
    const uint32_t row = perm[i];
    const uint32_t sp0 = slice_ptr[i >> 5], wmax = (slice_ptr[(i >> 5) + 1] - sp0) >> 5;
    uint32_t len = row != 0xffffffffu ? src.len(row) : 0u;
    if (len > cap) len = 0u;
    for (uint32_t k = 0; k < wmax; ++k) {
        const uint32_t o = sp0 + k * 32 + (i & 31);
        const uint64_t e = k < len ? src.at(row, k) : 0ull;
        sas[o] = k < len ? src.as[e] : 0.0;
        sja[o] = k < len ? src.ja[e] : 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// Row-major pitched ELL (pitch a multiple of 4 slots => rows 32-byte aligned): LANES lanes per
// row, two slots per lane and step (128-bit value / 64-bit index loads), loop bounded by the row
// length, shuffle reduction.  LANES = 32 is the reference's warp-per-row mapping
// (cudaSpMVWarpsPerRowELLNTrasposed, src/SpMV_CUDA.cu:116-135); smaller LANES (picked from K) serve
// the thread-per-row entry point cudaSpMVRowsELLNNTransposed (src/SpMV_CUDA.cu:99-115) without its
// lane-stride-pitch access pattern.
// ---------------------------------------------------------------------------------------------
template <int LANES, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
ell_rowmajor_kernel(const double* __restrict__ as, const uint32_t* __restrict__ ja, const uint32_t* __restrict__ rl,
                    uint64_t pitch, uint32_t M, uint32_t K, const double* __restrict__ x, double* __restrict__ y) {
    const uint32_t gt = blockIdx.x * BLOCK + threadIdx.x;
    const uint32_t row = gt / LANES, lane = gt % LANES;
    double acc = 0;
    if (row < M) {
        const uint32_t len = rl ? __ldg(rl + row) : K;
        const uint64_t base = (uint64_t) row * pitch;
        for (uint32_t k = 2 * lane; k < len; k += 2 * LANES) {
            const double2 v = ld_stream(reinterpret_cast<const double2*>(as + base + k));
            const uint2 c = ld_stream(reinterpret_cast<const uint2*>(ja + base + k));
            acc = fma(v.x, ld_x(x, c.x), acc);
            if (k + 1 < len) acc = fma(v.y, ld_x(x, c.y), acc);
        }
    }
    acc = subwarp_sum<LANES>(acc);
    if (lane == 0 && row < M) y[row] = acc;
}

// Warp per row with ROWS rows of a warp in flight at once (cudaSpMVWarpsPerRowELLNTrasposed, src/SpMV_CUDA.cu:116-135): a whole warp on a
// 27-entry row has 14 working lanes and one dependent chain load -> gather -> 5 shuffle steps per row; with one row per warp at a time that
// chain is exposed 2 M times (0.32 of the roofline on cfg2).  Here a warp owns ROWS consecutive rows: their loads are issued together, their
// gathers together, and the ROWS shuffle trees run interleaved.  Short rows (K <= 64): one double2 / uint2 per lane covers the row.
template <int ROWS, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
ell_rowmajor_warp_kernel(const double* __restrict__ as, const uint32_t* __restrict__ ja, const uint32_t* __restrict__ rl, uint64_t pitch,
                         uint32_t M, uint32_t K, const double* __restrict__ x, double* __restrict__ y) {
    const uint32_t warp = (blockIdx.x * BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t row0 = warp * ROWS;
    if (row0 >= M) return;
    double2 v[ROWS];
    uint2 c[ROWS];
    uint32_t len[ROWS];
    double acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t row = min(row0 + r, M - 1);  // a clamped duplicate row is computed and not stored
        len[r] = rl ? __ldg(rl + row) : K;
        acc[r] = 0.0;
    }
    for (uint32_t k = 2 * lane; k < K; k += 64) {  // one pass for K <= 64
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint64_t base = (uint64_t) min(row0 + r, M - 1) * pitch;
            v[r] = make_double2(0.0, 0.0);
            c[r] = make_uint2(0u, 0u);
            if (k < len[r]) {
                v[r] = ld_stream(reinterpret_cast<const double2*>(as + base + k));
                c[r] = ld_stream(reinterpret_cast<const uint2*>(ja + base + k));
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (k < len[r]) acc[r] = fma(v[r].x, ld_x(x, c[r].x), acc[r]);
            if (k + 1 < len[r]) acc[r] = fma(v[r].y, ld_x(x, c[r].y), acc[r]);
        }
    }
    // Four sums over 32 lanes with 6 exchanged doubles instead of 4 x 5 (the plain trees made the kernel shuffle-bound: 10 SHFL per row):
    // at distance 16 the lower half-warp keeps rows 0, 1 and takes the upper half's partials of them (and vice versa for rows 2, 3); at
    // distance 8 each quarter keeps one row; distances 4, 2, 1 finish it.  Row 2 * (lane >> 4) + ((lane >> 3) & 1) ends up in every lane of
    // its octet.
    static_assert(ROWS == 4, "the exchange below is written for four rows per warp");
    const bool up16 = lane & 16, up8 = lane & 8;
    double k0 = up16 ? acc[2] : acc[0], k1 = up16 ? acc[3] : acc[1];
    k0 += __shfl_xor_sync(0xffffffffu, up16 ? acc[0] : acc[2], 16, 32);
    k1 += __shfl_xor_sync(0xffffffffu, up16 ? acc[1] : acc[3], 16, 32);
    double k = up8 ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8, 32);
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off, 32);
    const uint32_t r = 2 * (lane >> 4) + ((lane >> 3) & 1);
    if ((lane & 7) == 0 && row0 + r < M) y[row0 + r] = k;
}

// ---------------------------------------------------------------------------------------------
// upload-side helpers (run once per matrix, not on the hot path)
// ---------------------------------------------------------------------------------------------
// 64-bit -> 32-bit index narrowing with optional rebase; flags values that do not fit
__global__ void narrow_u64_kernel(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n,
                                  uint64_t sub, int* __restrict__ overflow) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t v = src[i] - sub;
        if (v > 0xffffffffull) *overflow = 1;
        dst[i] = (uint32_t) v;
    }
}

// row-major (rows x K, 64-bit ids) host layout chunk -> column-major pitched device layout.
// 32 x 32 tiles through shared memory: coalesced on both sides.
__global__ void ell_transpose_kernel(const uint64_t* __restrict__ ja_rm, const double* __restrict__ as_rm,
                                     uint32_t rows, uint32_t K, uint32_t row_off, uint64_t pitch,
                                     uint32_t* __restrict__ ja_cm, double* __restrict__ as_cm, int* __restrict__ overflow) {
    __shared__ double t_as[32][33];
    __shared__ uint32_t t_ja[32][33];
    const uint32_t r_base = blockIdx.x * 32, k_base = blockIdx.y * 32;
    for (uint32_t i = threadIdx.y; i < 32; i += blockDim.y) {
        const uint32_t r = r_base + i, k = k_base + threadIdx.x;
        double v = 0;
        uint32_t c = 0;
        if (r < rows && k < K) {
            v = as_rm[(uint64_t) r * K + k];
            const uint64_t cc = ja_rm[(uint64_t) r * K + k];
            if (cc > 0xffffffffull) *overflow = 1;
            c = (uint32_t) cc;
        }
        t_as[i][threadIdx.x] = v;
        t_ja[i][threadIdx.x] = c;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.y; i < 32; i += blockDim.y) {
        const uint32_t k = k_base + i, r = r_base + threadIdx.x;
        if (r < rows && k < K) {
            as_cm[(uint64_t) k * pitch + row_off + r] = t_as[threadIdx.x][i];
            ja_cm[(uint64_t) k * pitch + row_off + r] = t_ja[threadIdx.x][i];
        }
    }
}

// row-major chunk -> row-major pitched narrow layout
__global__ void ell_repitch_kernel(const uint64_t* __restrict__ ja_rm, const double* __restrict__ as_rm, uint32_t rows,
                                   uint32_t K, uint32_t row_off, uint64_t pitch, uint32_t* __restrict__ ja_o,
                                   double* __restrict__ as_o, int* __restrict__ overflow) {
    const uint64_t n = (uint64_t) rows * K, stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r = i / K, k = i % K;
        const uint64_t cc = ja_rm[i];
        if (cc > 0xffffffffull) *overflow = 1;
        as_o[(row_off + r) * pitch + k] = as_rm[i];
        ja_o[(row_off + r) * pitch + k] = (uint32_t) cc;
    }
}

// effective row lengths of a row-major host-layout chunk when the caller has no RL vector:
// index of the last slot with AS != 0, plus one (trailing zero slots add nothing to the sum)
__global__ void ell_derive_rl_kernel(const double* __restrict__ as_rm, uint32_t rows, uint32_t K, uint32_t row_off,
                                     uint32_t* __restrict__ rl) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    uint32_t len = 0;
    for (uint32_t k = 0; k < K; ++k)
        if (as_rm[(uint64_t) r * K + k] != 0.0) len = k + 1;
    rl[row_off + r] = len;
}

__global__ void row_len_from_irp_kernel(const uint32_t* __restrict__ irp, uint32_t M, uint32_t* __restrict__ rl,
                                        uint32_t* __restrict__ kmax) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = 0;
    if (r < M) {
        len = irp[r + 1] - irp[r];
        if (rl) rl[r] = len;
    }
    len = __reduce_max_sync(0xffffffffu, len);
    if ((threadIdx.x & 31) == 0 && len) atomicMax(kmax, len);
}

// CSR -> ELL (either layout) on the device; padding slots get AS = 0, JA = 0 like the reference
// (src/lib/parser.c:245-252)
__global__ void csr_to_ell_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja,
                                  const double* __restrict__ as, uint32_t M, uint32_t K, uint64_t pitch, int colmajor,
                                  uint32_t* __restrict__ eja, double* __restrict__ eas) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const uint32_t s = irp[r], len = irp[r + 1] - s;
    for (uint32_t k = 0; k < K; ++k) {
        const uint64_t o = colmajor ? (uint64_t) k * pitch + r : (uint64_t) r * pitch + k;
        eas[o] = k < len ? as[s + k] : 0.0;
        eja[o] = k < len ? ja[s + k] : 0u;
    }
}

// Rows of medium length (VEC_MID < len <= TILE) for the vector kernels: one CTA per listed row, so that a sub-warp
// of the vector kernel never iterates over more than VEC_MID non-zeros while its neighbours idle.
template <int BLOCK, int POL = 0>
__global__ void __launch_bounds__(BLOCK)
csr_midrow_kernel(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja,
                  const double* __restrict__ as, const double* __restrict__ x, double* __restrict__ y, uint32_t len_lo) {
    __shared__ double s_red[BLOCK / 32];
    const uint32_t tid = threadIdx.x, row = rows[blockIdx.x];
    const uint32_t s = __ldg(irp + row), e = __ldg(irp + row + 1);
    if (e - s <= len_lo) return;  // block-uniform: the warp-per-row kernel owns this row
    double t = 0;
    for (uint32_t i = (s & ~1u) + 2 * tid; i < e; i += 2 * BLOCK) {
        const double2 v = ld_mat<POL>(reinterpret_cast<const double2*>(as + i));
        const uint2 c = ld_mat<POL>(reinterpret_cast<const uint2*>(ja + i));
        if (i >= s) t = fma(v.x, ld_xp<POL>(x, c.x), t);
        if (i + 1 < e) t = fma(v.y, ld_xp<POL>(x, c.y), t);
    }
    t = subwarp_sum<32>(t);
    if ((tid & 31) == 0) s_red[tid >> 5] = t;
    __syncthreads();
    if (tid == 0) {
        double tot = 0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; ++w) tot += s_red[w];
        y[row] = tot;
    }
}

// Same rows, one WARP per row (8 rows per CTA): most medium rows of a power-law matrix sit just above VEC_MID, where a 128-thread
// CTA spends its time in the block reduction; a warp keeps 64 rows in flight per SM instead of 16.  Rows above MIDW_MAX stay with
// the CTA-per-row kernel.
constexpr int MIDW_MAX = 1024;  // measured on R-MAT: 1024 + CTA-per-row above it beats a warp for every medium row (290 vs 308 us)
template <int BLOCK, int POL = 0>
__global__ void __launch_bounds__(BLOCK)
csr_midrow_warp_kernel(const uint32_t* __restrict__ rows, uint32_t nrows, const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja,
                       const double* __restrict__ as, const double* __restrict__ x, double* __restrict__ y, uint32_t len_lo, uint32_t len_hi) {
    const uint32_t w = (blockIdx.x * BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nrows) return;
    const uint32_t row = __ldg(rows + w);
    const uint32_t s = __ldg(irp + row), e = __ldg(irp + row + 1);
    if (e - s <= len_lo || e - s > len_hi) return;  // the other medium-row kernel owns this row
    double t = 0;
    for (uint32_t i = (s & ~1u) + 2 * lane; i < e; i += 64) {
        const double2 v = ld_mat<POL>(reinterpret_cast<const double2*>(as + i));
        const uint2 c = ld_mat<POL>(reinterpret_cast<const uint2*>(ja + i));
        if (i >= s) t = fma(v.x, ld_xp<POL>(x, c.x), t);
        if (i + 1 < e) t = fma(v.y, ld_xp<POL>(x, c.y), t);
    }
    t = subwarp_sum<32>(t);
    if (lane == 0) y[row] = t;
}

// Same rows for the BIT-EXACT kind: one warp per row.  A lane owns C = 8 CONSECUTIVE non-zeros of every 256-entry chunk (two 16-byte
// column loads, four 16-byte value loads, 8 independent gathers), multiplies them, and the row's running sum then travels through the
// lanes in order: at step l every lane adds its own 8 products to the sum so far, lane l's result is the one broadcast (one 64-bit shuffle
// per 8 non-zeros).  mul and add keep their separate roundings and the serial order: bit-identical to sgemvSerial
// (src/SpMV_CSR_OMP.c:229-250), with all loads 32 lanes wide.  What remains serial is the chain of dependent DADDs, one per non-zero.
// Two earlier versions, both measured on the R-MAT medium rows (25 M non-zeros; the tolerance kernel above needs 97 us): every product
// fetched by shuffle (2 SHFL per non-zero) 250 us; products parked in shared memory and read back as broadcasts 193-204 us (ncu: l1tex
// 85 %, one shared-memory wavefront per non-zero on top of the gathers; its partial-chunk loop also had the LDS latency in the chain).
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
csr_midrow_exact_kernel(const uint32_t* __restrict__ rows, uint32_t nrows, const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja,
                        const double* __restrict__ as, const double* __restrict__ x, double* __restrict__ y) {
    constexpr int C = 8, CH = 32 * C;
    const uint32_t w = (blockIdx.x * BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nrows) return;  // warp-uniform
    const uint32_t row = __ldg(rows + w);
    const uint32_t s = __ldg(irp + row), e = __ldg(irp + row + 1);
    double acc = 0;
    for (uint32_t base = s & ~3u; base < e; base += CH) {  // 16-byte aligned chunk start; entries before s / from e on are masked
        const uint32_t i0 = base + C * lane;
        double p[C];
        uint32_t c[C];
#pragma unroll
        for (int g = 0; g < C / 4; ++g) {
            uint4 q = make_uint4(0, 0, 0, 0);
            if (i0 + 4 * g < e) q = ld_stream(reinterpret_cast<const uint4*>(ja + i0 + 4 * g));  // overruns e by < 4: inside the arrays' slack
            c[4 * g] = q.x; c[4 * g + 1] = q.y; c[4 * g + 2] = q.z; c[4 * g + 3] = q.w;
        }
#pragma unroll
        for (int g = 0; g < C / 2; ++g) {
            double2 v = make_double2(0.0, 0.0);
            if (i0 + 2 * g < e) v = ld_stream(reinterpret_cast<const double2*>(as + i0 + 2 * g));
            p[2 * g] = v.x; p[2 * g + 1] = v.y;
        }
        // Masked entries (before s in the first chunk, from e on in the last) become +0.0 and are added like the rest: the running sum
        // starts at +0.0 and a round-to-nearest sum is -0.0 only if both operands are, so it is never -0.0 and "+ 0.0" is an exact no-op.
        // One unpredicated chain for every chunk (a predicated tail loop put compare + select latency into the chain: 225 us).
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const bool ok = i0 + j >= s && i0 + j < e;
            p[j] = ok ? __dmul_rn(p[j], ld_x(x, c[j])) : 0.0;
        }
        const uint32_t nl = (min(e, base + CH) - base + C - 1) / C;  // lanes that hold anything (warp-uniform)
#pragma unroll 4
        for (uint32_t l = 0; l < nl; ++l) {
            double t = acc;
#pragma unroll
            for (int j = 0; j < C; ++j) t = __dadd_rn(t, p[j]);
            acc = __shfl_sync(0xffffffffu, t, l);
        }
    }
    if (lane == 0) y[row] = acc;
}

// largest column id in a range of a column-id array / in the valid slots of a row range of a column-major ELL
// (plan of the pipelined host path: which x pieces a row chunk needs)
__global__ void colmax_flat_kernel(const uint32_t* __restrict__ ja, uint64_t n0, uint64_t n1, uint32_t* __restrict__ out) {
    uint32_t mx = 0;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = n0 + (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += stride) mx = max(mx, ja[i]);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}
__global__ void colmax_ell_cm_kernel(const uint32_t* __restrict__ ja, const uint32_t* __restrict__ rl, uint64_t pitch, uint32_t r0,
                                     uint32_t r1, uint32_t* __restrict__ out) {
    const uint32_t r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mx = 0;
    if (r < r1) {
        const uint32_t len = rl[r];
        for (uint32_t k = 0; k < len; ++k) mx = max(mx, ja[(uint64_t) k * pitch + r]);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}

__global__ void widen_u32_kernel(const uint32_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n, uint64_t add) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (uint64_t) src[i] + add;
}

}  // namespace spmvb200
