// xwin.cuh -- "x-window" CSR: the x gather served from SHARED MEMORY instead of L2.
//
// Why: on B200 a gather that misses L1 costs one 32-byte L2 sector request, and the L2 sustains only ~1 request
// per clock per SM (tools/microbench_gather.cu) -- a matrix whose columns are scattered over a window larger than
// L1 (cfg4: 32 columns in [i-2^15, i+2^15] = 512 KB of x per row) is bound at ~0.5 of the HBM roofline whatever the
// matrix stream does.  Shared memory gathers run at 3-5 per clock per SM.  So the matrix is re-tiled at build time:
//
//   row block  b = rows [b*R, (b+1)*R)                  one CTA; R = 32 * NW * ACC
//   window     w = columns [w*W, (w+1)*W)               W*8 bytes of x, staged by ONE TMA bulk copy
//   tile       (b, w) with at least one non-zero        tiles of a row block are stored in ascending window order
//
// Inside a tile the rows are cut into groups of 32 (a warp, lane = row) and each group is stored JAGGED slot-major:
// first the 1st non-zero (in this window) of every row that has one, then the 2nd of every row that has two, ...
// with no padding at all.  Lane l finds "its" entry of slot k at  off + popc(ballot(cnt > k) & lanemask_lt)  and the
// warp's loads are always contiguous (<= 256 B of values, <= 64 B of 16-bit window-local column ids).
// Per row and tile the format spends one byte (cnt), per group and tile four (offset): 10.1-10.4 B per non-zero
// instead of CSR's 12 -- the kernel moves fewer bytes than the "algorithmic" 12*nnz it is scored against.
//
// Kernel: NW warps.  The row block's x windows stream through a ring of NBUF shared-memory buffers (one mbarrier per
// slot for "landed", one counter per slot for "every warp has left", no CTA-wide barrier in the loop); lanes keep their
// rows' sums in registers across all tiles, adding products in ascending column order with separate mul/add
// roundings => bit-identical to sgemvSerial (src/SpMV_CSR_OMP.c:229-250) when the CSR rows are column-sorted.
#pragma once
#include "common.cuh"

namespace spmvb200 {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t ld_stream(const uint16_t* p) { return (uint32_t) __ldcs(reinterpret_cast<const unsigned short*>(p)); }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

constexpr int XW_MAX_NBUF = 8;

// One slot of one 32-row group, load half: lanes with cnt > k fetch their entry of the jagged slot
//   idx = off + popc(ballot(cnt > k) & lanemask_lt),  off += popc(ballot)
// with the two loads predicated (no branch, no zero fill).  Written in PTX so that the batch of U x ACC of these is
// emitted back to back (asm volatile keeps program order) ahead of the first use: 11 instructions per slot and group.
__device__ __forceinline__ void xw_slot_load(double& v, uint32_t& cc, uint32_t& off, const double* __restrict__ val,
                                             const uint16_t* __restrict__ col, uint32_t c, uint32_t k, uint32_t lt) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        ".reg .b32 m, r;\n"
        ".reg .b64 pa, pb;\n"
        "setp.gt.u32 q, %5, %6;\n"
        "vote.sync.ballot.b32 m, q, 0xffffffff;\n"
        "and.b32 r, m, %7;\n"
        "popc.b32 r, r;\n"
        "add.u32 r, r, %2;\n"
        "popc.b32 m, m;\n"
        "add.u32 %2, %2, m;\n"
        "mad.wide.u32 pa, r, 8, %3;\n"
        "mad.wide.u32 pb, r, 2, %4;\n"
        "mov.f64 %0, 0d0000000000000000;\n"
        "@q ld.global.cs.f64 %0, [pa];\n"
        "@q ld.global.cs.u16 %1, [pb];\n"
        "}\n"
        : "+d"(v), "+r"(cc), "+r"(off)
        : "l"(val), "l"(col), "r"(c), "r"(k), "r"(lt)
        : "memory");
}
// consume half: acc += v * xw[cc] with separate mul / add roundings (the order and rounding of sgemvSerial,
// src/SpMV_CSR_OMP.c:229-250).  Lanes with cnt <= k add (+0) * (+0): exact, since a running sum that starts at +0 is
// never -0; this keeps the two fp64 instructions unpredicated (ptxas turns predicated fp64 math into selects).
__device__ __forceinline__ void xw_slot_fma(double& acc, double v, uint32_t cc, uint32_t c, uint32_t k, uint32_t xw_saddr) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        ".reg .f64 x, p;\n"
        ".reg .b32 sa;\n"
        "setp.gt.u32 q, %3, %4;\n"
        "mad.lo.u32 sa, %2, 8, %5;\n"
        "mov.f64 x, 0d0000000000000000;\n"
        "@q ld.shared.f64 x, [sa];\n"
        "mul.rn.f64 p, %1, x;\n"
        "add.rn.f64 %0, %0, p;\n"
        "}\n"
        : "+d"(acc)
        : "d"(v), "r"(cc), "r"(c), "r"(k), "r"(xw_saddr)
        : "memory");
}

// x window of tile t -> ring slot s.  Called by one whole warp.  Aligned x: lane 0 issues one TMA bulk copy that
// completes on full[s]; otherwise (caller's x not 16-byte aligned) the warp copies the window itself.
__device__ __forceinline__ void xw_load_window(const double* __restrict__ x, uint32_t win, uint32_t W, uint32_t N, double* dst, uint64_t* bar,
                                               int x_aligned, uint32_t lane, uint64_t pol) {
    const uint64_t base = (uint64_t) win * W;
    const uint32_t n = (uint32_t) min((uint64_t) W, (uint64_t) N - base);
    if (x_aligned) {
        if (lane == 0) {
            const uint32_t n2 = n & ~1u;  // TMA moves multiples of 16 bytes; an odd tail element goes by hand
            if (n & 1u) dst[n - 1] = __ldg(x + base + n - 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this slot vs the async write
            mbar_arrive_expect_tx(bar, n2 * 8u);
            if (n2) bulk_g2s(dst, x + base, n2 * 8u, bar, pol);
        }
    } else {
        for (uint32_t j = lane; j < n; j += 32) dst[j] = __ldg(x + base + j);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    }
}

// No producer warp and no CTA-wide barrier in the loop: every warp counts itself out of a ring slot (shared-memory
// atomic, acq_rel); the LAST warp to leave slot s refills it with the window of tile t + nbuf.
// Latency: the matrix stream lands in registers (shared memory belongs to the x windows).  A warp issues the loads of
// U slots x ACC groups (2*U*ACC independent loads per lane) before it consumes the first, and fetches the next tile's
// counts/offsets while it works on the current tile.
// DBG (developer builds only): bit 0 = no x windows (no TMA, no waits), bit 1 = no shared-memory gather.
template <int NW, int ACC, int U, int DBG = 0>
__global__ void __launch_bounds__(32 * NW, 1)
xwin_kernel(const uint32_t* __restrict__ rb_tile0, const uint32_t* __restrict__ tile_win, const uint32_t* __restrict__ grp_off,
            const uint8_t* __restrict__ cnt, const uint16_t* __restrict__ col, const double* __restrict__ val,
            const double* __restrict__ x, double* __restrict__ y, uint32_t M, uint32_t N, uint32_t W, uint32_t nbuf, int x_aligned) {
    constexpr uint32_t R = 32u * NW * ACC, G = NW * ACC;
    extern __shared__ __align__(128) unsigned char xw_smem[];
    double* xs = reinterpret_cast<double*>(xw_smem);                       // nbuf windows of W doubles
    uint64_t* full = reinterpret_cast<uint64_t*>(xs + (size_t) nbuf * W);  // [nbuf] window landed
    uint32_t* done = reinterpret_cast<uint32_t*>(full + XW_MAX_NBUF);      // [nbuf] warps that have left the slot (monotonic)

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x;
    const uint32_t t0 = __ldg(rb_tile0 + b), t1 = __ldg(rb_tile0 + b + 1);
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nbuf; ++s) {
            mbar_init(full + s, 1);
            done[s] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();
    const uint64_t pol = policy_evict_last();
    if (!(DBG & 1) && warp == 0)
        for (uint32_t i = 0; i < nbuf && t0 + i < t1; ++i)
            xw_load_window(x, __ldg(tile_win + t0 + i), W, N, xs + (size_t) i * W, full + i, x_aligned, lane, pol);

    // warp owns row groups g = a*NW + warp (a < ACC), lane = row inside the group
    double acc[ACC];
    uint32_t c_nx[ACC], off_nx[ACC];
#pragma unroll
    for (int a = 0; a < ACC; ++a) {
        acc[a] = 0.0;
        c_nx[a] = 0;
        off_nx[a] = 0;
        if (t0 < t1) {
            const uint32_t g = a * NW + warp;
            c_nx[a] = __ldg(cnt + (size_t) t0 * R + g * 32u + lane);
            off_nx[a] = __ldg(grp_off + (size_t) t0 * G + g);
        }
    }
    const uint32_t lt = lanemask_lt();
    uint32_t s = 0, ph = 0;
    for (uint32_t t = t0; t < t1; ++t) {
        uint32_t c[ACC], off[ACC], kmax = 0;
#pragma unroll
        for (int a = 0; a < ACC; ++a) {
            c[a] = c_nx[a];
            off[a] = off_nx[a];
            kmax = max(kmax, c[a]);
        }
        if (t + 1 < t1) {  // next tile's metadata: in flight while this tile is processed
#pragma unroll
            for (int a = 0; a < ACC; ++a) {
                const uint32_t g = a * NW + warp;
                c_nx[a] = __ldg(cnt + (size_t) (t + 1) * R + g * 32u + lane);
                off_nx[a] = __ldg(grp_off + (size_t) (t + 1) * G + g);
            }
        }
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if (!(DBG & 1)) mbar_wait(full + s, ph);
        const uint32_t xw = smem_u32(xs + (size_t) s * W);
        double v[U][ACC];
        uint32_t cc[U][ACC];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int a = 0; a < ACC; ++a) { v[u][a] = 0.0; cc[u][a] = 0u; }
#pragma unroll 1
        for (uint32_t k0 = 0; k0 < kmax; k0 += U) {  // U slots x ACC groups in flight before the first use; slots past a row's count are predicated off
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int a = 0; a < ACC; ++a) xw_slot_load(v[u][a], cc[u][a], off[a], val, col, c[a], k0 + u, lt);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int a = 0; a < ACC; ++a) {
                    if (DBG & 2) acc[a] += v[u][a] * (double) cc[u][a];
                    else xw_slot_fma(acc[a], v[u][a], cc[u][a], c[a], k0 + u, xw);
                }
        }
        if (!(DBG & 1)) {
            __syncwarp();
            uint32_t last = 0;
            if (lane == 0) {
                uint32_t prev;
                asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(smem_u32(done + s)) : "memory");
                last = ((prev + 1u) % NW) == 0u;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last && t + nbuf < t1)
                xw_load_window(x, __ldg(tile_win + t + nbuf), W, N, xs + (size_t) s * W, full + s, x_aligned, lane, pol);
        }
        if (++s == nbuf) { s = 0; ph ^= 1u; }
    }
#pragma unroll
    for (int a = 0; a < ACC; ++a) {
        const uint32_t row = b * R + (a * NW + warp) * 32u + lane;
        if (row < M) y[row] = acc[a];
    }
}

// ---------------------------------------------------------------------------------------------
// construction on the device (once per matrix): mark -> scan -> tiles -> count -> scan -> fill
// ---------------------------------------------------------------------------------------------
// bit (rb, w) of the bitmap = some row of row block rb has a non-zero in window w; also flags unsorted rows
__global__ void xw_mark_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, uint32_t M, uint32_t R, uint32_t W,
                               uint32_t nwords, uint32_t* __restrict__ bitmap, int* __restrict__ unsorted) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const uint32_t s = irp[r], e = irp[r + 1];
    uint32_t* bm = bitmap + (size_t) (r / R) * nwords;
    uint32_t prev_w = 0xffffffffu, prev_c = 0;
    for (uint32_t j = s; j < e; ++j) {
        const uint32_t c = ja[j], w = c / W;
        if (j > s && c < prev_c) *unsorted = 1;
        prev_c = c;
        if (w != prev_w) {
            atomicOr(bm + (w >> 5), 1u << (w & 31));
            prev_w = w;
        }
    }
}
__global__ void xw_popc_kernel(const uint32_t* __restrict__ bitmap, uint64_t n, uint32_t* __restrict__ pc) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pc[i] = __popc(bitmap[i]);
    if (i == n) pc[i] = 0;
}
// tile list: word (rb, i) of the bitmap owns tiles [scan[rb*nwords+i], ...) in ascending window order
__global__ void xw_tiles_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ scan, uint32_t nrb, uint32_t nwords,
                                uint32_t* __restrict__ rb_tile0, uint32_t* __restrict__ tile_win, uint32_t* __restrict__ tile_rb) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x, n = (uint64_t) nrb * nwords;
    if (i > n) return;
    if (i == n) { rb_tile0[nrb] = scan[n]; return; }
    const uint32_t rb = (uint32_t) (i / nwords), wi = (uint32_t) (i % nwords);
    uint32_t t = scan[i], bits = bitmap[i];
    if (wi == 0) rb_tile0[rb] = t;
    while (bits) {
        const uint32_t bit = __ffs(bits) - 1;
        bits &= bits - 1;
        tile_win[t] = wi * 32 + bit;
        tile_rb[t] = rb;
        ++t;
    }
}
// first position in [s, e) of a sorted row with ja >= key
__device__ __forceinline__ uint32_t xw_lower_bound(const uint32_t* __restrict__ ja, uint32_t s, uint32_t e, uint64_t key) {
    while (s < e) {
        const uint32_t mid = s + (e - s) / 2;
        if ((uint64_t) ja[mid] < key) s = mid + 1; else e = mid;
    }
    return s;
}
// one warp per (tile, 32-row group): per-row entry count inside the tile's window, group total
template <bool SORTED>
__global__ void xw_count_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, uint32_t M, uint32_t R, uint32_t W,
                                const uint32_t* __restrict__ tile_win, const uint32_t* __restrict__ tile_rb, uint64_t ngroups,
                                uint8_t* __restrict__ cnt, uint32_t* __restrict__ grp_cnt, int* __restrict__ overflow) {
    const uint64_t wg = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31, G = R / 32;
    if (wg > ngroups) return;
    if (wg == ngroups) { if (lane == 0) grp_cnt[wg] = 0; return; }
    const uint32_t t = (uint32_t) (wg / G), g = (uint32_t) (wg % G);
    const uint32_t row = tile_rb[t] * R + g * 32 + lane;
    const uint64_t lo_c = (uint64_t) tile_win[t] * W, hi_c = lo_c + W;
    uint32_t c = 0;
    if (row < M) {
        const uint32_t s = irp[row], e = irp[row + 1];
        if (SORTED) {
            const uint32_t lo = xw_lower_bound(ja, s, e, lo_c);
            c = xw_lower_bound(ja, lo, e, hi_c) - lo;
        } else {
            for (uint32_t j = s; j < e; ++j) c += (ja[j] >= lo_c && ja[j] < hi_c);
        }
    }
    if (c > 255u) { *overflow = 1; c = 255u; }
    cnt[(size_t) t * R + g * 32 + lane] = (uint8_t) c;
    const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) grp_cnt[wg] = tot;
}
// one warp per (tile, group): jagged slot-major fill (the layout xwin_kernel reads)
template <bool SORTED>
__global__ void xw_fill_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, const double* __restrict__ as, uint32_t M,
                               uint32_t R, uint32_t W, const uint32_t* __restrict__ tile_win, const uint32_t* __restrict__ tile_rb,
                               uint64_t ngroups, const uint8_t* __restrict__ cnt, const uint32_t* __restrict__ grp_off,
                               uint16_t* __restrict__ col, double* __restrict__ val) {
    const uint64_t wg = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31, G = R / 32;
    if (wg >= ngroups) return;
    const uint32_t t = (uint32_t) (wg / G), g = (uint32_t) (wg % G);
    const uint32_t row = tile_rb[t] * R + g * 32 + lane;
    const uint64_t lo_c = (uint64_t) tile_win[t] * W, hi_c = lo_c + W;
    const uint32_t c = cnt[(size_t) t * R + g * 32 + lane];
    uint32_t j = 0, e = 0;
    if (row < M && c) {
        j = irp[row];
        e = irp[row + 1];
        if (SORTED) j = xw_lower_bound(ja, j, e, lo_c);
    }
    uint32_t off = grp_off[wg];
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, c), lt = lanemask_lt();
    for (uint32_t k = 0; k < kmax; ++k) {
        const bool on = c > k;
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        if (on) {
            if (!SORTED) while (j < e && !(ja[j] >= lo_c && ja[j] < hi_c)) ++j;
            const uint32_t idx = off + __popc(m & lt);
            val[idx] = as[j];
            col[idx] = (uint16_t) ((uint64_t) ja[j] - lo_c);
            ++j;
        }
        off += __popc(m);
    }
}

}  // namespace spmvb200
