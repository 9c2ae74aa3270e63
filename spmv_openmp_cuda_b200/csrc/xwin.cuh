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
// Inside a tile the rows are cut into groups of 32 (a warp, lane = row) and each group is stored JAGGED slot-major and
// SORTED: slot k holds the k-th non-zero (in this window) of every row that has one, rows ordered by decreasing count,
// with no padding at all.  A row's place p in that order is stored with its count (one u16 per row and tile), so the
// rows active in slot k are exactly the places [0, n_k) and lane l finds its entry at  (group base + S_k) + p_l  with
// S_k = n_0 + ... + n_{k-1} warp-uniform: per slot the kernel spends ONE ballot + popc and advances one per-lane
// entry index -- no per-slot rank (the SpMV is instruction-issue bound long before it is HBM
// bound if every non-zero costs a rank computation: measured 1.9 warp instructions per non-zero, 58 % issue
// utilisation at 0.6 of the roofline for the unsorted layout).  The warp's loads stay contiguous (<= 256 B of values,
// <= 64 B of 16-bit window-local column ids); lanes keep their own rows (sums stay in registers across tiles).
// Per row and tile the format spends two bytes, per group and tile four: 10.3-10.7 B per non-zero instead of CSR's 12.
//
// Kernel: NW warps.  The row block's x windows stream through a ring of NBUF shared-memory buffers (one mbarrier per
// slot for "landed", one counter per slot for "every warp has left", no CTA-wide barrier in the loop); lanes keep their
// rows' sums in registers across all tiles, adding products in ascending column order with separate mul/add
// roundings => bit-identical to sgemvSerial (src/SpMV_CSR_OMP.c:229-250) when the CSR rows are column-sorted.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace spmvb200 {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

constexpr int XW_MAX_NBUF = 8;

// Where the (count | place << 8) word of row (group g, lane) of tile t lives.  A warp owns the groups a * nw + warp (a < R / (32 nw)): the
// words of one lane are stored next to each other, so that the kernel fetches them with ONE load per lane and tile (two u16 = one u32
// for the default shape).
__host__ __device__ __forceinline__ size_t xw_cp_index(uint32_t t, uint32_t g, uint32_t lane, uint32_t R, uint32_t nw) {
    const uint32_t acc = R / (32u * nw), a = g / nw, warp = g % nw;
    return (size_t) t * R + (size_t) (warp * 32u + lane) * acc + a;
}

// One slot of one 32-row group, load half.  Lanes with cnt > k fetch their entry, entry index e (ONE 32-bit index per group serves the
// value and the id array: two address instructions per slot more than two running 64-bit pointers, six registers per lane less -- this
// kernel does not pay for ALU instructions, it pays for registers: -3 % on cfg4, profiles/r02ae_*, r02ag_*), which then advances by the
// slot's population n_k = popc(ballot(cnt > k)).  Inactive lanes keep v = 0 and cc = W (the index of a zero kept behind every x window),
// so the consume half needs no predicate.  PTX so that a batch of these is emitted back to back ahead of the first use (asm volatile
// keeps program order).
__device__ __forceinline__ void xw_slot_load(double& v, uint32_t& cc, uint32_t& e, const double* val, const uint16_t* col, uint32_t c,
                                             uint32_t k, uint32_t W) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        ".reg .b32 m;\n"
        ".reg .b64 a;\n"
        "setp.gt.u32 q, %5, %6;\n"
        "vote.sync.ballot.b32 m, q, 0xffffffff;\n"
        "popc.b32 m, m;\n"
        "mov.f64 %0, 0d0000000000000000;\n"
        "mov.b32 %1, %7;\n"
        "mad.wide.u32 a, %2, 8, %3;\n"
        "@q ld.global.cs.f64 %0, [a];\n"
        "mad.wide.u32 a, %2, 2, %4;\n"
        "@q ld.global.cs.u16 %1, [a];\n"
        "add.u32 %2, %2, m;\n"
        "}\n"
        : "=d"(v), "=r"(cc), "+r"(e)
        : "l"(val), "l"(col), "r"(c), "r"(k), "r"(W)
        : "memory");
}
// consume half: acc += v * xw[cc] with separate mul / add roundings (the order and rounding of sgemvSerial,
// src/SpMV_CSR_OMP.c:229-250).  Inactive lanes add (+0) * (+0): exact, since a running sum that starts at +0 is never -0.
__device__ __forceinline__ void xw_slot_fma(double& acc, double v, uint32_t cc, uint32_t xw_saddr) {
    asm volatile(
        "{\n"
        ".reg .f64 x, p;\n"
        ".reg .b32 sa;\n"
        "mad.lo.u32 sa, %2, 8, %3;\n"
        "ld.shared.f64 x, [sa];\n"
        "mul.rn.f64 p, %1, x;\n"
        "add.rn.f64 %0, %0, p;\n"
        "}\n"
        : "+d"(acc)
        : "d"(v), "r"(cc), "r"(xw_saddr)
        : "memory");
}

// x window of tile t -> ring slot s (slots are W+2 doubles apart: index W holds the zero inactive lanes read).
// Called by one whole warp.  Aligned x: lane 0 issues one TMA bulk copy that completes on full[s]; otherwise
// (caller's x not 16-byte aligned) the warp copies the window itself.
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

// U slots of all ACC groups: loads first (2*U*ACC independent loads per lane in flight), then the gathers and sums.
// The first batch of a tile waits for the tile's x window only AFTER its loads are out: window and matrix latencies
// overlap.
template <int ACC, int U>
__device__ __forceinline__ void xw_batch(double (&acc)[ACC], uint32_t (&e)[ACC], const double* __restrict__ val, const uint16_t* __restrict__ col,
                                         const uint32_t (&c)[ACC], uint32_t k0, uint32_t W, uint32_t xw, uint64_t* bar, uint32_t ph) {
    double v[U][ACC];
    uint32_t cc[U][ACC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int a = 0; a < ACC; ++a) xw_slot_load(v[u][a], cc[u][a], e[a], val, col, c[a], k0 + u, W);
    if (k0 == 0) mbar_wait(bar, ph);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int a = 0; a < ACC; ++a) xw_slot_fma(acc[a], v[u][a], cc[u][a], xw);
}

// the ACC 16-bit (count | place << 8) words of one lane, fetched with one load
template <int ACC> struct XwCp;
template <> struct XwCp<1> {
    uint32_t w;
    __device__ __forceinline__ void clear() { w = 0; }
    __device__ __forceinline__ void load(const uint16_t* p) { w = __ldg(p); }
    __device__ __forceinline__ uint32_t get(int) const { return w; }
};
template <> struct XwCp<2> {
    uint32_t w;
    __device__ __forceinline__ void clear() { w = 0; }
    __device__ __forceinline__ void load(const uint16_t* p) { w = __ldg(reinterpret_cast<const uint32_t*>(p)); }
    __device__ __forceinline__ uint32_t get(int a) const { return a ? w >> 16 : w & 0xffffu; }
};
template <> struct XwCp<4> {
    uint2 w;
    __device__ __forceinline__ void clear() { w = make_uint2(0u, 0u); }
    __device__ __forceinline__ void load(const uint16_t* p) { w = __ldg(reinterpret_cast<const uint2*>(p)); }
    __device__ __forceinline__ uint32_t get(int a) const { const uint32_t v = a < 2 ? w.x : w.y; return (a & 1) ? v >> 16 : v & 0xffffu; }
};
template <> struct XwCp<8> {
    uint4 w;
    __device__ __forceinline__ void clear() { w = make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ void load(const uint16_t* p) { w = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ uint32_t get(int a) const {
        const uint32_t v = a < 2 ? w.x : a < 4 ? w.y : a < 6 ? w.z : w.w;
        return (a & 1) ? v >> 16 : v & 0xffffu;
    }
};

// Does the CTA that owns row blocks [rb, rb1) take part in the fused neighbour synchronisation?  Yes if some of its rows are
// delivered to a peer, or if one of its x windows reaches outside the columns this rank owns (windows are ascending per row block).
__device__ __forceinline__ bool xw_cta_is_boundary(const PushArgs& push, uint32_t rb, uint32_t rb1, const uint32_t* __restrict__ rb_tile0,
                                                   const uint32_t* __restrict__ tile_win, uint32_t R, uint32_t W, uint32_t M, uint32_t N) {
    const uint64_t row_hi = min((uint64_t) rb1 * R, (uint64_t) M);
    if (push_rows_wanted(push, rb * R, (uint32_t) row_hi)) return true;
    for (uint32_t b = rb; b < rb1; ++b) {
        const uint32_t t0 = __ldg(rb_tile0 + b), t1 = __ldg(rb_tile0 + b + 1);
        if (t1 <= t0) continue;
        const uint64_t cmin = (uint64_t) __ldg(tile_win + t0) * W, cmax = min(((uint64_t) __ldg(tile_win + t1 - 1) + 1) * W, (uint64_t) N);
        if (cmin < push.own_lo || cmax > push.own_hi) return true;
    }
    return false;
}

// No producer warp and no CTA-wide barrier in the loop: every warp counts itself out of a ring slot (shared-memory
// atomic, acq_rel); the LAST warp to leave slot s refills it with the window of tile t + nbuf.
// Latency: the matrix stream lands in registers (shared memory belongs to the x windows).  A warp issues the loads of
// a batch of up to UMAX slots x ACC groups before it consumes the first -- the batch size is chosen per tile from the
// warp's longest row in that tile (switch over fully unrolled bodies), so that short tiles neither execute
// predicated-off slots nor leave a lone tail slot exposed to a full memory round trip -- and fetches the next tile's
// counts/offsets while it works on the current tile.
// CTA i owns the contiguous row blocks [cta_rb[i], cta_rb[i+1]) and runs their tiles as ONE sequence: tile ids are
// global and consecutive across row blocks, so the window ring and the metadata prefetch keep going at a row-block
// boundary; only the sums are written out and reset there.  One row block per CTA is the plain launch.
template <int NW, int ACC, int UMAX>
__global__ void __launch_bounds__(32 * NW, 1)
xwin_kernel(const uint32_t* __restrict__ cta_rb, const uint32_t* __restrict__ rb_tile0, const uint32_t* __restrict__ tile_win,
            const uint32_t* __restrict__ grp_off, const uint16_t* __restrict__ cp, const uint16_t* __restrict__ col,
            const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y, uint32_t M, uint32_t N, uint32_t W,
            uint32_t nbuf, int x_aligned, uint32_t rb_first, const PushArgs push) {
    constexpr uint32_t R = 32u * NW * ACC, G = NW * ACC;
    extern __shared__ __align__(128) unsigned char xw_smem[];
    double* xs = reinterpret_cast<double*>(xw_smem);                              // nbuf windows of W (+2) doubles
    const uint32_t WS = W + 2;                                                    // slot stride; xs[s*WS + W] = 0
    uint64_t* full = reinterpret_cast<uint64_t*>(xs + (size_t) nbuf * WS);        // [nbuf] window landed
    uint32_t* done = reinterpret_cast<uint32_t*>(full + XW_MAX_NBUF);             // [nbuf] warps that have left the slot (monotonic)

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // no split table: one row block per CTA (rotated under the fused neighbour synchronisation, see PushArgs::rb_rot)
    uint32_t rb = cta_rb ? __ldg(cta_rb + blockIdx.x) : rb_first + (push.rb_rot ? (blockIdx.x + push.rb_rot) % gridDim.x : blockIdx.x);
    const uint32_t rb1 = cta_rb ? __ldg(cta_rb + blockIdx.x + 1) : rb + 1;
    if (rb >= rb1) return;
    const uint32_t T0 = __ldg(rb_tile0 + rb), T1 = __ldg(rb_tile0 + rb1);  // this CTA's tiles
    __shared__ uint32_t s_boundary, s_delivers;
    if (threadIdx.x == 0) {
        // does any row of this CTA go to a peer?  Decided once per CTA (two compares per destination, no loads): the per-row range
        // tests of push_out in the epilogue of EVERY row were what the fused delivery cost (+2-3 % at 2 GPUs), not the stores
        s_delivers = push.n && push_rows_wanted(push, rb * R, (uint32_t) min((uint64_t) rb1 * R, (uint64_t) M));
        // fused neighbour synchronisation (common.cuh): only CTAs that deliver rows or read columns owned by a peer take part
        const bool boundary = push.nsync && push.cta_boundary[blockIdx.x];
        if (boundary) push_sync_wait(push);
        s_boundary = boundary;
        for (uint32_t s = 0; s < nbuf; ++s) {
            mbar_init(full + s, 1);
            done[s] = 0;
            xs[(size_t) s * WS + W] = 0.0;
            xs[(size_t) s * WS + W + 1] = 0.0;
        }
        mbar_fence_init();
    }
    __syncthreads();
    const bool delivers = s_delivers != 0;
    const uint64_t pol = policy_evict_last();
    if (warp == 0)
        for (uint32_t i = 0; i < nbuf && T0 + i < T1; ++i)
            xw_load_window(x, __ldg(tile_win + T0 + i), W, N, xs + (size_t) i * WS, full + i, x_aligned, lane, pol);

    // warp owns row groups g = a*NW + warp (a < ACC) of every row block, lane = row inside the group
    const uint32_t myrow = warp * 32u + lane;  // + rb*R + a*NW*32
    double acc[ACC];
    XwCp<ACC> cp_nx;  // my ACC rows' (count | place << 8) words of the next tile: one load (xw_cp_index)
    uint32_t off_nx[ACC];
    cp_nx.clear();
#pragma unroll
    for (int a = 0; a < ACC; ++a) {
        acc[a] = 0.0;
        off_nx[a] = 0;
        if (T0 < T1) off_nx[a] = __ldg(grp_off + (size_t) T0 * G + a * NW + warp);
    }
    if (T0 < T1) cp_nx.load(cp + (size_t) T0 * R + (size_t) myrow * ACC);
    uint32_t t_end = __ldg(rb_tile0 + rb + 1);
    // row blocks without tiles at the start of the range
    while (t_end == T0 && rb < rb1) {
#pragma unroll
        for (int a = 0; a < ACC; ++a) {
            const uint32_t row = rb * R + a * NW * 32u + myrow;
            if (row < M) {
                y[row] = 0.0;
                if (delivers) push_out(push, row, 0.0);
            }
        }
        if (++rb < rb1) t_end = __ldg(rb_tile0 + rb + 1);
    }
    uint32_t s = 0, ph = 0;
#pragma unroll 1
    for (uint32_t t = T0; t < T1; ++t) {
        uint32_t c[ACC], e[ACC], kmax = 0;
#pragma unroll
        for (int a = 0; a < ACC; ++a) {
            c[a] = cp_nx.get(a) & 0xffu;             // entries of my row in this tile
            e[a] = off_nx[a] + (cp_nx.get(a) >> 8);  // my first entry: group base + my place in the sorted order
            kmax = max(kmax, c[a]);
        }
        if (t + 1 < T1) {  // next tile's metadata (possibly the next row block's): in flight while this tile is processed
            cp_nx.load(cp + (size_t) (t + 1) * R + (size_t) myrow * ACC);
#pragma unroll
            for (int a = 0; a < ACC; ++a) off_nx[a] = __ldg(grp_off + (size_t) (t + 1) * G + a * NW + warp);
        }
        // the window id the LAST warp to leave this tile will refill the slot with: fetched by every warp now, so that the refill is one
        // TMA issue away when the slot frees up instead of a dependent load + a TMA issue (the ring's turn-around time bounds the tile rate
        // when a row block has more tiles than the ring has slots)
        const uint32_t win_nx = (t + nbuf < T1) ? __ldg(tile_win + t + nbuf) : 0u;
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        const uint32_t xw = smem_u32(xs + (size_t) s * WS);
        if (kmax == 0) mbar_wait(full + s, ph);  // nothing of this warp's rows here: still keep step with the ring
        uint32_t k0 = 0;
#pragma unroll 1
        while (k0 < kmax) {
            const uint32_t rem = kmax - k0;
            if (rem >= (uint32_t) UMAX) {
                xw_batch<ACC, UMAX>(acc, e, val, col, c, k0, W, xw, full + s, ph);
                k0 += UMAX;
            } else {
                switch (rem) {  // exact tail (or whole short tile) in one batch
                    case 1: xw_batch<ACC, 1>(acc, e, val, col, c, k0, W, xw, full + s, ph); break;
                    case 2: xw_batch<ACC, (UMAX > 2 ? 2 : 1)>(acc, e, val, col, c, k0, W, xw, full + s, ph); break;
                    case 3: xw_batch<ACC, (UMAX > 3 ? 3 : 1)>(acc, e, val, col, c, k0, W, xw, full + s, ph); break;
                    case 4: xw_batch<ACC, (UMAX > 4 ? 4 : 1)>(acc, e, val, col, c, k0, W, xw, full + s, ph); break;
                    case 5: xw_batch<ACC, (UMAX > 5 ? 5 : 1)>(acc, e, val, col, c, k0, W, xw, full + s, ph); break;
                    case 6: xw_batch<ACC, (UMAX > 6 ? 6 : 1)>(acc, e, val, col, c, k0, W, xw, full + s, ph); break;
                    default: xw_batch<ACC, (UMAX > 7 ? 7 : 1)>(acc, e, val, col, c, k0, W, xw, full + s, ph); break;
                }
                k0 = kmax;
            }
        }
        __syncwarp();
        uint32_t last = 0;
        if (lane == 0) {
            uint32_t prev;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(smem_u32(done + s)) : "memory");
            last = ((prev + 1u) % NW) == 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && t + nbuf < T1)
            xw_load_window(x, win_nx, W, N, xs + (size_t) s * WS, full + s, x_aligned, lane, pol);
        if (++s == nbuf) { s = 0; ph ^= 1u; }
        while (t + 1 == t_end && rb < rb1) {  // row block finished (and any tile-less row blocks after it)
#pragma unroll
            for (int a = 0; a < ACC; ++a) {
                const uint32_t row = rb * R + a * NW * 32u + myrow;
                if (row < M) {
                    y[row] = acc[a];
                    if (delivers) push_out(push, row, acc[a]);
                }
                acc[a] = 0.0;
            }
            if (++rb < rb1) t_end = __ldg(rb_tile0 + rb + 1);
        }
    }
    if (push.nsync && s_boundary) {  // CTA-uniform
        __syncthreads();
        if (threadIdx.x == 0) push_sync_signal(push);
    }
}

// how many CTAs of a launch (same grid shape as xwin_kernel) are boundary CTAs: counted once per handle, mode and partition
__global__ void xw_count_boundary_kernel(const uint32_t* __restrict__ cta_rb, const uint32_t* __restrict__ rb_tile0, const uint32_t* __restrict__ tile_win,
                                         uint32_t ncta, uint32_t R, uint32_t W, uint32_t M, uint32_t N, const PushArgs push, uint32_t* __restrict__ count,
                                         uint8_t* __restrict__ cta_flag) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncta) return;
    const uint32_t rb = cta_rb ? cta_rb[c] : (push.rb_rot ? (c + push.rb_rot) % ncta : c), rb1 = cta_rb ? cta_rb[c + 1] : rb + 1;
    const bool b = rb < rb1 && xw_cta_is_boundary(push, rb, rb1, rb_tile0, tile_win, R, W, M, N);
    cta_flag[c] = b ? 1 : 0;
    if (b) atomicAdd(count, 1u);
}

// row blocks per persistent CTA, balanced by non-zeros: CTA i starts at the first row block whose first entry index
// reaches i * NZ / ncta  (grp_off[rb_tile0[rb] * G] = entries before row block rb)
__global__ void xw_cta_split_kernel(const uint32_t* __restrict__ rb_tile0, const uint32_t* __restrict__ grp_off, uint32_t G, uint32_t nrb,
                                    uint64_t NZ, uint32_t ncta, uint32_t* __restrict__ cta_rb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > ncta) return;
    if (i == 0) { cta_rb[0] = 0; return; }
    if (i == ncta) { cta_rb[ncta] = nrb; return; }
    const uint64_t target = NZ * i / ncta;
    uint32_t lo = 0, hi = nrb;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((uint64_t) grp_off[(size_t) rb_tile0[mid] * G] < target) lo = mid + 1; else hi = mid;
    }
    cta_rb[i] = lo;
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
// one warp per (tile, 32-row group): per-row entry count inside the tile's window, the row's place in the group's
// order by decreasing count (ties by lane), group total
template <bool SORTED>
__global__ void xw_count_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, uint32_t M, uint32_t R, uint32_t W,
                                const uint32_t* __restrict__ tile_win, const uint32_t* __restrict__ tile_rb, uint64_t ngroups,
                                uint16_t* __restrict__ cp, uint32_t* __restrict__ grp_cnt, int* __restrict__ overflow, uint32_t nw) {
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
    uint32_t place = 0;
    for (uint32_t j = 0; j < 32; ++j) {
        const uint32_t cj = __shfl_sync(0xffffffffu, c, j);
        place += (cj > c) || (cj == c && j < lane);
    }
    cp[xw_cp_index(t, g, lane, R, nw)] = (uint16_t) (c | (place << 8));
    const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) grp_cnt[wg] = tot;
}
// one warp per (tile, group): jagged, sorted, slot-major fill (the layout xwin_kernel reads)
template <bool SORTED>
__global__ void xw_fill_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, const double* __restrict__ as, uint32_t M,
                               uint32_t R, uint32_t W, const uint32_t* __restrict__ tile_win, const uint32_t* __restrict__ tile_rb,
                               uint64_t ngroups, const uint16_t* __restrict__ cp, const uint32_t* __restrict__ grp_off,
                               uint16_t* __restrict__ col, double* __restrict__ val, uint32_t nw) {
    const uint64_t wg = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31, G = R / 32;
    if (wg >= ngroups) return;
    const uint32_t t = (uint32_t) (wg / G), g = (uint32_t) (wg % G);
    const uint32_t row = tile_rb[t] * R + g * 32 + lane;
    const uint64_t lo_c = (uint64_t) tile_win[t] * W, hi_c = lo_c + W;
    const uint32_t cpv = cp[xw_cp_index(t, g, lane, R, nw)], c = cpv & 0xffu, place = cpv >> 8;
    uint32_t j = 0, e = 0;
    if (row < M && c) {
        j = irp[row];
        e = irp[row + 1];
        if (SORTED) j = xw_lower_bound(ja, j, e, lo_c);
    }
    uint32_t off = grp_off[wg];
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, c);
    for (uint32_t k = 0; k < kmax; ++k) {
        const bool on = c > k;
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        if (on) {
            if (!SORTED) while (j < e && !(ja[j] >= lo_c && ja[j] < hi_c)) ++j;
            const uint32_t idx = off + place;  // rows with more than k entries are exactly the places [0, popc(m))
            val[idx] = as[j];
            col[idx] = (uint16_t) ((uint64_t) ja[j] - lo_c);
            ++j;
        }
        off += __popc(m);
    }
}

}  // namespace spmvb200
