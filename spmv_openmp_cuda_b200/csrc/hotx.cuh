// hotx.cuh -- "hot-x" hybrid for matrices whose COLUMN popularity is a power law (R-MAT, web / social graphs; BASELINE.json configs[2]).
//
// Why: on such matrices every kernel of this engine is bound by 32-byte sector traffic between L2 and the SMs -- one
// sector per gathered x value on top of 12 bytes of matrix stream -- not by HBM, and the gathers are anything but uniform: on R-MAT
// scale 22 the 16 384 most referenced columns (128 KB of x) take 41 % of all gathers.  L1 does not exploit that (20 % hit rate, and an
// evict_last / no_allocate policy pair changes nothing), so the hot part of x is cached EXPLICITLY:
//
//   * build (once per matrix, on the device): column histogram -> the H hottest columns -> column ids remapped so that a hot column's
//     id is its rank in [0, H) and a cold column c becomes c + H; a SELL-32-sigma copy of the rows of at most VEC_MID entries is built
//     from the remapped ids, rows longer than that keep their CSR order (remapped ids next to the parent's values).
//   * run: ONE persistent kernel, one CTA per SM.  Each CTA gathers the H hot x values into shared memory once, then its warps walk
//     the whole matrix as warp-sized work items, heaviest first: (a) rows of VEC_MID < len <= TILE, a warp per row, and the <= TILE-entry
//     segments of longer rows (partials combined in segment order by the last segment to finish, as in csr_longrow_kernel); (b) SELL
//     slices, a lane per row, left-to-right sums.  x[c] is a shared-memory read for c < H and a global gather of x[c - H] otherwise.
//   One launch instead of three concurrent ones, and 41 % fewer L2 sectors per SpMV.
// Rows of at most VEC_MID entries are summed in the serial order (bit-identical to sgemvSerial); longer rows by a shuffle tree: the
// candidate belongs to the tolerance kind SPMVB200_CSR_ADAPTIVE.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace spmvb200 {

// launch shapes: <threads per CTA, CTAs per SM>; the hot cache is H * 8 bytes of shared memory per CTA (H = 16384 with one CTA per SM,
// 8192 with three): more resident warps hide more gather latency, a bigger cache removes more gathers -- picked by timing

__global__ void hotx_hist_kernel(const uint32_t* __restrict__ ja, uint64_t nz, uint32_t* __restrict__ cnt) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; j < nz; j += stride) atomicAdd(cnt + ja[j], 1u);
}
__global__ void hotx_iota_kernel(uint32_t* __restrict__ v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
// remap[c] = c + H for every column, then remap[hot[i]] = i
__global__ void hotx_remap_init_kernel(uint32_t* __restrict__ remap, uint32_t n, uint32_t H) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) remap[i] = i + H;
}
__global__ void hotx_remap_hot_kernel(uint32_t* __restrict__ remap, const uint32_t* __restrict__ hot, uint32_t H) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H) remap[hot[i]] = i;
}
__global__ void hotx_apply_kernel(const uint32_t* __restrict__ ja, const uint32_t* __restrict__ remap, uint64_t nz, uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; j < nz; j += stride) out[j] = remap[ja[j]];
}

// sort keys for the slice order: 2^32-1 - slots of the slice (ascending sort = decreasing length)
__global__ void hotx_slice_keys_kernel(const uint32_t* __restrict__ slice_ptr, uint32_t nsl, uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nsl) {
        keys[i] = 0xffffffffu - (slice_ptr[i + 1] - slice_ptr[i]);
        ids[i] = i;
    }
}

// x value of remapped column id c: shared memory for the hot ones, L2 gather for the rest (both predicated: no branch)
__device__ __forceinline__ double hotx_get(const double* __restrict__ xs, const double* __restrict__ x, uint32_t c, uint32_t H) {
    double v;
    if (c < H) v = xs[c];
    else v = __ldg(x + (c - H));
    return v;
}

struct HotxArgs {
    // SELL part (rows of at most VEC_MID entries), remapped ids
    const uint32_t* slice_ptr;
    const uint32_t* perm;
    const uint32_t* rl;
    const double* sas;
    const uint32_t* sja;
    uint32_t nslices;
    // CSR part (rows longer than VEC_MID): parent's row pointer and values, remapped ids
    const uint32_t* irp;
    const uint32_t* ja;
    const double* as;
    const uint32_t* mid_rows;
    uint32_t nmid;
    const uint32_t* seg_tiles;
    uint32_t nseg;
    const TileDesc* desc;
    const LongRec* longrec;
    double* partial;
    uint32_t* ticket;
    // hot cache
    const uint32_t* hot_cols;
    uint32_t H;
    const uint32_t* slice_order;  // slices by decreasing length: the heavy ones first, so that the kernel's tail is made of tiny slices
};

template <int HOTX_BLOCK, int MINB>
__global__ void __launch_bounds__(HOTX_BLOCK, MINB)
hotx_kernel(const HotxArgs a, const double* __restrict__ x, double* __restrict__ y) {
    extern __shared__ __align__(16) double hx_s[];  // [H] hot x values
    const uint32_t H = a.H;
    for (uint32_t i = threadIdx.x; i < H; i += HOTX_BLOCK) hx_s[i] = __ldg(x + __ldg(a.hot_cols + i));
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * (HOTX_BLOCK / 32) + (threadIdx.x >> 5), nw = gridDim.x * (HOTX_BLOCK / 32);

    // (a) long rows and segments, a warp each: 4 non-zeros per lane in flight (two 16-byte value loads + two 8-byte id loads per step)
    const uint32_t nrow_items = a.nmid + a.nseg;
    for (uint32_t it = gw; it < nrow_items; it += nw) {
        uint32_t s, e, row = 0, tile = 0, rec = 0;
        const bool seg = it >= a.nmid;
        if (!seg) {
            row = __ldg(a.mid_rows + it);
            s = __ldg(a.irp + row);
            e = __ldg(a.irp + row + 1);
        } else {
            tile = __ldg(a.seg_tiles + (it - a.nmid));
            s = a.desc[tile].nnz0;
            e = a.desc[tile + 1].nnz0;
            rec = a.desc[tile].aux;
        }
        double t0 = 0, t1 = 0;
        for (uint32_t i = (s & ~1u) + 2 * lane; i < e; i += 128) {
            const uint32_t i2 = i + 64;
            const double2 v0 = ld_stream(reinterpret_cast<const double2*>(a.as + i));
            const uint2 c0 = ld_stream(reinterpret_cast<const uint2*>(a.ja + i));
            double2 v1 = make_double2(0.0, 0.0);
            uint2 c1 = make_uint2(0u, 0u);
            if (i2 < e) {
                v1 = ld_stream(reinterpret_cast<const double2*>(a.as + i2));
                c1 = ld_stream(reinterpret_cast<const uint2*>(a.ja + i2));
            }
            if (i >= s) t0 = fma(v0.x, hotx_get(hx_s, x, c0.x, H), t0);
            if (i + 1 < e) t0 = fma(v0.y, hotx_get(hx_s, x, c0.y, H), t0);
            if (i2 < e) t1 = fma(v1.x, hotx_get(hx_s, x, c1.x, H), t1);
            if (i2 + 1 < e) t1 = fma(v1.y, hotx_get(hx_s, x, c1.y, H), t1);
        }
        const double t = subwarp_sum<32>(t0 + t1);
        if (lane == 0) {
            if (!seg) {
                y[row] = t;
            } else {
                const LongRec lr = a.longrec[rec];
                a.partial[tile] = t;
                __threadfence();
                const uint32_t done = atomicAdd(a.ticket + rec, 1u);
                if (done == lr.ntiles - 1) {  // last segment of the row to finish: add the partials in segment order
                    __threadfence();
                    double acc = 0;
                    for (uint32_t k = 0; k < lr.ntiles; ++k) acc += __ldcg(a.partial + lr.first_tile + k);
                    y[lr.row] = acc;
                    a.ticket[rec] = 0;
                }
            }
        }
    }
    // (b) SELL slices, a lane per row, 4 slots in flight, left-to-right sums
    for (uint32_t so = gw; so < a.nslices; so += nw) {
        const uint32_t sl = __ldg(a.slice_order + so);
        const uint32_t i = sl * 32 + lane;
        const uint32_t len = __ldg(a.rl + i);
        const uint32_t sp0 = __ldg(a.slice_ptr + sl), sp1 = __ldg(a.slice_ptr + sl + 1);
        const uint32_t wmax = (sp1 - sp0) >> 5;
        const double* pa = a.sas + sp0 + lane;
        const uint32_t* pj = a.sja + sp0 + lane;
        double acc = 0;
        uint32_t k = 0;
        for (; k + 4 <= wmax; k += 4) {
            double v[4], xv[4];
            uint32_t c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = k + u < len;
                v[u] = ok ? ld_stream(pa + (k + u) * 32) : 0.0;
                c[u] = ok ? ld_stream(pj + (k + u) * 32) : H;  // H = cold column 0: never dereferenced (predicate below)
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = (k + u < len) ? hotx_get(hx_s, x, c[u], H) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k + u < len) acc = __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
        }
        for (; k < wmax; ++k)
            if (k < len) acc = __dadd_rn(acc, __dmul_rn(ld_stream(pa + k * 32), hotx_get(hx_s, x, ld_stream(pj + k * 32), H)));
        const uint32_t row = __ldg(a.perm + i);
        if (row != 0xffffffffu) y[row] = acc;
    }
}

}  // namespace spmvb200
