// plan.cuh -- the CSR row-block plan, built ON THE DEVICE (no host pass over the row pointer).
//
// Tiles are whole rows with <= STREAM_TILE non-zeros and <= STREAM_TILE_ROWS rows.  A row starts a new tile when
//   * its start offset falls into another quantile of width Q = TILE - SPECIAL than the previous row's, or
//   * it or its predecessor is "special" (longer than SPECIAL: such rows get tiles of their own), or
//   * its index is a multiple of TILE_ROWS.
// A tile's rows (except a special one, alone in its tile) all start inside one quantile, so the tile holds fewer
// than Q + SPECIAL = TILE non-zeros; SPECIAL = min(256, longest row), so tiles come out 87-100 % full in general and
// ~99 % full for matrices with short rows only.  Rows longer than TILE become ceil(len/TILE)
// segment tiles that share a LongRec.  Everything is flags -> exclusive scans -> scatter.
// Replaces the host loops a CPU-side planner would need for 3e7 rows; part of SURVEY.md §8f-1 (on-device
// construction of the "adaptive row-block map").
#pragma once
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace spmvb200 {

constexpr uint32_t PLAN_SPECIAL_MAX = 256;  // SPECIAL = min(this, longest row): regular matrices get Q = TILE - Lmax

__device__ __forceinline__ uint32_t plan_row_len(const uint32_t* irp, uint32_t r) { return irp[r + 1] - irp[r]; }

// per row: number of tiles that start there (0 = continues the previous tile), long-row flag, segment count
__global__ void plan_count_kernel(const uint32_t* __restrict__ irp, uint32_t M, uint32_t PLAN_SPECIAL, uint32_t* __restrict__ tiles_at,
                                  uint32_t* __restrict__ long_at, uint32_t* __restrict__ segs_at) {
    const uint32_t PLAN_Q = STREAM_TILE - PLAN_SPECIAL;
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > M) return;
    if (r == M) { tiles_at[r] = 0; long_at[r] = 0; segs_at[r] = 0; return; }
    const uint32_t len = plan_row_len(irp, r);
    bool start = r == 0 || len > PLAN_SPECIAL || (r % STREAM_TILE_ROWS) == 0;
    if (!start) start = plan_row_len(irp, r - 1) > PLAN_SPECIAL || irp[r] / PLAN_Q != irp[r - 1] / PLAN_Q;
    const bool is_long = len > (uint32_t) STREAM_TILE;
    const uint32_t nseg = is_long ? (len + STREAM_TILE - 1) / STREAM_TILE : 0u;
    tiles_at[r] = start ? (is_long ? nseg : 1u) : 0u;
    long_at[r] = is_long ? 1u : 0u;
    segs_at[r] = nseg;
}

__global__ void plan_scatter_kernel(const uint32_t* __restrict__ irp, uint32_t M, uint32_t NZ, const uint32_t* __restrict__ tiles_at,
                                    const uint32_t* __restrict__ tile_idx, const uint32_t* __restrict__ long_idx,
                                    const uint32_t* __restrict__ seg_idx, uint32_t ntiles, TileDesc* __restrict__ desc,
                                    LongRec* __restrict__ longrec, uint32_t* __restrict__ seg_tiles) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == M) desc[ntiles] = TileDesc{M, NZ, 0u, 0u};  // sentinel
    if (r >= M || tiles_at[r] == 0) return;
    const uint32_t t0 = tile_idx[r], n = tiles_at[r], s0 = irp[r];
    if (plan_row_len(irp, r) <= (uint32_t) STREAM_TILE) {
        desc[t0] = TileDesc{r, s0, 0u, 0u};
        return;
    }
    const uint32_t rec = long_idx[r], sb = seg_idx[r];
    longrec[rec] = LongRec{r, t0, n, 0u};
    for (uint32_t s = 0; s < n; ++s) {
        desc[t0 + s] = TileDesc{r | SEG_FLAG, s0 + s * STREAM_TILE, rec, 0u};
        seg_tiles[sb + s] = t0 + s;
    }
}

struct MidRowPred {
    const uint32_t* irp;
    __device__ bool operator()(uint32_t r) const {
        const uint32_t len = irp[r + 1] - irp[r];
        return len > (uint32_t) VEC_MID && len <= (uint32_t) STREAM_TILE;
    }
};

// span k starts at the first row whose start offset reaches k*NZ/ns (nnz-balanced contiguous spans)
__global__ void plan_spans_kernel(const uint32_t* __restrict__ irp, uint32_t M, uint64_t NZ, uint32_t ns, uint32_t* __restrict__ span_b) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > ns) return;
    if (k == 0) { span_b[0] = 0; return; }
    if (k == ns) { span_b[ns] = M; return; }
    const uint32_t target = (uint32_t) (NZ * k / ns);
    uint32_t lo = 0, hi = M;  // lower_bound over irp[0..M]
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (irp[mid] < target) lo = mid + 1; else hi = mid;
    }
    span_b[k] = lo;
}

__global__ void plan_split_desc_kernel(const TileDesc* __restrict__ desc, uint32_t n, uint32_t* __restrict__ row0, uint32_t* __restrict__ nnz0) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { row0[i] = desc[i].row0; nnz0[i] = desc[i].nnz0; }
}

}  // namespace spmvb200
