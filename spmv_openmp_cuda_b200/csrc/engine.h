// engine.h -- private definition of the opaque handle of include/spmv_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace spmvb200 {
struct TileDesc;
struct LongRec;
// Plan of the pipelined host path (spmvb200_spmv_host): x goes up in pieces, row chunks start as soon as the
// pieces they read have landed, y chunks go down while later chunks compute (three streams, full-duplex PCIe).
struct HostPipe {
    int kind = -1, cand = -1, nch = 0, nch_req = 0, npieces = 0;
    std::vector<uint64_t> row_b;   // nch+1 row bounds
    std::vector<uint32_t> tile_b;  // nch+1 tile bounds (CSR stream kernel)
    std::vector<uint64_t> x_b;     // nch+1 bounds of the x pieces: chunk k reads only x[0, x_b[k+1])
    cudaStream_t s_up = nullptr, s_comp = nullptr, s_down = nullptr;
    cudaStream_t s_up2 = nullptr, s_down2 = nullptr;  // optional second copy stream per direction (SPMVB200_HOST_COPY_STREAMS=2)
    std::vector<cudaEvent_t> x_ready, k_start, k_end;  // k_start / k_end carry time stamps: recorded only when the caller asks for kernel_ms
    std::vector<cudaEvent_t> k_done;                   // chunk finished, no time stamp (tools/pipe_lab.cu: 32 timing events cost 0.65 ms per call)
};
constexpr uint64_t PAD = 16;  // readable slack after ja/as (vector loads and 16-byte aligned TMA tiles overrun)
}  // namespace spmvb200

// Device-resident matrix.  Indices are 32 bit, values fp64.
//   CSR          : irp[M+1], ja[NZ+PAD], as[NZ+PAD]  + row-block plan (desc/longrec/partial/ticket)
//   ELL colmajor : ja/as[K*pitch+PAD] (slot k of row r at k*pitch+r, pitch = M rounded up to 64), rl[M]
//   ELL rowmajor : ja/as[M*pitch+PAD] (slot k of row r at r*pitch+k, pitch = K rounded up to 4),  rl[M]
struct spmvb200_matrix {
    int format = 0;
    uint64_t M = 0, N = 0, NZ = 0, K = 0, pitch = 0, slots = 0;
    uint32_t* irp = nullptr;
    uint32_t* ja = nullptr;
    double* as = nullptr;
    uint32_t* rl = nullptr;
    uint16_t* ja16 = nullptr;  // column-major ELL: 16-bit column offsets from (row + ja16_base), when the matrix allows it
    int32_t ja16_base = 0;
    // SELL: irp = slice_ptr[nslices+1], rl = row lengths in sorted order, perm = sorted position -> row (0xffffffff = padding)
    uint32_t* perm = nullptr;
    uint64_t Mpad = 0;
    // x-window CSR (xwin.cuh): row blocks of xw_R rows x windows of xw_W columns; as = values in tile order
    uint32_t xw_R = 0, xw_W = 0, xw_nrb = 0, xw_ntiles = 0, xw_nbuf = 0, xw_nw = 0, xw_sorted = 1;
    uint32_t* xw_cta_rb = nullptr;    // [xw_ncta+1] row blocks of a persistent CTA (balanced by non-zeros)
    uint32_t xw_ncta = 0;
    int xw_mode = -1;                 // -1 not tuned yet, 0 one CTA per row block, 1 persistent CTAs (xw_cta_rb); first-use timing
    spmvb200_matrix* xw_child = nullptr;  // CSR handle: x-window or SELL copy built (and kept, if it won) by the adaptive mode's tuning run
    uint32_t lmax = 0;                    // CSR: longest row
    // SPMVB200_CSR_ROWS (the bit-exact kind): -1 not tuned yet, 0 stream kernel, 12 x-window copy, 13 SELL copy (all three sum in the serial order)
    int tuned_x = -1;
    spmvb200_matrix* x_child = nullptr;
    float tuned_x_ms[3] = {0, 0, 0};
    uint32_t* xw_rb_tile0 = nullptr;  // [nrb+1] first tile of a row block
    uint32_t* xw_tile_win = nullptr;  // [ntiles] window id
    uint32_t* xw_grp_off = nullptr;   // [ntiles*R/32+1] first entry of a (tile, 32-row group)
    uint16_t* xw_cnt = nullptr;       // [ntiles*R] entries of a row inside a tile | its place in the group's sorted order << 8
    uint16_t* xw_col = nullptr;       // [NZ+PAD] window-local column ids
    int own = 1;
    // hot-x hybrid (hotx.cuh; adaptive kind on matrices with power-law column popularity): the H hottest columns, the CSR column ids
    // remapped (hot column -> its rank < H, cold column c -> c + H) and a SELL copy of the short rows built from the remapped ids
    uint32_t* hot_cols = nullptr;
    uint32_t hot_H = 0;
    uint32_t* ja_hot = nullptr;
    spmvb200_matrix* hot_sell = nullptr;
    float hot_cover = 0.f;  // fraction of the non-zeros whose column is hot
    uint32_t* hot_slice_order = nullptr;  // slices of hot_sell by decreasing length
    int hot_shape = 0;                    // 0: 1024 threads x 1 CTA per SM, 1: 512 threads x 3 CTAs per SM (first-use timing)
    // stand-alone SELL handle of a skewed matrix: rows longer than VEC_MID live in `tail` (a compact CSR handle of just those rows,
    // run by the per-row / per-segment kernels next to the slices); tail_map[i] = original row of tail row i, tail_y = its scratch output
    spmvb200_matrix* tail = nullptr;
    uint32_t* tail_map = nullptr;
    double* tail_y = nullptr;
    // CSR stream plan
    spmvb200::TileDesc* desc = nullptr;
    spmvb200::LongRec* longrec = nullptr;
    double* partial = nullptr;
    uint32_t* ticket = nullptr;
    uint32_t* span_b = nullptr;     // nnz-balanced contiguous row spans, one per persistent CTA (vector-span kernel)
    uint32_t nspans = 0;
    uint32_t* mid_rows = nullptr;   // rows with VEC_MID < len <= TILE (vector kernels give them a CTA each)
    uint32_t nmid = 0;
    uint32_t* seg_tiles = nullptr;  // indices of the segment tiles (rows longer than one tile)
    uint32_t ntiles = 0, nlong = 0, nseg = 0;
    int l1pol = 0;  // L1 policy of the gather-bound kernels on this handle (common.cuh ld_mat / ld_xp)
    int vec_lanes = 32;
    int vec_tuned = 0;  // CSR: the sub-warp width of SPMVB200_CSR_ROWS_WARP has been timed on this matrix (first use)
    // SPMVB200_CSR_ADAPTIVE: candidate chosen by the first-use tuning run (-1 = not tuned yet)
    int tuned = -1;
    float tuned_ms[16] = {0};
    std::vector<uint32_t> h_tile_row0, h_tile_nnz0;  // host copy of the tile plan (chunking of the host path)
    spmvb200::HostPipe* pipe = nullptr;
    // host-path scratch
    double* d_x = nullptr;
    double* d_y = nullptr;
    void* flush = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // side streams of the vector path: medium-row and long-row CTAs run next to the main kernel (fork / join by events)
    cudaStream_t s_tail[2] = {nullptr, nullptr};
    cudaEvent_t e_fork = nullptr, e_tail[2] = {nullptr, nullptr};
};

namespace spmvb200 {
int finish_csr(spmvb200_matrix* m);  // build the plan once irp/ja/as are on the device
}
