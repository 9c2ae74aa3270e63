/*
 * spmv_b200.h -- C ABI of the B200-native SpMV engine (libspmv_b200.so).
 *
 * Drop-in boundary for the GPU path of andreadiiorio/SpMV_openMP_CUDA (y = A*x, fp64, CSR and
 * ELLPACK).  Plain C: pointers and sizes only, no torch / C++ types.  Every entry point names the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions (the reference's): every function returns 0 (EXIT_SUCCESS) on success
 * and non-zero on failure; a diagnostic goes to stderr (like ERRPRINT, src/include/macros.h:57-58)
 * and is kept for spmvb200_last_error().  Index arrays coming from the host use the reference's
 * element type `ulong` == uint64_t (src/include/sparseMatrix.h:26-32); on the device the engine
 * narrows them to 32 bit (all supported matrices have M, N, NZ < 2^32 / 2^31, checked at upload).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails loudly.
 *
 * Threading: call from one host thread per matrix handle (the reference driver is single threaded
 * around the GPU path, src/main.cu:192-248).  Device-pointer entry points are asynchronous on the
 * stream given; host-pointer entry points synchronise before returning.
 */
#ifndef SPMV_B200_H
#define SPMV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPMVB200_VERSION 200

/* Kernel selectors.  Values 0..4 are the reference's CUDA compute modes in table order
 * (SpmvCUDA_CSRFuncs[0..1], SpmvCUDA_ELLFuncs[0..2], src/include/SpMV.h:130-142; mode strings
 * src/include/SpMV.h:37-41); later values are new modes appended after them. */
typedef enum {
    SPMVB200_CSR_ROWS          = 0, /* replaces cudaSpMVRowsCSR                 src/SpMV_CUDA.cu:33-49.  Bit-identical to sgemvSerial
                                       (src/SpMV_CSR_OMP.c:229-250) for every row of at most 2048 non-zeros; a longer row is split into
                                       2048-entry segments whose partial sums are combined in segment order: deterministic, within
                                       1e-12 * sum|a_ij x_j| of the serial sum, but not the same bits */
    SPMVB200_CSR_ROWS_WARP     = 1, /* replaces cudaSpMVWarpPerRowCSR           src/SpMV_CUDA.cu:52-73   */
    SPMVB200_ELL_ROWS          = 2, /* replaces cudaSpMVRowsELL (column-major)  src/SpMV_CUDA.cu:79-96   */
    SPMVB200_ELL_ROWS_NT       = 3, /* replaces cudaSpMVRowsELLNNTransposed     src/SpMV_CUDA.cu:99-115  */
    SPMVB200_ELL_ROWS_WARP_NT  = 4, /* replaces cudaSpMVWarpsPerRowELLNTrasposed src/SpMV_CUDA.cu:116-135 */
    SPMVB200_CSR_ADAPTIVE      = 5, /* new: row-length driven split (short rows streamed, long rows shared) */
    SPMVB200_SELL_ROWS         = 6, /* new: sliced ELL (SELL-32-sigma), thread per row, no global padding */
    SPMVB200_XWIN_ROWS         = 7, /* new: x-window CSR -- x gathered from shared-memory windows, thread per row */
    SPMVB200_KIND_COUNT        = 8
} spmvb200_kind;

/* Storage formats a handle can hold. */
typedef enum {
    SPMVB200_FMT_CSR          = 0,
    SPMVB200_FMT_ELL_COLMAJOR = 1, /* pitched, column-major: slot k of row r at k*pitch + r */
    SPMVB200_FMT_ELL_ROWMAJOR = 2, /* pitched, row-major:    slot k of row r at r*pitch + k */
    SPMVB200_FMT_SELL         = 3, /* sliced ELL: rows sorted by length inside windows of sigma rows, slices of 32 rows stored
                                      column-major and padded only to the slice's longest row */
    SPMVB200_FMT_XWIN         = 4  /* x-window CSR: row blocks x column windows, jagged slot-major groups of 32 rows, 16-bit
                                      window-local column ids; the kernel stages each x window in shared memory */
} spmvb200_format;

typedef struct spmvb200_matrix spmvb200_matrix; /* opaque, device resident */

/* ------------------------------------------------------------------ library / device */
const char* spmvb200_last_error(void);
int spmvb200_version(void);
/* number of SpMV kernels this process has launched through the library (all handles) */
unsigned long long spmvb200_launch_count(void);
int spmvb200_device_count(int* count);
int spmvb200_set_device(int device);
/* name, SM count, L2 bytes, total memory of the current device */
int spmvb200_device_info(char* name, size_t name_len, int* sm_count, size_t* l2_bytes, size_t* mem_bytes);

/* ------------------------------------------------------------------ upload  (host -> device)
 * Replaces spMatCpyCSR (src/commons/cudaUtils.cu:20-55).  Copies rows [row_begin,row_end) of a host
 * CSR matrix (reference layout: IRP[M+1], JA[NZ], AS[NZ], 64-bit indices) to the current device,
 * narrowing indices to 32 bit on the device, and builds the row-block plan used by the
 * SPMVB200_CSR_* kernels.  row_begin=0,row_end=M uploads everything; a sub-range is one GPU's
 * slice of a row-block partition (column ids stay global).  The host arrays are not modified and
 * may be freed afterwards. */
int spmvb200_csr_upload(uint64_t M, uint64_t N, const uint64_t* irp, const uint64_t* ja, const double* as,
                        uint64_t row_begin, uint64_t row_end, spmvb200_matrix** out);

/* Replaces ellTranspose + spMatCpyELL / spMatCpyELLNNPitched (src/commons/sparseUtils.c:145-185,
 * src/commons/cudaUtils.cu:56-140).  Input is the reference's ROW-MAJOR host ELL (M x K, zero padded:
 * AS=0, JA=0, src/lib/parser.c:217-296); `rl` is the row-length vector (mat->RL under -DROWLENS) or
 * NULL, in which case effective lengths are derived on the device (trailing AS==0 slots).
 * The column-major transposition for SPMVB200_FMT_ELL_COLMAJOR happens on the device. */
int spmvb200_ell_upload(uint64_t M, uint64_t N, uint64_t K, const uint64_t* ja, const double* as,
                        const uint64_t* rl, uint64_t row_begin, uint64_t row_end, int format,
                        spmvb200_matrix** out);

/* Adopt arrays that already live on the current device in the engine's narrow layout (used by the
 * on-device generators and by callers that build matrices on the GPU).  irp32 has M+1 entries,
 * ja32/as have NZ entries followed by at least 8 readable padding entries.  With own != 0 the
 * handle frees the arrays (cudaFree) when destroyed. */
int spmvb200_csr_adopt_device(uint64_t M, uint64_t N, uint64_t NZ, uint32_t* d_irp32, uint32_t* d_ja32,
                              double* d_as, int own, spmvb200_matrix** out);

/* Build an ELL handle (either layout) on the device from a CSR handle. */
int spmvb200_ell_from_csr(const spmvb200_matrix* csr, int format, spmvb200_matrix** out);

/* Build a SELL-32-sigma handle on the device from a CSR handle (sigma: sorting window in rows, a multiple of 32;
 * 0 = default 16384).  The successor of the ELL early-exit kernel for matrices with mixed short/long rows
 * (BASELINE.json configs[4]; the reference author's notes discuss SELL-C-sigma): coalesced like ELL, padding bounded
 * by the length spread inside a window instead of the global maximum. */
int spmvb200_sell_from_csr(const spmvb200_matrix* csr, uint32_t sigma, spmvb200_matrix** out);

/* Build an x-window CSR handle on the device from a CSR handle.  For matrices whose rows touch a column range too wide
 * for L1 but local enough that a block of rows_per_block rows needs only a few windows of window_cols columns (banded,
 * stencil, FEM-like: BASELINE.json configs[3]), the x gather is then served from shared memory instead of L2 -- on B200 the
 * L2 sector request rate (~1 per clock per SM) caps a gather-from-L2 SpMV at ~0.5 of the HBM roofline.
 * rows_per_block: power of two in [512, 4096] (0 = 2048); window_cols: even, <= 65536 (0 = 8192).  Fails if a row has more
 * than 255 non-zeros inside one window.  Column-sorted rows give results bit-identical to sgemvSerial. */
int spmvb200_xwin_from_csr(const spmvb200_matrix* csr, uint32_t rows_per_block, uint32_t window_cols, spmvb200_matrix** out);
/* geometry of an x-window handle; moved_bytes = everything one SpMV reads/writes including the x windows (from L2) */
int spmvb200_xwin_info(const spmvb200_matrix* m, uint32_t* rows_per_block, uint32_t* window_cols, uint32_t* ntiles,
                       uint32_t* ring, uint64_t* moved_bytes);

/* Replaces cudaFreeSpmat (src/include/cudaUtils.h:70-78). */
int spmvb200_free(spmvb200_matrix* m);

/* ------------------------------------------------------------------ queries */
int spmvb200_dims(const spmvb200_matrix* m, uint64_t* M, uint64_t* N, uint64_t* NZ, uint64_t* K, int* format);
/* algorithmic bytes of one SpMV with this handle:
 *   CSR: 12*NZ + 4*(M+1) + 8*N + 8*M      ELL: 12*NZ + 4*M + 8*N + 8*M      SELL: 12*NZ + 8*M + 8*N + 8*M  */
uint64_t spmvb200_algorithmic_bytes(const spmvb200_matrix* m);
/* bytes of device memory the handle's arrays occupy (includes ELL padding and the plan) */
uint64_t spmvb200_device_bytes(const spmvb200_matrix* m);
/* width of the column ids the handle's kernel reads: 32, or 16 when the engine found a narrower lossless encoding
 * (x-window CSR: window-local ids; column-major ELL: offsets from the row index when max(col-row) - min(col-row) < 2^16,
 * true of every stencil / banded matrix).  The algorithmic bytes above always count 32-bit ids. */
int spmvb200_index_bits(const spmvb200_matrix* m);
/* 1 if `kind` can run on this handle's format */
int spmvb200_kind_supported(const spmvb200_matrix* m, int kind);
const char* spmvb200_kind_name(int kind); /* the reference's mode string, e.g. "CUDA_CSR_ROWS" */
/* SPMVB200_CSR_ADAPTIVE times its candidate kernels (stream with 8 CTAs/SM, stream with a large L1,
 * vector with 2..32 lanes per row) on the handle's first adaptive launch and keeps the fastest;
 * this returns the winner's name ("" before that first launch). */
int spmvb200_adaptive_choice(const spmvb200_matrix* m, char* name, size_t len);
/* The bit-exact kinds pick at first use too: SPMVB200_CSR_ROWS between the stream kernel and an x-window or SELL copy,
 * SPMVB200_ELL_ROWS (when the ELL rectangle is >= 1.25 x the non-zeros) between the column-major kernel and a SELL copy built
 * from the ELL arrays; every candidate reproduces sgemvSerial (src/SpMV_CSR_OMP.c:229-250) bit for bit.  Returns
 * "stream" | "ell" | "xwindow" | "sell" ("" before the first launch of that kind). */
int spmvb200_exact_choice(const spmvb200_matrix* m, char* name, size_t len);

/* ------------------------------------------------------------------ compute
 * Device-resident SpMV: y[0..rows) = A[row_begin..row_end) * x, d_x has N doubles, d_y has
 * (row_end-row_begin) doubles, both on the handle's device.  Asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = default stream).  This is what replaces the launch
 *   f<<<Conf.gridSize,Conf.blockSize>>>(dMat,dVect,Conf,dOutV)   src/main.cu:233, test/SpMV_test.cu:112
 * -- the engine picks its own launch geometry. */
int spmvb200_spmv_device(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, void* stream);

/* First-use picks.  SPMVB200_CSR_ROWS / _ROWS_WARP / _ADAPTIVE, SPMVB200_ELL_ROWS and SPMVB200_XWIN_ROWS choose their kernel (and may
 * build a re-tiled copy of the matrix) the first time they run on a handle.  That first call BLOCKS: it allocates, synchronises the
 * device and -- in timed mode -- times candidates with CUDA events.  spmvb200_tune does it explicitly (d_y is scratch); afterwards
 * spmvb200_spmv_device is a pure asynchronous launch.  If the first launch of an unpicked handle happens on a stream that is being
 * CAPTURED, nothing is picked: the kind's plain kernel (stream / sub-warp / column-major ELL) is recorded into the graph instead. */
int spmvb200_tune(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, void* stream);
/* 0 (default) = timed: fastest candidate on this GPU, may differ from run to run, so the tolerance kinds (_ROWS_WARP, _ADAPTIVE) may
 * return different bits in different processes.  1 = deterministic: the pick is a pure function of the matrix structure; nothing
 * is timed.  Also settable with the environment variable SPMVB200_TUNE=deterministic.  Process-wide; the bit-exact kinds return
 * the same bits in either mode. */
int spmvb200_set_tuning_mode(int mode);
int spmvb200_get_tuning_mode(void);
/* Read / install a handle's picks (8 integers, see csrc/engine.cu): record them once, install them in every later process or on
 * every rank of a multi-GPU job, and all of them run the same kernels. */
int spmvb200_tuning_get(const spmvb200_matrix* m, int32_t picks[8]);
int spmvb200_tuning_set(spmvb200_matrix* m, const int32_t picks[8]);

/* Same launch, with FUSED OUTPUT DELIVERY for x <- y iterations over the GPUs of one box: besides
 * d_y, the kernel that computes row r stores it to dst[p][r + row_offset] for every destination p whose range
 * [lo[p], hi[p]) contains the global index r + row_offset.  Destinations are device-visible pointers -- typically the next
 * iteration's x on the other GPUs, mapped with spmvb200_ipc_open -- so the all-gather (for banded matrices: the halo
 * exchange) happens as posted NVLink stores while the SpMV is still running, not as a collective after it.  The ELL
 * column-major and x-window kernels deliver from their epilogue; other kinds add one pass over y.  A cross-GPU barrier
 * (spmvb200_peer_barrier) must separate an iteration's deliveries from the next iteration's reads. */
typedef struct {
    int      n;          /* destinations, 0..8 */
    double*  dst[8];
    uint64_t lo[8], hi[8];
    uint64_t row_offset; /* global index of the handle's row 0 */
} spmvb200_push;
int spmvb200_spmv_device_push(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, const spmvb200_push* push, void* stream);

/* Deliver rows that already sit in device memory (e.g. this GPU's freshly uploaded slice of x) the same way. */
int spmvb200_push_rows(const double* d_rows, uint64_t nrows, const spmvb200_push* push, void* stream);

/* Peer-memory plumbing for one-process-per-GPU jobs: export a cudaMalloc'ed buffer as a 64-byte CUDA IPC handle, map
 * another process's handle (peer access is enabled on first use), unmap. */
int spmvb200_ipc_export(void* d_ptr, unsigned char handle[64]);
int spmvb200_ipc_open(const unsigned char handle[64], void** d_ptr);
int spmvb200_ipc_close(void* d_ptr);
/* Barrier across the n GPUs of a box, as a one-block kernel on `stream`: d_flags[q] is GPU q's flag array (n zero-initialised
 * uint32_t; the local one for q == rank, peer-mapped otherwise); epoch must grow by one per use.  Every GPU's earlier work on
 * its stream (including stores to peer memory) is visible to the others' later work. */
int spmvb200_peer_barrier(uint32_t* const* d_flags, int n, int rank, uint32_t epoch, void* stream);

/* ------------------------------------------------------------------ row-block partition over the GPUs of one box
 * One process per GPU: rank g holds rows [splits[g], splits[g+1]) of a square matrix as handle `m` (global column
 * ids) and a replicated x.  This follows the reference's own CPU decomposition -- contiguous row blocks, spmvRowsBlocksCSR,
 * src/SpMV_CSR_OMP.c:65-99 with the block arithmetic of src/include/macros.h:33-36 -- with one block per GPU.  The shard owns
 * `nbuf` (2..4) x buffers of N doubles; peers map them through CUDA IPC.  Rendezvous: every rank calls _export, the caller
 * all-gathers the blobs (spmvb200_shard_blob_bytes each, rank order) with whatever it has (MPI, torch.distributed, a file), every rank
 * calls _connect.  col_range = {smallest, largest} column id the rank's rows reference (NULL: taken from a CSR handle; whole x for
 * other formats) -- it decides which rows each peer needs (a halo for banded matrices).
 *   _step      : x[dst][my rows] = A_local * x[src], and the rows the peers read are stored into THEIR x[dst] by the SpMV kernel's
 *                epilogue (posted NVLink stores), then a flag barrier across the GPUs; asynchronous on `stream`.
 *   _spmv_host : the whole host-buffer step in one call: x_slice (my rows of x, host) up over this GPU's PCIe link, halo rows first
 *                and delivered to the peers + barrier while the rest uploads, row chunks as their x pieces land, y chunks down while
 *                later chunks compute; returns when y_slice (my rows of y, host) is complete.  Every rank must make the call. */
typedef struct spmvb200_shard spmvb200_shard;
int spmvb200_shard_create(spmvb200_matrix* m, int kind, int rank, int world, const uint64_t* splits, int nbuf,
                          const uint64_t* col_range, spmvb200_shard** out);
size_t spmvb200_shard_blob_bytes(const spmvb200_shard* s);
int spmvb200_shard_export(spmvb200_shard* s, unsigned char* blob);
int spmvb200_shard_connect(spmvb200_shard* s, const unsigned char* blobs);
double* spmvb200_shard_x(spmvb200_shard* s, int buf);
int spmvb200_shard_halo_rows(const spmvb200_shard* s, uint64_t* rows); /* rows of mine delivered to peers per step (sum over peers) */
int spmvb200_shard_step(spmvb200_shard* s, int src, int dst, void* stream);
int spmvb200_shard_barrier(spmvb200_shard* s, void* stream); /* the flag barrier across the GPUs on its own; every rank calls it */
int spmvb200_shard_spmv_host(spmvb200_shard* s, const double* x_slice, double* y_slice, float* kernel_ms);
int spmvb200_shard_free(spmvb200_shard* s);

/* Iterated SpMV on one GPU (square matrices): x <- A x, `iters` times, ping-pong between d_a (holds x on entry) and d_b; the
 * result is in d_b if iters is odd, else in d_a.  use_graph != 0 captures the launch pair in a CUDA graph (no launch latency
 * between consecutive SpMVs).  *total_ms (may be NULL) = CUDA-event time of all iterations.  Synchronises before returning. */
int spmvb200_iterate_device(spmvb200_matrix* m, int kind, double* d_a, double* d_b, int iters, int use_graph, void* stream,
                            float* total_ms);

/* Host-buffer SpMV with the semantics of the reference's SPMV_INTERF
 *   int f(spmat* mat, double* x, CONFIG* cfg, double* y)        src/include/SpMV.h:63-64
 * x (N doubles) is copied to the device, the kernel runs, y (rows doubles) is copied back; the call
 * returns after y is complete.  *kernel_ms (may be NULL) receives the CUDA-event time of the kernel
 * alone -- the value a driver stores in ElapsedInternal (src/include/config.h:112).  Long vectors run as a pipeline
 * of row chunks (x pieces up, kernel chunks, y chunks down); asking for kernel_ms puts time-stamped events between the chunks,
 * which costs ~10 % of the call on 268 MB vectors (tools/pipe_lab.cu): pass NULL when the kernel time is not needed. */
int spmvb200_spmv_host(spmvb200_matrix* m, int kind, const double* x, double* y, float* kernel_ms);
/* Pageable x / y (what the reference driver passes: malloc, src/main.cu:155,181) move at about half the page-locked rate and cannot
 * overlap with the kernel.  spmvb200_host_register page-locks a caller buffer IN PLACE (cudaHostRegister; no-op for memory that is
 * page-locked already) until spmvb200_host_unregister(p) (p == NULL: all; also done by spmvb200_cache_drop(NULL)) -- the one call a
 * driver adds next to its malloc, and the other before its free().  A buffer freed while registered leaves a stale registration
 * behind (later CUDA calls on a re-used address fail with "invalid argument"), which is why the library does not register
 * buffers on its own unless SPMVB200_HOST_REGISTER=auto is set (then: the second time the same address and size come back). */
int spmvb200_host_register(const void* p, size_t bytes);
/* Page-locked vectors allocated by the library (cudaHostAlloc) -- the other one-line change: malloc -> spmvb200_host_alloc, free ->
 * spmvb200_host_free.  Fastest kind of host buffer: in-place registration of malloc'ed memory keeps its 4 KB pages, which cost ~18 % of
 * the duplex copy rate on the B200 box (tools/hostmem_lab.cu).  NULL on failure (spmvb200_last_error). */
void* spmvb200_host_alloc(size_t bytes);
int spmvb200_host_free(void* p);
int spmvb200_host_registered(const void* p); /* 1 if this library holds a registration of p */
int spmvb200_host_unregister(const void* p);

/* Repeat the kernel `reps` times on device-resident vectors and return per-repetition CUDA-event
 * times in ms (times_ms[reps]); with flush_l2 != 0 a buffer larger than L2 is read (1: leaves clean lines) or
 * overwritten (2: leaves dirty lines, whose write-back the timed kernel then pays for) between repetitions, outside
 * the timed region.  Mirrors the timing loop of testSpMVImplCuda
 * (test/SpMV_test.cu:103-145) with events instead of a host stopwatch. */
int spmvb200_time_device(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, int reps, int flush_l2,
                         float* times_ms);

/* ------------------------------------------------------------------ host adapter cache
 * One-call form for a driver that only has host arrays (what an SPMV_INTERF adapter needs): the
 * device copy is created on first use and cached under `key` (e.g. the host spmat*), so the
 * upload is not repeated on later calls.  is_ell selects ELL (ja/as row-major M x K) or CSR input. */
int spmvb200_cached_spmv(const void* key, int kind, int is_ell, uint64_t M, uint64_t N, uint64_t K,
                         const uint64_t* irp, const uint64_t* ja, const double* as, const uint64_t* rl,
                         const double* x, double* y, double* elapsed_internal_s);
/* The cached copy is re-uploaded when the pointers, the dimensions or a content fingerprint change (row pointer / row lengths plus
 * 4096 evenly spaced (JA, AS) samples).  An in-place edit confined to values the fingerprint does not sample is NOT seen: call
 * spmvb200_cache_drop(key) after editing a matrix in place. */
int spmvb200_cache_drop(const void* key); /* key == NULL drops everything */

/* ------------------------------------------------------------------ comparators (host arrays; no compute path)
 * The strict check of SURVEY.md §8c: |y_i - yref_i| <= tau * sum_j |a_ij x_j| for every row, NaN / Inf anywhere in y fails
 * (the reference's doubleVectorsDiff, src/commons/utils.c:362-393, is an absolute 7e-4 and lets never-written outputs pass).
 * n_bad = rows that fail, worst_ratio = max |dy| / sum|a x|. */
int spmvb200_compare_strict_csr(uint64_t M, const uint64_t* irp, const uint64_t* ja, const double* as, const double* x,
                                const double* y_ref, const double* y, double tau, uint64_t* n_bad, double* worst_ratio);
/* doubleVectorsDiff semantics (largest |a-b|, failed when it exceeds the threshold; reference: 7e-4), but NaN fails */
int spmvb200_compare_abs(uint64_t n, const double* a, const double* b, double threshold, int* failed, double* max_diff);

/* ------------------------------------------------------------------ device vectors (plumbing) */
int spmvb200_dmalloc(void** d_ptr, size_t bytes);
int spmvb200_dfree(void* d_ptr);
int spmvb200_h2d(void* d_dst, const void* h_src, size_t bytes);
int spmvb200_d2h(void* h_dst, const void* d_src, size_t bytes);
int spmvb200_sync(void);
int spmvb200_h2d_async(void* d_dst, const void* h_src, size_t bytes, void* stream); /* h_src pinned for a true async copy */
int spmvb200_d2h_async(void* h_dst, const void* d_src, size_t bytes, void* stream);
int spmvb200_stream_sync(void* stream);

/* ------------------------------------------------------------------ synthetic workloads
 * Seeded, counter-based generators of the BASELINE.json matrices; every row (R-MAT: every edge) is a
 * pure function of (seed, index) so the host and the device produce identical matrices and any row
 * range can be generated on its own.  Not part of the reference (it reads Matrix Market files);
 * they exist because 1e9 non-zeros cannot go through a text file.   kind:
 *   1 = 5-point 2-D Laplacian on an n x n grid            (p0 = n)
 *   2 = 27-point 3-D stencil on an nx x ny x nz grid      (p0 = nx, p1 = ny, p2 = nz)
 *   4 = random banded: p1 non-zeros per row in [i-p0, i+p0], stratified     (p0 = half width, p1 = nnz/row, p2 = M)
 *   5 = mixed rows: length p1 with probability p3/2^32, else length 4, uniform columns (p0 = M, p1 = K_max, p3 = prob)
 * (kind 3, R-MAT, is edge based: see spmvb200_synth_rmat_*). */
typedef struct {
    int      kind;
    uint64_t seed;
    uint64_t p0, p1, p2, p3;
} spmvb200_synth;

int spmvb200_synth_dims(const spmvb200_synth* s, uint64_t* M, uint64_t* N);
/* host, reference layout (64-bit): row lengths, then fill given the local row pointer
 * (irp_local[0] = 0 for row_begin) */
int spmvb200_synth_rowlen_host(const spmvb200_synth* s, uint64_t row_begin, uint64_t row_end, uint64_t* rl);
int spmvb200_synth_fill_host(const spmvb200_synth* s, uint64_t row_begin, uint64_t row_end,
                             const uint64_t* irp_local, uint64_t* ja, double* as);
/* device: build rows [row_begin,row_end) directly as a CSR handle on the current device */
int spmvb200_synth_csr_device(const spmvb200_synth* s, uint64_t row_begin, uint64_t row_end, spmvb200_matrix** out);
/* R-MAT (a,b,c,d = .57,.19,.19,.05): 64-bit keys row<<32|col of edges [e_begin,e_end) on the host;
 * on the device the whole matrix (keys sorted, duplicates merged, values hashed from (row,col)). */
int spmvb200_synth_rmat_keys_host(int scale, uint64_t seed, uint64_t e_begin, uint64_t e_end, uint64_t* keys);
int spmvb200_synth_rmat_values_host(uint64_t seed, uint64_t n, const uint64_t* keys, double* as);
int spmvb200_synth_rmat_csr_device(int scale, uint64_t n_edges, uint64_t seed, spmvb200_matrix** out);
/* x_i = U(-1,1) * scale, hashed from (seed, i) */
int spmvb200_synth_vector_host(uint64_t seed, uint64_t begin, uint64_t end, double scale, double* x);
int spmvb200_synth_vector_device(uint64_t seed, uint64_t begin, uint64_t end, double scale, double* d_x);

/* smallest / largest column id a CSR handle references: the part of x its rows read (halo planning of the multi-GPU iteration;
 * both 0 for a handle without non-zeros) */
int spmvb200_col_range(const spmvb200_matrix* m, uint64_t* col_min, uint64_t* col_max);

/* copy a handle's narrow CSR arrays back to the host (tests / CPU baseline on device-built matrices) */
int spmvb200_csr_download(const spmvb200_matrix* m, uint64_t* irp, uint64_t* ja, double* as);

#ifdef __cplusplus
}
#endif

/* ------------------------------------------------------------------ reference-typed adapters
 * Available when this header is included AFTER the reference's src/include/sparseMatrix.h and
 * SpMV.h: functions of type SPMV (src/include/SpMV.h:63) that can be appended to SpmvCSRFuncs /
 * SpmvELLFuncs (src/include/SpMV.h:144-159) and run by the unchanged testSpMVImplOMP
 * (test/SpMV_test.cu:67-101).  They are compiled in the CALLER's translation unit, so they always
 * see the caller's own `spmat` layout (-DROWLENS, __CUDACC__: SURVEY.md §2.3-10). */
#if defined(SPARSEMATRIX) && defined(_SPMV)
#ifdef ROWLENS
#define SPMVB200_RL_(m) ((const uint64_t*) (m)->RL)
#else
#define SPMVB200_RL_(m) ((const uint64_t*) 0)
#endif
#define SPMVB200_DEFINE_CSR_ADAPTER(NAME, KIND)                                                              \
    static inline int NAME(spmat* mat, double* x, CONFIG* cfg, double* y) {                                  \
        (void) cfg;                                                                                          \
        return spmvb200_cached_spmv(mat, KIND, 0, mat->M, mat->N, 0, (const uint64_t*) mat->IRP,             \
                                    (const uint64_t*) mat->JA, mat->AS, SPMVB200_RL_(mat), x, y,             \
                                    &ElapsedInternal);                                                       \
    }
#define SPMVB200_DEFINE_ELL_ADAPTER(NAME, KIND)                                                              \
    static inline int NAME(spmat* mat, double* x, CONFIG* cfg, double* y) {                                  \
        (void) cfg;                                                                                          \
        return spmvb200_cached_spmv(mat, KIND, 1, mat->M, mat->N, mat->MAX_ROW_NZ, (const uint64_t*) 0,      \
                                    (const uint64_t*) mat->JA, mat->AS, SPMVB200_RL_(mat), x, y,             \
                                    &ElapsedInternal);                                                       \
    }
SPMVB200_DEFINE_CSR_ADAPTER(b200SpMVRowsCSR, SPMVB200_CSR_ROWS)
SPMVB200_DEFINE_CSR_ADAPTER(b200SpMVWarpPerRowCSR, SPMVB200_CSR_ROWS_WARP)
SPMVB200_DEFINE_CSR_ADAPTER(b200SpMVAdaptiveCSR, SPMVB200_CSR_ADAPTIVE)
SPMVB200_DEFINE_CSR_ADAPTER(b200SpMVRowsSELL, SPMVB200_SELL_ROWS) /* CSR in, SELL-32-sigma built on the device */
SPMVB200_DEFINE_CSR_ADAPTER(b200SpMVRowsXWIN, SPMVB200_XWIN_ROWS) /* CSR in, x-window CSR built on the device */
SPMVB200_DEFINE_ELL_ADAPTER(b200SpMVRowsELL, SPMVB200_ELL_ROWS)
SPMVB200_DEFINE_ELL_ADAPTER(b200SpMVRowsELLNNTransposed, SPMVB200_ELL_ROWS_NT)
SPMVB200_DEFINE_ELL_ADAPTER(b200SpMVWarpsPerRowELLNTrasposed, SPMVB200_ELL_ROWS_WARP_NT)
#endif /* reference headers present */

#endif /* SPMV_B200_H */
