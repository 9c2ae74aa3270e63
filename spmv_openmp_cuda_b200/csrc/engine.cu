// engine.cu -- the thin C-ABI layer of include/spmv_b200.h: device allocation, H2D upload with
// on-device 64->32-bit narrowing / ELL transposition, plan building, kernel launch, D2H.
// Replaces src/commons/cudaUtils.cu (spMatCpyCSR/ELL*, cudaFreeSpmat) and the launch+sync+download
// code of src/main.cu:192-248 / test/SpMV_test.cu:103-145.  No CPU compute path exists here.
// One translation unit in five files: this one (globals, plumbing, queries, timing loop, adapter cache) textually includes
// engine_formats.inl (plan + uploads + SELL / x-window construction), engine_launch.inl (launchers, first-use tuning, kind switch),
// engine_multigpu.inl (fused delivery, IPC, barrier, CUDA-graph iteration) and engine_hostpath.inl (spmvb200_spmv_host).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/spmv_b200.h"
#include "engine.h"
#include "kernels.cuh"
#include "plan.cuh"
#include "xwin.cuh"
#include "hotx.cuh"

namespace spmvb200 {
thread_local char g_err[512] = "";
thread_local int g_quiet = 0;
static std::atomic<unsigned long long> g_launches{0};
}  // namespace spmvb200
using namespace spmvb200;

// -------------------------------------------------------------------------------------------------
static void destroy_pipe(HostPipe* p) {
    if (!p) return;
    for (auto e : p->x_ready) cudaEventDestroy(e);
    for (auto e : p->k_start) cudaEventDestroy(e);
    for (auto e : p->k_end) cudaEventDestroy(e);
    for (auto e : p->k_done) cudaEventDestroy(e);
    if (p->s_up) cudaStreamDestroy(p->s_up);
    if (p->s_comp) cudaStreamDestroy(p->s_comp);
    if (p->s_down) cudaStreamDestroy(p->s_down);
    if (p->s_up2) cudaStreamDestroy(p->s_up2);
    if (p->s_down2) cudaStreamDestroy(p->s_down2);
    delete p;
}

static void free_arrays(spmvb200_matrix* m) {
    if (m->xw_child) {
        spmvb200_free(m->xw_child);
        m->xw_child = nullptr;
    }
    if (m->x_child) {
        spmvb200_free(m->x_child);
        m->x_child = nullptr;
    }
    if (m->tail) {
        spmvb200_free(m->tail);
        m->tail = nullptr;
    }
    if (m->hot_sell) {
        spmvb200_free(m->hot_sell);
        m->hot_sell = nullptr;
    }
    cudaFree(m->hot_cols);
    cudaFree(m->ja_hot);
    cudaFree(m->hot_slice_order);
    m->hot_cols = nullptr;
    m->ja_hot = nullptr;
    m->hot_slice_order = nullptr;
    cudaFree(m->tail_map);
    cudaFree(m->tail_y);
    if (m->own) {
        cudaFree(m->irp);
        cudaFree(m->ja);
        cudaFree(m->as);
        cudaFree(m->rl);
        cudaFree(m->ja16);
        cudaFree(m->perm);
    }
    cudaFree(m->xw_rb_tile0);
    cudaFree(m->xw_cta_rb);
    cudaFree(m->xw_tile_win);
    cudaFree(m->xw_grp_off);
    cudaFree(m->xw_cnt);
    cudaFree(m->xw_col);
    cudaFree(m->desc);
    cudaFree(m->longrec);
    cudaFree(m->partial);
    cudaFree(m->ticket);
    cudaFree(m->seg_tiles);
    cudaFree(m->span_b);
    cudaFree(m->mid_rows);
    cudaFree(m->d_x);
    cudaFree(m->d_y);
    cudaFree(m->flush);
    for (int i = 0; i < 2; ++i) {
        if (m->s_tail[i]) cudaStreamDestroy(m->s_tail[i]);
        if (m->e_tail[i]) cudaEventDestroy(m->e_tail[i]);
    }
    if (m->e_fork) cudaEventDestroy(m->e_fork);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    destroy_pipe(m->pipe);
    m->pipe = nullptr;
}

extern "C" const char* spmvb200_last_error(void) { return g_err; }
extern "C" int spmvb200_version(void) { return SPMVB200_VERSION; }
extern "C" unsigned long long spmvb200_launch_count(void) { return g_launches.load(); }

extern "C" int spmvb200_device_count(int* count) {
    if (!count) return fail("device_count: null argument");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail("cudaGetDeviceCount -> %s (no CUDA device: this engine has no CPU fallback)", cudaGetErrorString(e));
    }
    return 0;
}
extern "C" int spmvb200_set_device(int device) {
    CU_TRY(cudaSetDevice(device));
    return 0;
}
extern "C" int spmvb200_device_info(char* name, size_t name_len, int* sm_count, size_t* l2_bytes, size_t* mem_bytes) {
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    CU_TRY(cudaGetDeviceProperties(&p, dev));
    if (name && name_len) snprintf(name, name_len, "%s", p.name);
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (l2_bytes) *l2_bytes = (size_t) p.l2CacheSize;
    if (mem_bytes) *mem_bytes = p.totalGlobalMem;
    return 0;
}

// ------------------------------------------------------------------------------------------------- plumbing
extern "C" int spmvb200_dmalloc(void** d_ptr, size_t bytes) {
    if (!d_ptr) return fail("dmalloc: null argument");
    CU_TRY(cudaMalloc(d_ptr, bytes ? bytes : 16));
    return 0;
}
extern "C" int spmvb200_dfree(void* d_ptr) {
    CU_TRY(cudaFree(d_ptr));
    return 0;
}
extern "C" int spmvb200_h2d(void* d, const void* h, size_t bytes) {
    CU_TRY(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int spmvb200_d2h(void* h, const void* d, size_t bytes) {
    CU_TRY(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int spmvb200_sync(void) {
    CU_TRY(cudaDeviceSynchronize());
    return 0;
}

#include "engine_formats.inl"

extern "C" int spmvb200_free(spmvb200_matrix* m) {
    if (!m) return 0;
    free_arrays(m);
    delete m;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("free: %s", cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------------- queries
extern "C" int spmvb200_dims(const spmvb200_matrix* m, uint64_t* M, uint64_t* N, uint64_t* NZ, uint64_t* K, int* format) {
    if (!m) return fail("dims: null handle");
    if (M) *M = m->M;
    if (N) *N = m->N;
    if (NZ) *NZ = m->NZ;
    if (K) *K = m->K;
    if (format) *format = m->format;
    return 0;
}
extern "C" uint64_t spmvb200_algorithmic_bytes(const spmvb200_matrix* m) {
    if (!m) return 0;
    if (m->format == SPMVB200_FMT_CSR || m->format == SPMVB200_FMT_XWIN) return 12 * m->NZ + 4 * (m->M + 1) + 8 * m->N + 8 * m->M;
    if (m->format == SPMVB200_FMT_SELL) return 12 * m->NZ + 8 * m->M + 8 * m->N + 8 * m->M;  // row length + permutation per row
    return 12 * m->NZ + 4 * m->M + 8 * m->N + 8 * m->M;
}
extern "C" uint64_t spmvb200_device_bytes(const spmvb200_matrix* m) {
    if (!m) return 0;
    if (m->format == SPMVB200_FMT_CSR)
        return (m->NZ + PAD) * 12 + (m->M + 1) * 4 + ((uint64_t) m->ntiles + 1) * sizeof(TileDesc) + (uint64_t) m->ntiles * 8 +
               (uint64_t) m->nlong * (sizeof(LongRec) + 4);
    if (m->format == SPMVB200_FMT_SELL) return (m->slots + PAD) * 12 + m->Mpad * 8 + (m->Mpad / 32 + 1) * 4;
    if (m->format == SPMVB200_FMT_XWIN)
        return (m->NZ + PAD) * 10 + (uint64_t) m->xw_ntiles * m->xw_R * 2 + ((uint64_t) m->xw_ntiles * (m->xw_R / 32) + 1) * 4 +
               (uint64_t) m->xw_ntiles * 4 + ((uint64_t) m->xw_nrb + 1) * 4;
    return (m->slots + PAD) * (m->ja16 ? 14 : 12) + m->M * 4;
}
extern "C" int spmvb200_index_bits(const spmvb200_matrix* m) {
    if (!m) return 0;
    if (m->format == SPMVB200_FMT_XWIN) return 16;
    if (m->format == SPMVB200_FMT_ELL_COLMAJOR && m->x_child && m->tuned_x > 0) return 32;  // ELL_ROWS runs its SELL copy (32-bit ids)
    if (m->format == SPMVB200_FMT_ELL_COLMAJOR && m->ja16 && !getenv("SPMVB200_ELL_NO_EARLY_EXIT")) return 16;
    return 32;
}
extern "C" int spmvb200_kind_supported(const spmvb200_matrix* m, int kind) {
    if (!m) return 0;
    switch (kind) {
        case SPMVB200_CSR_ROWS:
        case SPMVB200_CSR_ROWS_WARP:
        case SPMVB200_CSR_ADAPTIVE: return m->format == SPMVB200_FMT_CSR;
        case SPMVB200_ELL_ROWS: return m->format == SPMVB200_FMT_ELL_COLMAJOR;
        case SPMVB200_SELL_ROWS: return m->format == SPMVB200_FMT_SELL;
        case SPMVB200_XWIN_ROWS: return m->format == SPMVB200_FMT_XWIN;
        case SPMVB200_ELL_ROWS_NT:
        case SPMVB200_ELL_ROWS_WARP_NT: return m->format == SPMVB200_FMT_ELL_ROWMAJOR;
        default: return 0;
    }
}
extern "C" const char* spmvb200_kind_name(int kind) {
    switch (kind) {  // mode strings of src/include/SpMV.h:37-41 (+ the appended mode)
        case SPMVB200_CSR_ROWS: return "CUDA_CSR_ROWS";
        case SPMVB200_CSR_ROWS_WARP: return "CUDA_CSR_ROWS_WARP";
        case SPMVB200_ELL_ROWS: return "CUDA_ELL_ROWS";
        case SPMVB200_ELL_ROWS_NT: return "CUDA_ELL_ROWS_WARP_NN_TRANSPOSED_1T";
        case SPMVB200_ELL_ROWS_WARP_NT: return "CUDA_ELL_ROWS_WARP_NN_TRANSPOSED";
        case SPMVB200_CSR_ADAPTIVE: return "CUDA_CSR_ADAPTIVE";
        case SPMVB200_SELL_ROWS: return "CUDA_SELL_ROWS";
        case SPMVB200_XWIN_ROWS: return "CUDA_CSR_XWINDOW_ROWS";
        default: return "?";
    }
}

#include "engine_launch.inl"

extern "C" int spmvb200_spmv_device(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, void* stream) {
    if (!m || !d_x || !d_y) return fail("spmv_device: null argument");
    if (prefer_smem_once()) return 1;
    return launch(m, kind, d_x, d_y, (cudaStream_t) stream);
}

// Make the first-use pick of a self-tuning kind NOW (blocking: allocates, synchronises, and -- in timed mode -- times candidates on
// d_x / d_y, which are scratch here: d_y is overwritten).  After it, spmvb200_spmv_device is a pure asynchronous launch and may be
// captured into a CUDA graph with the tuned kernel inside.
extern "C" int spmvb200_tune(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, void* stream) {
    if (!m || !d_x || !d_y) return fail("tune: null argument");
    if (!spmvb200_kind_supported(m, kind)) return fail("kind %d (%s) cannot run on format %d", kind, spmvb200_kind_name(kind), m->format);
    if (prefer_smem_once()) return 1;
    if (!needs_tuning(m, kind)) return 0;
    if (stream_capturing((cudaStream_t) stream)) return fail("tune: the stream is capturing; tune before cudaStreamBeginCapture");
    if (launch(m, kind, d_x, d_y, (cudaStream_t) stream)) return 1;
    CU_TRY(cudaStreamSynchronize((cudaStream_t) stream));
    return 0;
}
extern "C" int spmvb200_set_tuning_mode(int mode) {
    if (mode != 0 && mode != 1) return fail("set_tuning_mode: 0 = timed, 1 = deterministic");
    g_tune_mode = mode;
    return 0;
}
extern "C" int spmvb200_get_tuning_mode(void) { return tune_mode(); }
// t[0] adaptive candidate (-1: not picked yet)   t[1] exact-kind / ELL pick (-1; 0 plain kernel; 12 x-window copy; 13 SELL copy)
// t[2] lanes of the warp kind (0: not picked)     t[3] x-window launch shape of this handle or of the adaptive kind's copy (-1)
// t[4] x-window launch shape of the exact kind's copy (-1)                                 t[5..7] reserved (0)
extern "C" int spmvb200_tuning_get(const spmvb200_matrix* m, int32_t t[8]) {
    if (!m || !t) return fail("tuning_get: null argument");
    for (int i = 0; i < 8; ++i) t[i] = 0;
    t[0] = m->tuned;
    t[1] = m->tuned_x;
    t[2] = m->vec_tuned ? m->vec_lanes : 0;
    t[3] = m->format == SPMVB200_FMT_XWIN ? m->xw_mode : (m->tuned == CAND_XWIN && m->xw_child ? m->xw_child->xw_mode : -1);
    t[4] = (m->tuned_x == CAND_XWIN && m->x_child) ? m->x_child->xw_mode : -1;
    return 0;
}
// Install picks recorded earlier (same matrix, e.g. another process or another GPU of the job) instead of timing: every rank / run
// then executes the same kernels and the tolerance kinds return the same bits.  Builds the re-tiled copies the picks name.
extern "C" int spmvb200_tuning_set(spmvb200_matrix* m, const int32_t t[8]) {
    if (!m || !t) return fail("tuning_set: null argument");
    if (m->format == SPMVB200_FMT_XWIN) {
        if (t[3] >= 0) m->xw_mode = t[3] ? 1 : 0;
        return 0;
    }
    const bool csr = m->format == SPMVB200_FMT_CSR, ell = m->format == SPMVB200_FMT_ELL_COLMAJOR;
    if (csr && t[0] >= 0) {
        if (t[0] >= N_CAND) return fail("tuning_set: adaptive candidate %d out of range", t[0]);
        if (m->xw_child) { spmvb200_free(m->xw_child); m->xw_child = nullptr; }
        m->tuned = -1;
        hotx_drop(m);
        if (t[0] == CAND_HOTX) {
            if (hotx_build_quiet(m)) return fail("tuning_set: the matrix does not fit the hot-x hybrid");
        } else if (t[0] == CAND_XWIN || t[0] == CAND_SELL) {
            if (build_child(m, t[0], false, &m->xw_child)) return fail("tuning_set: the matrix does not fit the %s copy", CAND_NAME[t[0]]);
            if (t[0] == CAND_XWIN) m->xw_child->xw_mode = t[3] >= 0 ? (t[3] ? 1 : 0) : xwin_mode_rule(m->xw_child);
        }
        m->tuned = t[0];
    }
    if ((csr || ell) && t[1] >= 0) {
        if (t[1] != 0 && t[1] != CAND_SELL && !(csr && t[1] == CAND_XWIN)) return fail("tuning_set: exact-kind pick %d not valid for this format", t[1]);
        if (m->x_child) { spmvb200_free(m->x_child); m->x_child = nullptr; }
        m->tuned_x = -1;
        if (csr && t[1] != 0) {
            if (build_child(m, t[1], true, &m->x_child)) return fail("tuning_set: the matrix does not fit the %s copy", CAND_NAME[t[1]]);
            if (t[1] == CAND_XWIN) m->x_child->xw_mode = t[4] >= 0 ? (t[4] ? 1 : 0) : xwin_mode_rule(m->x_child);
        } else if (ell && t[1] == CAND_SELL) {
            const int q = g_quiet;
            g_quiet = 1;
            const int rc = sell_build(m, 0, 0xffffffffu, &m->x_child);
            g_quiet = q;
            if (rc) return fail("tuning_set: SELL copy of the ELL handle could not be built");
        }
        m->tuned_x = t[1];
    }
    if (csr && t[2] > 0) {
        if (t[2] != 2 && t[2] != 4 && t[2] != 8 && t[2] != 16 && t[2] != 32) return fail("tuning_set: %d lanes", t[2]);
        m->vec_lanes = t[2];
        m->vec_tuned = 1;
    }
    if (m->pipe) { destroy_pipe(m->pipe); m->pipe = nullptr; }
    return 0;
}

#include "engine_multigpu.inl"

#include "engine_hostpath.inl"

#include "engine_shard.inl"

// L2 flush by READING a buffer larger than L2: leaves the cache full of CLEAN lines.  (A memset leaves it full of dirty
// lines, whose write-back the next kernel's fills then pay for -- up to one extra byte written per byte read.)
__global__ void flush_read_kernel(const uint4* p, size_t n, uint4* sink) {
    uint4 a = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 v = p[i];
        a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
    }
    if ((a.x ^ a.y ^ a.z ^ a.w) == 0x9e3779b9u) *sink = a;  // never true for a zeroed buffer: keeps the loads alive
}

extern "C" int spmvb200_time_device(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, int reps, int flush_l2,
                                    float* times_ms) {
    if (!m || !d_x || !d_y || reps <= 0 || !times_ms) return fail("time_device: bad arguments");
    if (prefer_smem_once() || ensure_events(m)) return 1;
    const size_t FLUSH = 512ull << 20;  // > 126 MB L2
    if (flush_l2 && !m->flush) {
        CU_TRY(cudaMalloc(&m->flush, FLUSH));
        CU_TRY(cudaMemset(m->flush, 0, FLUSH));
    }
    for (int i = 0; i < reps; ++i) {
        if (flush_l2 == 2) CU_TRY(cudaMemsetAsync(m->flush, 0, FLUSH, 0));  // dirty flush (what round 1 first measured with)
        else if (flush_l2) flush_read_kernel<<<1184, 512>>>((const uint4*) m->flush, FLUSH / 16, (uint4*) m->flush);
        CU_TRY(cudaEventRecord(m->ev0, 0));
        if (launch(m, kind, d_x, d_y, 0)) return 1;
        CU_TRY(cudaEventRecord(m->ev1, 0));
        CU_TRY(cudaEventSynchronize(m->ev1));
        CU_TRY(cudaEventElapsedTime(times_ms + i, m->ev0, m->ev1));
    }
    return 0;
}

extern "C" int spmvb200_adaptive_choice(const spmvb200_matrix* m, char* name, size_t len) {
    if (!m || !name || !len) return fail("adaptive_choice: bad arguments");
    snprintf(name, len, "%s", m->tuned >= 0 ? CAND_NAME[m->tuned] : "");
    return 0;
}

extern "C" int spmvb200_exact_choice(const spmvb200_matrix* m, char* name, size_t len) {
    if (!m || !name || !len) return fail("exact_choice: bad arguments");
    const bool ell = m->format == SPMVB200_FMT_ELL_COLMAJOR;
    snprintf(name, len, "%s", m->tuned_x < 0 ? "" : m->tuned_x == CAND_XWIN ? "xwindow" : m->tuned_x == CAND_SELL ? "sell" : ell ? "ell" : "stream");
    return 0;
}

// smallest and largest column id a CSR handle references (which part of x its rows read: halo planning of the multi-GPU iteration)
__global__ void colminmax_kernel(const uint32_t* __restrict__ ja, uint64_t n, uint32_t* __restrict__ out) {
    uint32_t mn = 0xffffffffu, mx = 0;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t c = ja[i];
        mn = min(mn, c);
        mx = max(mx, c);
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, mn);
        atomicMax(out + 1, mx);
    }
}
extern "C" int spmvb200_col_range(const spmvb200_matrix* m, uint64_t* col_min, uint64_t* col_max) {
    if (!m || m->format != SPMVB200_FMT_CSR || !col_min || !col_max) return fail("col_range: needs a CSR handle");
    uint32_t* d = nullptr;
    CU_TRY(cudaMalloc(&d, 8));
    const uint32_t init[2] = {0xffffffffu, 0u};
    CU_TRY(cudaMemcpy(d, init, 8, cudaMemcpyHostToDevice));
    if (m->NZ) colminmax_kernel<<<1184, 256>>>(m->ja, m->NZ, d);
    uint32_t h[2];
    cudaError_t e = cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail("col_range: %s", cudaGetErrorString(e));
    *col_min = m->NZ ? h[0] : 0;
    *col_max = m->NZ ? h[1] : 0;
    return 0;
}

extern "C" int spmvb200_csr_download(const spmvb200_matrix* m, uint64_t* irp, uint64_t* ja, double* as) {
    if (!m || m->format != SPMVB200_FMT_CSR) return fail("csr_download: not a CSR handle");
    uint64_t* tmp = nullptr;
    const uint64_t n = std::max<uint64_t>(m->NZ, m->M + 1);
    CU_TRY(cudaMalloc(&tmp, n * 8));
    int rc = 0;
    if (irp) {
        widen_u32_kernel<<<1184, 256>>>(m->irp, tmp, m->M + 1, 0);
        rc |= cudaMemcpy(irp, tmp, (m->M + 1) * 8, cudaMemcpyDeviceToHost) != cudaSuccess;
    }
    if (ja && m->NZ) {
        widen_u32_kernel<<<1184, 256>>>(m->ja, tmp, m->NZ, 0);
        rc |= cudaMemcpy(ja, tmp, m->NZ * 8, cudaMemcpyDeviceToHost) != cudaSuccess;
    }
    if (as && m->NZ) rc |= cudaMemcpy(as, m->as, m->NZ * 8, cudaMemcpyDeviceToHost) != cudaSuccess;
    cudaFree(tmp);
    if (rc) return fail("csr_download: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// ------------------------------------------------------------------------------------------------- comparators (host arrays)
// The reference's check (doubleVectorsDiff, src/commons/utils.c:362-393) is an absolute 7e-4 on |x| < 3e-5 inputs and its NaN
// handling lets "never written" outputs pass (SURVEY.md §2.3-1,8).  The strict check a GPU kernel has to pass here:
//   |y_i - yref_i| <= tau * sum_j |a_ij x_j|   for every row, any NaN / Inf in y fails.   (pure host loops over the caller's arrays:
// a comparator, not a compute path -- the SpMV itself never runs on the CPU.)
extern "C" int spmvb200_compare_strict_csr(uint64_t M, const uint64_t* irp, const uint64_t* ja, const double* as, const double* x,
                                           const double* y_ref, const double* y, double tau, uint64_t* n_bad, double* worst_ratio) {
    if (!irp || !y_ref || !y || (irp[M] && (!ja || !as || !x))) return fail("compare_strict_csr: null argument");
    uint64_t bad = 0;
    double worst = 0.0;
#pragma omp parallel for reduction(+ : bad) reduction(max : worst) schedule(static)
    for (int64_t r = 0; r < (int64_t) M; ++r) {
        double scale = 0.0;
        for (uint64_t j = irp[r]; j < irp[r + 1]; ++j) scale += fabs(as[j] * x[ja[j]]);
        const double d = fabs(y[r] - y_ref[r]);
        if (!(d <= tau * scale) || y[r] != y[r] || fabs(y[r]) > 1.7976931348623157e308) {
            if (!(d == 0.0 && scale == 0.0)) ++bad;
        }
        const double ratio = scale > 0 ? d / scale : (d > 0 ? 1e300 : 0.0);
        if (ratio == ratio && ratio > worst) worst = ratio;
    }
    if (n_bad) *n_bad = bad;
    if (worst_ratio) *worst_ratio = worst;
    return 0;
}
// doubleVectorsDiff with the reference's threshold semantics but NaN-aware: returns the largest |a-b|, *failed = 1 if any difference
// exceeds `threshold` (the reference uses DOUBLE_DIFF_THREASH = 7e-4, src/include/config.h:101) or any entry is NaN.
extern "C" int spmvb200_compare_abs(uint64_t n, const double* a, const double* b, double threshold, int* failed, double* max_diff) {
    if ((n && (!a || !b)) || !failed) return fail("compare_abs: null argument");
    double mx = 0.0;
    int f = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const double d = fabs(a[i] - b[i]);
        if (d != d) { f = 1; continue; }
        if (d > mx) mx = d;
    }
    if (mx > threshold) f = 1;
    *failed = f;
    if (max_diff) *max_diff = mx;
    return 0;
}

// ------------------------------------------------------------------------------------------------- adapter cache
namespace {
struct CacheKey {
    const void* key;
    int fmt;
    bool operator<(const CacheKey& o) const { return key != o.key ? key < o.key : fmt < o.fmt; }
};
struct CacheVal {
    spmvb200_matrix* m;
    const void *ja, *as, *irp, *rl;
    uint64_t M, N, K, NZ, fp;
};
// Content fingerprint of the caller's host arrays: the row pointer (CSR) / row-length vector (every entry up to 2^20 rows) plus 4096 evenly spaced samples
// of JA and AS.  A freed-and-reallocated matrix of the same shape at the same addresses, or an in-place edit that touches the
// structure or any sampled entry, changes it; an edit confined to unsampled values does not -- such callers must call
// spmvb200_cache_drop(key) (stated in the header and in INTEGRATION.md).  ~2 us per 10^5 rows: noise next to the PCIe copies.
uint64_t host_fingerprint(uint64_t M, uint64_t slots, const uint64_t* irp, const uint64_t* ja, const double* as, const uint64_t* rl) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ M ^ (slots << 1);
    const uint64_t rstep = std::max<uint64_t>(1, M >> 20);  // every row up to 2^20 rows, ~10^6 evenly spaced ones beyond
    if (irp) {
        for (uint64_t r = 0; r <= M; r += rstep) h = (h ^ irp[r]) * 0x100000001B3ull;
        h = (h ^ irp[M]) * 0x100000001B3ull;
    }
    if (rl) for (uint64_t r = 0; r < M; r += rstep) h = (h ^ rl[r]) * 0x100000001B3ull;
    const uint64_t step = std::max<uint64_t>(1, slots / 4096);
    for (uint64_t j = 0; j < slots; j += step) {
        uint64_t bits;
        memcpy(&bits, as + j, 8);
        h = (h ^ ja[j]) * 0x100000001B3ull;
        h = (h ^ bits) * 0x100000001B3ull;
    }
    if (slots) {
        uint64_t bits;
        memcpy(&bits, as + slots - 1, 8);
        h = (h ^ ja[slots - 1] ^ bits) * 0x100000001B3ull;
    }
    return h;
}
std::map<CacheKey, CacheVal> g_cache;
std::mutex g_cache_mu;
}  // namespace

extern "C" int spmvb200_cached_spmv(const void* key, int kind, int is_ell, uint64_t M, uint64_t N, uint64_t K,
                                    const uint64_t* irp, const uint64_t* ja, const double* as, const uint64_t* rl,
                                    const double* x, double* y, double* elapsed_internal_s) {
    int fmt = SPMVB200_FMT_CSR;
    if (kind == SPMVB200_ELL_ROWS) fmt = SPMVB200_FMT_ELL_COLMAJOR;
    else if (kind == SPMVB200_ELL_ROWS_NT || kind == SPMVB200_ELL_ROWS_WARP_NT) fmt = SPMVB200_FMT_ELL_ROWMAJOR;
    else if (kind == SPMVB200_SELL_ROWS) fmt = SPMVB200_FMT_SELL;  // built on the device from the CSR input
    else if (kind == SPMVB200_XWIN_ROWS) fmt = SPMVB200_FMT_XWIN;  // likewise
    if ((fmt == SPMVB200_FMT_CSR || fmt == SPMVB200_FMT_SELL || fmt == SPMVB200_FMT_XWIN) == (is_ell != 0)) return fail("cached_spmv: kind %d does not match the %s input", kind, is_ell ? "ELL" : "CSR");
    if (!ja && M) return fail("cached_spmv: null JA");
    if (!as && M && (is_ell ? K != 0 : (irp && irp[M] != 0))) return fail("cached_spmv: null AS");
    if (!is_ell && !irp) return fail("cached_spmv: null IRP");
    const uint64_t NZ = is_ell ? M * K : irp[M];
    const uint64_t fp = host_fingerprint(M, NZ, irp, ja, as, is_ell ? rl : nullptr);
    spmvb200_matrix* m = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_cache.find({key, fmt});
        if (it != g_cache.end() && (it->second.ja != ja || it->second.as != as || it->second.irp != irp || it->second.rl != rl || it->second.M != M ||
                                    it->second.N != N || it->second.K != K || it->second.NZ != NZ || it->second.fp != fp)) {
            spmvb200_free(it->second.m);  // the host matrix behind this key changed
            g_cache.erase(it);
            it = g_cache.end();
        }
        if (it == g_cache.end()) {
            int rc = is_ell ? spmvb200_ell_upload(M, N, K, ja, as, rl, 0, M, fmt, &m) : spmvb200_csr_upload(M, N, irp, ja, as, 0, M, &m);
            if (rc) return 1;
            if (fmt == SPMVB200_FMT_SELL || fmt == SPMVB200_FMT_XWIN) {
                spmvb200_matrix* conv = nullptr;
                rc = fmt == SPMVB200_FMT_SELL ? spmvb200_sell_from_csr(m, 0, &conv) : spmvb200_xwin_from_csr(m, 0, 0, &conv);
                spmvb200_free(m);
                if (rc) return 1;
                m = conv;
            }
            g_cache[{key, fmt}] = {m, ja, as, irp, rl, M, N, K, NZ, fp};
        } else {
            m = it->second.m;
        }
    }
    float ms = 0;
    if (spmvb200_spmv_host(m, kind, x, y, &ms)) return 1;
    if (elapsed_internal_s) *elapsed_internal_s = (double) ms * 1e-3;
    return 0;
}

extern "C" int spmvb200_cache_drop(const void* key) {
    if (!key) spmvb200_host_unregister(nullptr);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto it = g_cache.begin(); it != g_cache.end();) {
        if (!key || it->first.key == key) {
            spmvb200_free(it->second.m);
            it = g_cache.erase(it);
        } else {
            ++it;
        }
    }
    return 0;
}
