// dropin/b200_cudaUtils.cu -- drop-in for src/commons/cudaUtils.cu: the H2D helpers declared in
// src/include/cudaUtils.h:60-68 (spMatCpyCSR, spMatCpyELL, spMatCpyELLNNPitched), same signatures, same
// ownership rules (the caller cudaMalloc's the device struct and later frees the members through a host copy
// with cudaFreeSpmat, src/main.cu:195,215,280), same return convention (EXIT_SUCCESS / EXIT_FAILURE with the
// members freed), producing the REFERENCE's device layout so that the reference kernels or
// dropin/b200_SpMV_CUDA.cu can consume it.  Compile with the reference's headers on the include path.
// Defects of the reference that are not reproduced: element sizes come from the pointee
// types, RL is allocated and copied with its real length (for a transposed ELL struct that is MAX_ROW_NZ rows).
//
// On top of that layout every uploader builds -- on the device, from the arrays it has just uploaded -- the narrow VIEW of
// dropin/b200_view.h (32-bit ids, SELL-32-sigma slices / column-major ELL, effective row lengths) and hides it where the
// reference's own cudaFreeSpmat releases it.  B200_DROPIN_PLAIN=1 in the environment skips the view (compatibility tier only).
#include <cub/cub.cuh>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "cudaUtils.h"
#include "b200_view.h"

namespace {
struct Upload {  // allocations of one upload; released together on failure
    void* ptr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    bool ok = true;
    const char* tag;
    explicit Upload(const char* t) : tag(t) {}
    template <typename T>
    T* flat(const T* host, size_t count, size_t extra_bytes = 0) {
        T* d = nullptr;
        if (!ok) return d;
        if (cudaErr(cudaMalloc(&d, sizeof(T) * (count ? count : 1) + extra_bytes), tag)) { ok = false; return nullptr; }
        ptr[n++] = d;
        if (count && cudaErr(cudaMemcpy(d, host, sizeof(T) * count, dirUp), tag)) ok = false;
        return d;
    }
    template <typename T>
    T* pitched(const T* host, size_t width, size_t height, size_t* pitch_elems) {
        T* d = nullptr;
        size_t pitch_bytes = 0;
        if (!ok) return d;
        if (cudaErr(cudaMallocPitch(&d, &pitch_bytes, sizeof(T) * (width ? width : 1), height ? height : 1), tag)) { ok = false; return nullptr; }
        ptr[n++] = d;
        if (width && height && cudaErr(cudaMemcpy2D(d, pitch_bytes, host, sizeof(T) * width, sizeof(T) * width, height, dirUp), tag)) ok = false;
        *pitch_elems = pitch_bytes / sizeof(T);
        return d;
    }
    int finish(spmat* dst, const spmat& local) {
        if (ok && cudaErr(cudaMemcpy(dst, &local, sizeof(local), dirUp), tag)) ok = false;
        if (!ok)
            for (int i = 0; i < n; ++i) cudaFree(ptr[i]);
        return ok ? EXIT_SUCCESS : EXIT_FAILURE;
    }
};

spmat header_of(const spmat* m) {
    spmat d;
    memset(&d, 0, sizeof(d));
    d.M = m->M;
    d.N = m->N;
    d.NZ = m->NZ;
    d.MAX_ROW_NZ = m->MAX_ROW_NZ;
    return d;
}
// ellTranspose swaps M <-> MAX_ROW_NZ and sets N = original M (src/commons/sparseUtils.c:168-171)
bool is_transposed_ell(const spmat* m) { return m->N == m->MAX_ROW_NZ && m->IRP == NULL; }
#ifdef ROWLENS
// the row-length vector of a transposed struct has MAX_ROW_NZ entries, not M
size_t rl_entries(const spmat* m) { return is_transposed_ell(m) ? m->MAX_ROW_NZ : m->M; }
#endif
bool want_view() {
    static const bool plain = getenv("B200_DROPIN_PLAIN") != nullptr;
    return !plain;
}
inline size_t up256(size_t b) { return (b + 255) & ~(size_t) 255; }

// ------------------------------------------------------------------------------------------------ device-side builders
__global__ void narrow_kernel(const ulong* __restrict__ src, unsigned* __restrict__ dst, size_t n, int* __restrict__ overflow) {
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const ulong v = src[i];
        if (v > 0xffffffffull) *overflow = 1;
        dst[i] = (unsigned) v;
    }
}
// sort key of row r: window (r / sigma) in the high word, 2^32-1 - length in the low word => ascending sort = descending
// length inside a window; rows longer than B200_LONG_ROW (and padding rows) count as empty here
__global__ void sell_keys_kernel(const unsigned* __restrict__ irp, unsigned M, unsigned Mpad, unsigned sigma, unsigned long long* __restrict__ keys,
                                 unsigned* __restrict__ vals, unsigned* __restrict__ long_rows, unsigned* __restrict__ nlong) {
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Mpad) return;
    unsigned len = r < M ? irp[r + 1] - irp[r] : 0u;
    const bool is_long = len > B200_LONG_ROW;
    if (is_long) {
        long_rows[atomicAdd(nlong, 1u)] = r;
        len = 0u;
    }
    keys[r] = ((unsigned long long) (r / sigma) << 32) | (unsigned long long) (0xffffffffu - len);
    vals[r] = (r < M && !is_long) ? r : 0xffffffffu;
}
__global__ void sell_slices_kernel(const unsigned long long* __restrict__ keys_sorted, unsigned Mpad, unsigned* __restrict__ rl_sorted,
                                   unsigned long long* __restrict__ slice_slots) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Mpad) return;
    const unsigned len = 0xffffffffu - (unsigned) (keys_sorted[i] & 0xffffffffull);
    rl_sorted[i] = len;
    if ((i & 31) == 0) slice_slots[i >> 5] = (unsigned long long) len * 32;  // first row of a slice is its longest
    if (i == 0) slice_slots[Mpad >> 5] = 0;
}
__global__ void narrow64_kernel(const unsigned long long* __restrict__ src, unsigned* __restrict__ dst, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (unsigned) src[i];
}
__global__ void sell_fill_kernel(const unsigned* __restrict__ irp, const unsigned* __restrict__ ja, const double* __restrict__ as,
                                 const unsigned* __restrict__ perm, const unsigned* __restrict__ slice_ptr, const unsigned* __restrict__ rl_sorted,
                                 unsigned Mpad, unsigned* __restrict__ sja, double* __restrict__ sas) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Mpad) return;
    const unsigned row = perm[i], len = rl_sorted[i];
    const unsigned sp0 = slice_ptr[i >> 5], wmax = (slice_ptr[(i >> 5) + 1] - sp0) >> 5;
    const unsigned s = (row != 0xffffffffu && len) ? irp[row] : 0u;
    for (unsigned k = 0; k < wmax; ++k) {
        const unsigned o = sp0 + k * 32 + (i & 31);
        sas[o] = k < len ? as[s + k] : 0.0;
        sja[o] = k < len ? ja[s + k] : 0u;
    }
}
// ELL: effective row length = last slot with a non-zero value + 1 (the reference pads with AS = 0, JA = 0: src/lib/parser.c:245-252)
__global__ void ell_rl_kernel(const double* __restrict__ as, const ulong* __restrict__ rl64, unsigned rows, unsigned K, size_t row_stride,
                              size_t slot_stride, unsigned* __restrict__ rl32) {
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    if (rl64) {
        rl32[r] = (unsigned) min((ulong) K, rl64[r]);
        return;
    }
    unsigned len = 0;
    for (unsigned k = 0; k < K; ++k)
        if (as[(size_t) r * row_stride + (size_t) k * slot_stride] != 0.0) len = k + 1;
    rl32[r] = len;
}
// ids (and optionally values) of an ELL struct -> column-major narrow arrays:  dst[k * pitch + r] = src[r * row_stride + k * slot_stride]
__global__ void ell_to_cm_kernel(const ulong* __restrict__ ja, const double* __restrict__ as, unsigned rows, unsigned K, size_t ja_row, size_t ja_slot,
                                 size_t as_row, size_t as_slot, size_t pitch, unsigned* __restrict__ ja32, double* __restrict__ as_cm,
                                 int* __restrict__ overflow) {
    __shared__ ulong tj[32][33];
    __shared__ double ta[32][33];
    // tile of 32 rows x 32 slots; reads run along whichever direction is contiguous in the source, writes along rows
    const unsigned r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const bool src_rowmajor = ja_slot == 1;
    for (unsigned j = threadIdx.y; j < 32; j += blockDim.y) {
        const unsigned r = src_rowmajor ? r0 + j : r0 + threadIdx.x, k = src_rowmajor ? k0 + threadIdx.x : k0 + j;
        if (r < rows && k < K) {
            const ulong c = ja[(size_t) r * ja_row + (size_t) k * ja_slot];
            if (c > 0xffffffffull) *overflow = 1;
            if (src_rowmajor) { tj[j][threadIdx.x] = c; if (as_cm) ta[j][threadIdx.x] = as[(size_t) r * as_row + (size_t) k * as_slot]; }
            else { tj[threadIdx.x][j] = c; if (as_cm) ta[threadIdx.x][j] = as[(size_t) r * as_row + (size_t) k * as_slot]; }
        }
    }
    __syncthreads();
    for (unsigned j = threadIdx.y; j < 32; j += blockDim.y) {  // tj[row in tile][slot in tile]
        const unsigned r = r0 + threadIdx.x, k = k0 + j;
        if (r < rows && k < K) {
            ja32[(size_t) k * pitch + r] = (unsigned) tj[threadIdx.x][j];
            if (as_cm) as_cm[(size_t) k * pitch + r] = ta[threadIdx.x][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------ CSR view
// Builds IRP' = [IRP (M+1 x 8 B) | B200ViewCSR | irp32 | ja32 | slice_ptr | perm | rl_sorted | long_rows | sja | sas] from the
// device arrays d_ja / d_as already uploaded; returns the new IRP allocation (or NULL: caller falls back to a plain IRP upload).
ulong* build_csr_view(const spmat* m, const ulong* d_ja, const double* d_as, size_t* hdr_addr) {
    const size_t M = m->M, NZ = m->NZ;
    if (M == 0 || M >= 0x7fffffffull || NZ >= 0xfffffff0ull || m->N > 0xffffffffull) return nullptr;
    const unsigned Mpad = (unsigned) ((M + 31) / 32 * 32), nsl = Mpad / 32, sigma = 16384;
    unsigned *t_irp32 = nullptr, *t_vals = nullptr, *t_perm = nullptr, *t_rl = nullptr, *t_long = nullptr, *t_cnt = nullptr;
    unsigned long long *t_k0 = nullptr, *t_k1 = nullptr, *t_slots = nullptr, *t_scan = nullptr;
    ulong* t_irp64 = nullptr;
    int* t_of = nullptr;
    void* tmp = nullptr;
    ulong* out = nullptr;
    bool ok = false;
    do {
        if (cudaMalloc(&t_irp64, (M + 1) * 8) || cudaMalloc(&t_irp32, (M + 1) * 4) || cudaMalloc(&t_vals, (size_t) Mpad * 4) ||
            cudaMalloc(&t_perm, (size_t) Mpad * 4) || cudaMalloc(&t_rl, (size_t) Mpad * 4) || cudaMalloc(&t_long, M * 4) ||
            cudaMalloc(&t_cnt, 4) || cudaMalloc(&t_k0, (size_t) Mpad * 8) || cudaMalloc(&t_k1, (size_t) Mpad * 8) ||
            cudaMalloc(&t_slots, ((size_t) nsl + 1) * 8) || cudaMalloc(&t_scan, ((size_t) nsl + 1) * 8) || cudaMalloc(&t_of, 4))
            break;
        if (cudaMemcpy(t_irp64, m->IRP, (M + 1) * 8, dirUp) || cudaMemset(t_cnt, 0, 4) || cudaMemset(t_of, 0, 4)) break;
        narrow_kernel<<<592, 256>>>(t_irp64, t_irp32, M + 1, t_of);
        sell_keys_kernel<<<(Mpad + 255) / 256, 256>>>(t_irp32, (unsigned) M, Mpad, sigma, t_k0, t_vals, t_long, t_cnt);
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, t_k0, t_k1, t_vals, t_perm, (int) Mpad);
        cub::DeviceScan::ExclusiveSum(nullptr, b2, t_slots, t_scan, (int) nsl + 1);
        if (cudaMalloc(&tmp, (b1 > b2 ? b1 : b2) + 16)) break;
        if (cub::DeviceRadixSort::SortPairs(tmp, b1, t_k0, t_k1, t_vals, t_perm, (int) Mpad) != cudaSuccess) break;
        sell_slices_kernel<<<(Mpad + 255) / 256, 256>>>(t_k1, Mpad, t_rl, t_slots);
        if (cub::DeviceScan::ExclusiveSum(tmp, b2, t_slots, t_scan, (int) nsl + 1) != cudaSuccess) break;
        unsigned long long slots = 0;
        unsigned nlong = 0;
        int of = 0;
        if (cudaMemcpy(&slots, t_scan + nsl, 8, dirDown) || cudaMemcpy(&nlong, t_cnt, 4, dirDown) || cudaMemcpy(&of, t_of, 4, dirDown)) break;
        if (of || slots >= 0xfffffff0ull) break;
        // slices must stay nearly padding-free to be worth a second copy of the values
        if ((double) slots > 1.5 * (double) NZ + 1024.0) break;
        // carve the final allocation
        size_t off = up256((M + 1) * 8);
        const size_t o_hdr = off;        off = up256(off + sizeof(B200ViewCSR));
        const size_t o_irp32 = off;      off = up256(off + (M + 1) * 4);
        const size_t o_ja32 = off;       off = up256(off + (NZ + 16) * 4);
        const size_t o_sp = off;         off = up256(off + ((size_t) nsl + 1) * 4);
        const size_t o_perm = off;       off = up256(off + (size_t) Mpad * 4);
        const size_t o_rl = off;         off = up256(off + (size_t) Mpad * 4);
        const size_t o_long = off;       off = up256(off + ((size_t) nlong + 1) * 4);
        const size_t o_sja = off;        off = up256(off + ((size_t) slots + 16) * 4);
        const size_t o_sas = off;        off = up256(off + ((size_t) slots + 16) * 8);
        char* base = nullptr;
        if (cudaMalloc(&base, off)) break;
        out = reinterpret_cast<ulong*>(base);
        B200ViewCSR h;
        memset(&h, 0, sizeof(h));
        h.magic = B200_VIEW_MAGIC;
        h.Mpad = Mpad;
        h.nslices = nsl;
        h.nlong = nlong;
        const double mean = (double) NZ / (double) M;
        h.lanes = 2;
        while (h.lanes < 32 && h.lanes * 2 < mean) h.lanes *= 2;
        h.irp32 = reinterpret_cast<unsigned*>(base + o_irp32);
        h.ja32 = reinterpret_cast<unsigned*>(base + o_ja32);
        h.slice_ptr = reinterpret_cast<unsigned*>(base + o_sp);
        h.perm = reinterpret_cast<unsigned*>(base + o_perm);
        h.rl_sorted = reinterpret_cast<unsigned*>(base + o_rl);
        h.long_rows = reinterpret_cast<unsigned*>(base + o_long);
        h.sja = reinterpret_cast<unsigned*>(base + o_sja);
        h.sas = reinterpret_cast<double*>(base + o_sas);
        if (cudaMemcpy(base, t_irp64, (M + 1) * 8, cudaMemcpyDeviceToDevice) || cudaMemcpy(base + o_hdr, &h, sizeof(h), dirUp) ||
            cudaMemcpy(base + o_irp32, t_irp32, (M + 1) * 4, cudaMemcpyDeviceToDevice) ||
            cudaMemcpy(base + o_perm, t_perm, (size_t) Mpad * 4, cudaMemcpyDeviceToDevice) ||
            cudaMemcpy(base + o_rl, t_rl, (size_t) Mpad * 4, cudaMemcpyDeviceToDevice) ||
            (nlong && cudaMemcpy(base + o_long, t_long, (size_t) nlong * 4, cudaMemcpyDeviceToDevice)) ||
            cudaMemset(base + o_ja32 + NZ * 4, 0, 64) || cudaMemset(base + o_sja + slots * 4, 0, 64) || cudaMemset(base + o_sas + slots * 8, 0, 128))
            break;
        narrow64_kernel<<<(nsl + 1 + 255) / 256, 256>>>(t_scan, reinterpret_cast<unsigned*>(base + o_sp), nsl + 1);
        if (NZ) narrow_kernel<<<1184, 256>>>(d_ja, reinterpret_cast<unsigned*>(base + o_ja32), NZ, t_of);
        sell_fill_kernel<<<(Mpad + 255) / 256, 256>>>(h.irp32, h.ja32, d_as, h.perm, h.slice_ptr, h.rl_sorted, Mpad,
                                                       reinterpret_cast<unsigned*>(base + o_sja), reinterpret_cast<double*>(base + o_sas));
        if (cudaDeviceSynchronize() || cudaMemcpy(&of, t_of, 4, dirDown) || of) break;
        *hdr_addr = reinterpret_cast<size_t>(base + o_hdr);
        ok = true;
    } while (0);
    cudaFree(t_irp64); cudaFree(t_irp32); cudaFree(t_vals); cudaFree(t_perm); cudaFree(t_rl); cudaFree(t_long); cudaFree(t_cnt);
    cudaFree(t_k0); cudaFree(t_k1); cudaFree(t_slots); cudaFree(t_scan); cudaFree(t_of); cudaFree(tmp);
    if (!ok) {
        cudaFree(out);
        cudaGetLastError();
        return nullptr;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------ ELL view
// d is the device-layout struct being prepared (JA / AS uploaded, pitches set).  transposed: the struct is column-major already
// (as[k * pitchAS + row]) and only ids + lengths are narrowed; otherwise a column-major copy of ids AND values is built.
ulong* build_ell_view(const spmat* host, const spmat& d, bool transposed, const ulong* d_rl) {
    const size_t rows = transposed ? host->MAX_ROW_NZ : host->M, K = transposed ? host->M : host->MAX_ROW_NZ;
    if (rows == 0 || K == 0 || rows >= 0x7fffffffull || K > 0xffffffffull || host->N > 0xffffffffull) return nullptr;
    const size_t pitch = (rows + 63) / 64 * 64;
    size_t off = up256(sizeof(B200ViewELL));
    const size_t o_rl = off;   off = up256(off + rows * 4);
    const size_t o_ja = off;   off = up256(off + (K * pitch + 16) * 4);
    const size_t o_as = off;   if (!transposed) off = up256(off + (K * pitch + 16) * 8);
    char* base = nullptr;
    int* t_of = nullptr;
    if (cudaMalloc(&base, off)) { cudaGetLastError(); return nullptr; }
    bool ok = false;
    do {
        if (cudaMalloc(&t_of, 4) || cudaMemset(t_of, 0, 4) || cudaMemset(base, 0, off)) break;
        B200ViewELL h;
        memset(&h, 0, sizeof(h));
        h.magic = B200_VIEW_MAGIC;
        h.rows = (unsigned) rows;
        h.K = (unsigned) K;
        h.pitch = pitch;
        h.rl32 = reinterpret_cast<unsigned*>(base + o_rl);
        h.ja32 = reinterpret_cast<unsigned*>(base + o_ja);
        h.as_cm = transposed ? nullptr : reinterpret_cast<double*>(base + o_as);
        if (cudaMemcpy(base, &h, sizeof(h), dirUp)) break;
        // source strides in elements: transposed struct = slot-major rows of the pitched 2-D allocation
        const size_t ja_row = transposed ? 1 : d.pitchJA, ja_slot = transposed ? d.pitchJA : 1;
        const size_t as_row = transposed ? 1 : d.pitchAS, as_slot = transposed ? d.pitchAS : 1;
        ell_rl_kernel<<<(unsigned) ((rows + 255) / 256), 256>>>(d.AS, d_rl, (unsigned) rows, (unsigned) K, as_row, as_slot,
                                                                 reinterpret_cast<unsigned*>(base + o_rl));
        dim3 grid((unsigned) ((rows + 31) / 32), (unsigned) ((K + 31) / 32));
        ell_to_cm_kernel<<<grid, dim3(32, 8)>>>(d.JA, d.AS, (unsigned) rows, (unsigned) K, ja_row, ja_slot, as_row, as_slot, pitch,
                                                reinterpret_cast<unsigned*>(base + o_ja), transposed ? nullptr : reinterpret_cast<double*>(base + o_as), t_of);
        int of = 0;
        if (cudaDeviceSynchronize() || cudaMemcpy(&of, t_of, 4, dirDown) || of) break;
        ok = true;
    } while (0);
    cudaFree(t_of);
    if (!ok) {
        cudaFree(base);
        cudaGetLastError();
        return nullptr;
    }
    return reinterpret_cast<ulong*>(base);
}
}  // namespace

int spMatCpyCSR(spmat* m, spmat* dst) {
    Upload up("spMatCpyCSR");
    spmat d = header_of(m);
    d.JA = up.flat(m->JA, m->NZ);
    d.AS = up.flat(m->AS, m->NZ, 128);  // slack: the sub-warp kernel's 16-byte loads may overrun the last entry
    size_t hdr = 0;
    ulong* irp_view = (up.ok && want_view()) ? build_csr_view(m, d.JA, d.AS, &hdr) : nullptr;
    if (irp_view) {
        d.IRP = irp_view;
        up.ptr[up.n++] = irp_view;
        d.pitchJA = (size_t) B200_VIEW_MAGIC;  // unused for CSR, zero after the reference's own uploader: the view's marker
        d.pitchAS = hdr;
    } else {
        d.IRP = up.flat(m->IRP, m->M + 1);
    }
#ifdef ROWLENS
    d.RL = up.flat(m->RL, m->M);
#endif
    return up.finish(dst, d);
}

static int ell_upload_common(spmat* m, spmat* dst, bool pitched, const char* tag) {
    Upload up(tag);
    spmat d = header_of(m);
    if (pitched) {
        // @m is row-major M x MAX_ROW_NZ (after ellTranspose: K x rows) -> pitched 2-D allocations, pitch in elements
        d.JA = up.pitched(m->JA, m->MAX_ROW_NZ, m->M, &d.pitchJA);
        d.AS = up.pitched(m->AS, m->MAX_ROW_NZ, m->M, &d.pitchAS);
    } else {
        d.JA = up.flat(m->JA, m->MAX_ROW_NZ * m->M);
        d.AS = up.flat(m->AS, m->MAX_ROW_NZ * m->M);
        d.pitchJA = d.pitchAS = m->MAX_ROW_NZ;
    }
    const ulong* d_rl = nullptr;
#ifdef ROWLENS
    d.RL = up.flat(m->RL, rl_entries(m));
    d_rl = d.RL;
#endif
    if (up.ok && want_view()) {
        ulong* view = build_ell_view(m, d, is_transposed_ell(m), d_rl);
        if (view) {
            d.IRP = view;  // unused for ELL and NULL after the reference's uploader; cudaFreeSpmat cudaFree()s it
            up.ptr[up.n++] = view;
        }
    }
    return up.finish(dst, d);
}

int spMatCpyELL(spmat* m, spmat* dst) { return ell_upload_common(m, dst, true, "spMatCpyELL"); }

int spMatCpyELLNNPitched(spmat* m, spmat* dst) { return ell_upload_common(m, dst, false, "spMatCpyELLNNPitched"); }
