// synth.cu -- seeded synthetic matrices of the BASELINE.json configs, on the host (reference
// layout, 64-bit indices, OpenMP) and on the device (engine layout, 32-bit indices).
// Every row (R-MAT: every edge) is a pure function of (seed, index): host and device agree bit for
// bit and any row range can be produced independently (multi-GPU partitions, bounded CPU samples).
// Not part of the reference (it reads Matrix Market text, src/lib/parser.c); 1e9 non-zeros cannot
// go through a text file.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "../../include/spmv_b200.h"
#include "engine.h"
#include "common.cuh"

using namespace spmvb200;

namespace {

constexpr uint64_t VAL_SALT = 0xA5A5A5A55A5A5A5Aull;

struct Dims {
    uint64_t M, N;
};

__host__ __device__ inline Dims synth_dims(const spmvb200_synth& s) {
    switch (s.kind) {
        case 1: return {s.p0 * s.p0, s.p0 * s.p0};
        case 2: return {s.p0 * s.p1 * s.p2, s.p0 * s.p1 * s.p2};
        case 4: return {s.p2, s.p2};
        case 5: return {s.p0, s.p0};
        default: return {0, 0};
    }
}

__host__ __device__ inline uint32_t synth_row_len(const spmvb200_synth& s, uint64_t r) {
    switch (s.kind) {
        case 1: {
            const uint64_t n = s.p0, iy = r / n, ix = r % n;
            return 1 + (iy > 0) + (ix > 0) + (ix + 1 < n) + (iy + 1 < n);
        }
        case 2: {
            const uint64_t nx = s.p0, ny = s.p1, nz = s.p2;
            const uint64_t ix = r % nx, iy = (r / nx) % ny, iz = r / (nx * ny);
            const uint32_t cx = 1 + (ix > 0) + (ix + 1 < nx), cy = 1 + (iy > 0) + (iy + 1 < ny), cz = 1 + (iz > 0) + (iz + 1 < nz);
            return cx * cy * cz;
        }
        case 4: {
            const uint64_t M = s.p2, w = s.p0;
            const uint64_t lo = r > w ? r - w : 0, hi = r + w < M ? r + w : M - 1;
            const uint64_t W = hi - lo + 1;
            return (uint32_t) (s.p1 < W ? s.p1 : W);
        }
        case 5: {
            const bool is_long = (hash2(s.seed, r) >> 32) < s.p3;
            const uint64_t len = is_long ? s.p1 : 4;
            return (uint32_t) (len < s.p0 ? len : s.p0);
        }
        default: return 0;
    }
}

// k strictly increasing columns in [lo, lo+W): one uniform draw from each of k equal strata
template <typename IdxT>
__host__ __device__ inline void stratified_row(uint64_t seed, uint64_t r, uint64_t lo, uint64_t W, uint32_t k, IdxT* ja, double* as) {
    for (uint32_t j = 0; j < k; ++j) {
        const uint64_t s_lo = lo + (W * j) / k, s_hi = lo + (W * (j + 1)) / k;  // non-empty because k <= W
        ja[j] = (IdxT) (s_lo + urange(hash3(seed, r, j), s_hi - s_lo));
        as[j] = usym(hash3(seed ^ VAL_SALT, r, j));
    }
}

template <typename IdxT>
__host__ __device__ inline void synth_row_fill(const spmvb200_synth& s, uint64_t r, IdxT* ja, double* as) {
    switch (s.kind) {
        case 1: {
            const uint64_t n = s.p0, iy = r / n, ix = r % n;
            uint32_t k = 0;
            if (iy > 0) { ja[k] = (IdxT) (r - n); as[k++] = -1.0; }
            if (ix > 0) { ja[k] = (IdxT) (r - 1); as[k++] = -1.0; }
            ja[k] = (IdxT) r; as[k++] = 4.0;
            if (ix + 1 < n) { ja[k] = (IdxT) (r + 1); as[k++] = -1.0; }
            if (iy + 1 < n) { ja[k] = (IdxT) (r + n); as[k++] = -1.0; }
            break;
        }
        case 2: {
            const int64_t nx = (int64_t) s.p0, ny = (int64_t) s.p1, nz = (int64_t) s.p2;
            const int64_t ix = (int64_t) (r % nx), iy = (int64_t) ((r / nx) % ny), iz = (int64_t) (r / (nx * ny));
            uint32_t k = 0;
            for (int dz = -1; dz <= 1; ++dz) {
                if (iz + dz < 0 || iz + dz >= nz) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    if (iy + dy < 0 || iy + dy >= ny) continue;
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (ix + dx < 0 || ix + dx >= nx) continue;
                        ja[k] = (IdxT) ((int64_t) r + (dz * ny + dy) * nx + dx);
                        as[k++] = (dz == 0 && dy == 0 && dx == 0) ? 26.0 : -1.0;
                    }
                }
            }
            break;
        }
        case 4: {
            const uint64_t M = s.p2, w = s.p0;
            const uint64_t lo = r > w ? r - w : 0, hi = r + w < M ? r + w : M - 1;
            stratified_row<IdxT>(s.seed, r, lo, hi - lo + 1, synth_row_len(s, r), ja, as);
            break;
        }
        case 5: stratified_row<IdxT>(s.seed, r, 0, s.p0, synth_row_len(s, r), ja, as); break;
        default: break;
    }
}

__host__ __device__ inline uint64_t rmat_key(int scale, uint64_t seed, uint64_t e) {
    uint64_t row = 0, col = 0;
    for (int l = 0; l < scale; ++l) {
        const double u = u01(hash3(seed, e, (uint64_t) l));
        // quadrant probabilities a,b,c,d = .57,.19,.19,.05
        const int rb = u >= 0.76, cb = (u >= 0.57 && u < 0.76) || u >= 0.95;
        row = (row << 1) | (uint64_t) rb;
        col = (col << 1) | (uint64_t) cb;
    }
    return (row << 32) | col;
}
__host__ __device__ inline double rmat_value(uint64_t seed, uint64_t key) { return usym(hash3(seed ^ VAL_SALT, key >> 32, key & 0xffffffffull)); }

int check_synth(const spmvb200_synth* s) {
    if (!s) return fail("synth: null descriptor");
    const Dims d = synth_dims(*s);
    if (d.M == 0) return fail("synth: unknown kind %d or empty dimensions", s->kind);
    if (s->kind == 4 && (s->p1 == 0 || s->p1 > 2 * s->p0 + 1)) return fail("synth banded: nnz/row %llu must be in [1, 2w+1]", (unsigned long long) s->p1);
    if (s->kind == 5 && s->p1 == 0) return fail("synth mixed: K_max must be >= 1");
    return 0;
}

// ----------------------------------------------------------------------------- device kernels
__global__ void rowlen_kernel(spmvb200_synth s, uint64_t row_begin, uint32_t rows, uint32_t* len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) len[i] = synth_row_len(s, row_begin + i);
    if (i == rows) len[i] = 0;
}
__global__ void fill_kernel(spmvb200_synth s, uint64_t row_begin, uint32_t rows, const uint32_t* __restrict__ irp, uint32_t* ja, double* as) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) synth_row_fill<uint32_t>(s, row_begin + i, ja + irp[i], as + irp[i]);
}
__global__ void rmat_keys_kernel(int scale, uint64_t seed, uint64_t n, uint64_t* keys) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) keys[e] = rmat_key(scale, seed, e);
}
__global__ void rmat_finish_kernel(uint64_t seed, uint64_t n, const uint64_t* __restrict__ keys, uint32_t* ja, double* as, uint32_t* rowcount) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i];
        ja[i] = (uint32_t) (k & 0xffffffffull);
        as[i] = rmat_value(seed, k);
        atomicAdd(rowcount + (k >> 32), 1u);
    }
}
__global__ void vector_kernel(uint64_t seed, uint64_t begin, uint64_t n, double scale, double* x) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = usym(hash2(seed, begin + i)) * scale;
}

int exclusive_scan_u32(uint32_t* d_in, uint32_t* d_out, size_t n) {
    void* tmp = nullptr;
    size_t bytes = 0;
    CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_in, d_out, n));
    CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 16));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, d_in, d_out, n);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail("device scan failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace

// ----------------------------------------------------------------------------- C ABI
extern "C" int spmvb200_synth_dims(const spmvb200_synth* s, uint64_t* M, uint64_t* N) {
    if (check_synth(s)) return 1;
    const Dims d = synth_dims(*s);
    if (M) *M = d.M;
    if (N) *N = d.N;
    return 0;
}

extern "C" int spmvb200_synth_rowlen_host(const spmvb200_synth* s, uint64_t row_begin, uint64_t row_end, uint64_t* rl) {
    if (check_synth(s)) return 1;
    if (row_begin > row_end || row_end > synth_dims(*s).M || !rl) return fail("synth_rowlen_host: bad row range");
    const spmvb200_synth c = *s;
#pragma omp parallel for schedule(static)
    for (int64_t r = (int64_t) row_begin; r < (int64_t) row_end; ++r) rl[r - row_begin] = synth_row_len(c, (uint64_t) r);
    return 0;
}

extern "C" int spmvb200_synth_fill_host(const spmvb200_synth* s, uint64_t row_begin, uint64_t row_end, const uint64_t* irp_local,
                                        uint64_t* ja, double* as) {
    if (check_synth(s)) return 1;
    if (row_begin > row_end || row_end > synth_dims(*s).M || !irp_local || !ja || !as) return fail("synth_fill_host: bad arguments");
    const spmvb200_synth c = *s;
#pragma omp parallel for schedule(static)
    for (int64_t r = (int64_t) row_begin; r < (int64_t) row_end; ++r) {
        const uint64_t o = irp_local[r - row_begin];
        synth_row_fill<uint64_t>(c, (uint64_t) r, ja + o, as + o);
    }
    return 0;
}

extern "C" int spmvb200_synth_csr_device(const spmvb200_synth* s, uint64_t row_begin, uint64_t row_end, spmvb200_matrix** out) {
    if (!out) return fail("synth_csr_device: null output");
    *out = nullptr;
    if (check_synth(s)) return 1;
    const Dims d = synth_dims(*s);
    if (row_begin > row_end || row_end > d.M) return fail("synth_csr_device: bad row range");
    const uint64_t rows = row_end - row_begin;
    if (rows >= 0x7fffffffull) return fail("synth_csr_device: too many rows");
    uint32_t *len = nullptr, *irp = nullptr, *ja = nullptr;
    double* as = nullptr;
    int rc = 0;
    do {
        if ((rc = cudaMalloc(&len, (rows + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&irp, (rows + 1) * 4) != cudaSuccess)) break;
        rowlen_kernel<<<(unsigned) ((rows + 1 + 255) / 256), 256>>>(*s, row_begin, (uint32_t) rows, len);
        if ((rc = exclusive_scan_u32(len, irp, rows + 1))) break;
        uint32_t nz = 0;
        if ((rc = cudaMemcpy(&nz, irp + rows, 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        // the 32-bit scan must not have wrapped: compare with a 64-bit host-side bound
        {
            uint64_t maxlen = s->kind == 1 ? 5 : s->kind == 2 ? 27 : s->kind == 4 ? s->p1 : std::max<uint64_t>(4, s->p1);
            if (rows * maxlen >= 0xfffffff0ull && s->kind != 4) { rc = fail("synth_csr_device: nnz may exceed 32 bits"); break; }
            if (s->kind == 4 && rows * s->p1 >= 0xfffffff0ull) { rc = fail("synth_csr_device: nnz exceeds 32 bits"); break; }
        }
        if ((rc = cudaMalloc(&ja, ((uint64_t) nz + PAD) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&as, ((uint64_t) nz + PAD) * 8) != cudaSuccess)) break;
        cudaMemset(ja + nz, 0, PAD * 4);
        cudaMemset(as + nz, 0, PAD * 8);
        if (rows) fill_kernel<<<(unsigned) ((rows + 255) / 256), 256>>>(*s, row_begin, (uint32_t) rows, irp, ja, as);
        if ((rc = cudaDeviceSynchronize() != cudaSuccess)) break;
        rc = spmvb200_csr_adopt_device(rows, d.N, nz, irp, ja, as, 1, out);
    } while (0);
    cudaFree(len);
    if (rc) {
        if (cudaPeekAtLastError() != cudaSuccess) fail("synth_csr_device: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(irp);
        cudaFree(ja);
        cudaFree(as);
        return 1;
    }
    return 0;
}

extern "C" int spmvb200_synth_rmat_keys_host(int scale, uint64_t seed, uint64_t e_begin, uint64_t e_end, uint64_t* keys) {
    if (scale < 1 || scale > 31 || e_begin > e_end || !keys) return fail("synth_rmat_keys_host: bad arguments");
#pragma omp parallel for schedule(static)
    for (int64_t e = (int64_t) e_begin; e < (int64_t) e_end; ++e) keys[e - e_begin] = rmat_key(scale, seed, (uint64_t) e);
    return 0;
}
extern "C" int spmvb200_synth_rmat_values_host(uint64_t seed, uint64_t n, const uint64_t* keys, double* as) {
    if (!keys || !as) return fail("synth_rmat_values_host: null argument");
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t) n; ++i) as[i] = rmat_value(seed, keys[i]);
    return 0;
}

extern "C" int spmvb200_synth_rmat_csr_device(int scale, uint64_t n_edges, uint64_t seed, spmvb200_matrix** out) {
    if (!out) return fail("synth_rmat_csr_device: null output");
    *out = nullptr;
    if (scale < 1 || scale > 30 || n_edges == 0 || n_edges >= 0xfffffff0ull) return fail("synth_rmat_csr_device: bad arguments");
    const uint64_t M = 1ull << scale;
    uint64_t *k0 = nullptr, *k1 = nullptr, *d_num = nullptr;
    uint32_t *cnt = nullptr, *irp = nullptr, *ja = nullptr;
    double* as = nullptr;
    void* tmp = nullptr;
    int rc = 0;
    do {
        if ((rc = cudaMalloc(&k0, n_edges * 8) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&k1, n_edges * 8) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&d_num, 8) != cudaSuccess)) break;
        rmat_keys_kernel<<<2368, 256>>>(scale, seed, n_edges, k0);
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, b1, k0, k1, n_edges, 0, 32 + scale);
        cub::DeviceSelect::Unique(nullptr, b2, k1, k0, d_num, n_edges);
        if ((rc = cudaMalloc(&tmp, std::max(b1, b2) + 16) != cudaSuccess)) break;
        if ((rc = cub::DeviceRadixSort::SortKeys(tmp, b1, k0, k1, n_edges, 0, 32 + scale) != cudaSuccess)) break;
        if ((rc = cub::DeviceSelect::Unique(tmp, b2, k1, k0, d_num, n_edges) != cudaSuccess)) break;
        uint64_t nz = 0;
        if ((rc = cudaMemcpy(&nz, d_num, 8, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        cudaFree(k1); k1 = nullptr;
        cudaFree(tmp); tmp = nullptr;
        if ((rc = cudaMalloc(&cnt, (M + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&irp, (M + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&ja, (nz + PAD) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&as, (nz + PAD) * 8) != cudaSuccess)) break;
        cudaMemset(cnt, 0, (M + 1) * 4);
        cudaMemset(ja + nz, 0, PAD * 4);
        cudaMemset(as + nz, 0, PAD * 8);
        rmat_finish_kernel<<<2368, 256>>>(seed, nz, k0, ja, as, cnt);
        if ((rc = exclusive_scan_u32(cnt, irp, M + 1))) break;
        if ((rc = cudaDeviceSynchronize() != cudaSuccess)) break;
        rc = spmvb200_csr_adopt_device(M, M, nz, irp, ja, as, 1, out);
    } while (0);
    cudaFree(k0);
    cudaFree(k1);
    cudaFree(d_num);
    cudaFree(cnt);
    cudaFree(tmp);
    if (rc) {
        if (cudaPeekAtLastError() != cudaSuccess) fail("synth_rmat_csr_device: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(irp);
        cudaFree(ja);
        cudaFree(as);
        return 1;
    }
    return 0;
}

extern "C" int spmvb200_synth_vector_host(uint64_t seed, uint64_t begin, uint64_t end, double scale, double* x) {
    if (begin > end || !x) return fail("synth_vector_host: bad arguments");
#pragma omp parallel for schedule(static)
    for (int64_t i = (int64_t) begin; i < (int64_t) end; ++i) x[i - begin] = usym(hash2(seed, (uint64_t) i)) * scale;
    return 0;
}
extern "C" int spmvb200_synth_vector_device(uint64_t seed, uint64_t begin, uint64_t end, double scale, double* d_x) {
    if (begin > end || !d_x) return fail("synth_vector_device: bad arguments");
    if (end > begin) vector_kernel<<<1184, 256>>>(seed, begin, end - begin, scale, d_x);
    CU_TRY(cudaPeekAtLastError());
    return 0;
}
