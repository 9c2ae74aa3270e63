// engine_formats.inl -- part of engine.cu (textually included there: one translation unit, file-local helpers stay static).
// formats: CSR row-block plan, uploads with on-device narrowing / ELL transposition, SELL-32-sigma and x-window construction.

// ------------------------------------------------------------------------------------------------- CSR plan
static int scan_u32(void* tmp, size_t tmp_bytes, const uint32_t* in, uint32_t* out, size_t n) {
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int) n);
    if (e != cudaSuccess) return fail("plan scan: %s", cudaGetErrorString(e));
    return 0;
}

// Builds, entirely on the device (plan.cuh), the row-block tiles of the stream kernel, the long-row records and
// segment list, the medium-row list and the contiguous spans of the vector kernels.
int spmvb200::finish_csr(spmvb200_matrix* m) {
    const uint32_t M = (uint32_t) m->M;
    const size_t n1 = (size_t) M + 1;
    uint32_t last = 0;
    CU_TRY(cudaMemcpy(&last, m->irp + M, 4, cudaMemcpyDeviceToHost));
    if (last != m->NZ) return fail("CSR row pointer inconsistent: IRP[M]=%u, NZ=%llu", last, (unsigned long long) m->NZ);
    uint32_t *cnt = nullptr, *idx = nullptr, *d_num = nullptr;  // cnt: tiles_at | long_at | segs_at ; idx: their scans
    void* tmp = nullptr;
    uint32_t *h_r0 = nullptr, *h_n0 = nullptr;
    int rc = 0;
    do {
        if ((rc = cudaMalloc(&cnt, 3 * n1 * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&idx, 3 * n1 * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&d_num, 16) != cudaSuccess)) break;
        size_t b_scan = 0, b_sel = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b_scan, cnt, idx, (int) n1);
        thrust::counting_iterator<uint32_t> rows_it(0);
        MidRowPred pred{m->irp};
        cub::DeviceSelect::If(nullptr, b_sel, rows_it, cnt, d_num, (int) M, pred);
        const size_t tmp_bytes = std::max(b_scan, b_sel) + 16;
        if ((rc = cudaMalloc(&tmp, tmp_bytes) != cudaSuccess)) break;
        uint32_t lmax = 0;
        if (M) {
            cudaMemset(d_num, 0, 4);
            row_len_from_irp_kernel<<<(M + 255) / 256, 256>>>(m->irp, M, nullptr, d_num);
            if ((rc = cudaMemcpy(&lmax, d_num, 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        }
        m->lmax = lmax;
        const uint32_t special = std::max<uint32_t>(1, std::min<uint32_t>(PLAN_SPECIAL_MAX, lmax));
        plan_count_kernel<<<(unsigned) ((n1 + 255) / 256), 256>>>(m->irp, M, special, cnt, cnt + n1, cnt + 2 * n1);
        for (int a = 0; a < 3 && !rc; ++a) rc = scan_u32(tmp, tmp_bytes, cnt + a * n1, idx + a * n1, n1);
        if (rc) break;
        uint32_t tot[3];
        for (int a = 0; a < 3; ++a)
            if ((rc = cudaMemcpy(&tot[a], idx + a * n1 + M, 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        if (rc) break;
        m->ntiles = tot[0];
        m->nlong = tot[1];
        m->nseg = tot[2];
        if ((rc = cudaMalloc(&m->desc, ((size_t) m->ntiles + 1) * sizeof(TileDesc)) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->longrec, std::max<size_t>(1, m->nlong) * sizeof(LongRec)) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->seg_tiles, std::max<size_t>(1, m->nseg) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->partial, std::max<size_t>(1, m->ntiles) * sizeof(double)) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->ticket, std::max<size_t>(1, m->nlong) * 4) != cudaSuccess)) break;
        cudaMemset(m->ticket, 0, std::max<size_t>(1, m->nlong) * 4);
        plan_scatter_kernel<<<(unsigned) ((n1 + 255) / 256), 256>>>(m->irp, M, (uint32_t) m->NZ, cnt, idx, idx + n1, idx + 2 * n1, m->ntiles, m->desc,
                                                                   m->longrec, m->seg_tiles);
        // rows the vector kernels hand to csr_midrow_kernel (cnt is free again: reuse it as the output list)
        if (M) {
            if ((rc = cudaDeviceSynchronize() != cudaSuccess)) break;
            if ((rc = cub::DeviceSelect::If(tmp, b_sel, rows_it, cnt, d_num, (int) M, pred) != cudaSuccess)) break;
            if ((rc = cudaMemcpy(&m->nmid, d_num, 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        }
        if ((rc = cudaMalloc(&m->mid_rows, std::max<size_t>(1, m->nmid) * 4) != cudaSuccess)) break;
        if (m->nmid && (rc = cudaMemcpy(m->mid_rows, cnt, (size_t) m->nmid * 4, cudaMemcpyDeviceToDevice) != cudaSuccess)) break;
        // contiguous nnz-balanced row spans for the persistent vector kernel: SPANS_PER_SM big CTAs per SM
        int dev = 0, sms = 148, per_sm = 2;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (const char* e = getenv("SPMVB200_SPANS_PER_SM")) per_sm = std::max(1, atoi(e));
        m->nspans = (uint32_t) std::min<uint64_t>((uint64_t) sms * per_sm, std::max<uint64_t>(1, m->M));
        if ((rc = cudaMalloc(&m->span_b, ((size_t) m->nspans + 1) * 4) != cudaSuccess)) break;
        plan_spans_kernel<<<(m->nspans + 1 + 255) / 256, 256>>>(m->irp, M, m->NZ, m->nspans, m->span_b);
        // host copy of the tile bounds (chunking of the pipelined host path): two flat arrays
        const uint32_t nt1 = m->ntiles + 1;
        if ((rc = cudaMalloc(&h_r0, (size_t) nt1 * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&h_n0, (size_t) nt1 * 4) != cudaSuccess)) break;
        plan_split_desc_kernel<<<(nt1 + 255) / 256, 256>>>(m->desc, nt1, h_r0, h_n0);
        m->h_tile_row0.resize(nt1);
        m->h_tile_nnz0.resize(nt1);
        if ((rc = cudaMemcpy(m->h_tile_row0.data(), h_r0, (size_t) nt1 * 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        if ((rc = cudaMemcpy(m->h_tile_nnz0.data(), h_n0, (size_t) nt1 * 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
    } while (0);
    cudaFree(cnt);
    cudaFree(idx);
    cudaFree(d_num);
    cudaFree(tmp);
    cudaFree(h_r0);
    cudaFree(h_n0);
    if (rc) {
        if (cudaPeekAtLastError() != cudaSuccess || !g_err[0]) fail("CSR plan: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    // sub-warp width of the vector kernel from the mean row length (2 non-zeros per lane and step)
    const double mean = m->M ? (double) m->NZ / (double) m->M : 0.0;
    int lanes = 2;
    while (lanes < 32 && lanes * 2 < mean) lanes *= 2;
    if (const char* e = getenv("SPMVB200_VEC_LANES")) lanes = atoi(e);  // developer knob
    m->vec_lanes = lanes;
    return 0;
}

// Structure check at upload time: every kernel gathers x[col] unchecked and walks [irp[r], irp[r+1]) unchecked, so a malformed
// input (column id >= N from a wrong header, a row pointer that decreases) must become an error return here, not an out-of-bounds
// device read later.  flags[0]: row pointer not monotone, flags[1]: column id >= N.  One streaming pass over the ids.
__global__ void csr_validate_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, uint64_t M, uint64_t NZ, uint64_t N,
                                    int* __restrict__ flags) {
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x, t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    bool bad_r = false, bad_c = false;
    for (uint64_t r = t; r < M; r += stride) bad_r |= irp[r] > irp[r + 1];
    for (uint64_t j = t; j < NZ; j += stride) bad_c |= (uint64_t) ja[j] >= N;
    if (bad_r) flags[0] = 1;
    if (bad_c) flags[1] = 1;
}
__global__ void ell_validate_kernel(const uint32_t* __restrict__ ja, const uint32_t* __restrict__ rl, uint64_t pitch, uint32_t M, uint32_t K,
                                    int colmajor, uint64_t N, int* __restrict__ flags) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const uint32_t len = min(rl[r], K);
    bool bad = false;
    for (uint32_t k = 0; k < len; ++k) bad |= (uint64_t) ja[colmajor ? (uint64_t) k * pitch + r : (uint64_t) r * pitch + k] >= N;
    if (bad) flags[1] = 1;
}
static int validate_csr(const spmvb200_matrix* m, const char* who) {
    int* d = nullptr;
    CU_TRY(cudaMalloc(&d, 8));
    CU_TRY(cudaMemset(d, 0, 8));
    csr_validate_kernel<<<1184, 256>>>(m->irp, m->ja, m->M, m->NZ, m->N, d);
    int h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail("%s: structure check: %s", who, cudaGetErrorString(e));
    if (h[0]) return fail("%s: the row pointer is not monotone (IRP[r] > IRP[r+1] for some row)", who);
    if (h[1]) return fail("%s: a column id is >= N=%llu", who, (unsigned long long) m->N);
    return 0;
}
static int validate_ell(const spmvb200_matrix* m, const char* who) {
    if (!m->M || !m->K) return 0;
    int* d = nullptr;
    CU_TRY(cudaMalloc(&d, 8));
    CU_TRY(cudaMemset(d, 0, 8));
    ell_validate_kernel<<<(unsigned) ((m->M + 255) / 256), 256>>>(m->ja, m->rl, m->pitch, (uint32_t) m->M, (uint32_t) m->K,
                                                                     m->format == SPMVB200_FMT_ELL_COLMAJOR, m->N, d);
    int h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail("%s: structure check: %s", who, cudaGetErrorString(e));
    if (h[1]) return fail("%s: a column id is >= N=%llu", who, (unsigned long long) m->N);
    return 0;
}

static int check_dims(uint64_t M, uint64_t N, uint64_t NZ) {
    if (M >= 0x7fffffffull || N > 0xffffffffull || NZ >= 0xfffffff0ull)
        return fail("matrix too large for 32-bit device indices: M=%llu N=%llu NZ=%llu", (unsigned long long) M,
                    (unsigned long long) N, (unsigned long long) NZ);
    return 0;
}

// chunked H2D + narrowing of a 64-bit index array (bounded staging buffer)
static int upload_narrow(const uint64_t* h_src, uint64_t n, uint64_t sub, uint32_t* d_dst, int* d_overflow) {
    const uint64_t CH = 1ull << 25;  // 32 Mi elements = 256 MB staging
    uint64_t* stage = nullptr;
    if (!n) return 0;
    CU_TRY(cudaMalloc(&stage, std::min(n, CH) * 8));
    for (uint64_t o = 0; o < n; o += CH) {
        const uint64_t c = std::min(CH, n - o);
        cudaError_t e = cudaMemcpy(stage, h_src + o, c * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(stage);
            return fail("H2D of index chunk failed: %s", cudaGetErrorString(e));
        }
        narrow_u64_kernel<<<1184, 256>>>(stage, d_dst + o, c, sub, d_overflow);
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(stage);
    if (e != cudaSuccess) return fail("index narrowing failed: %s", cudaGetErrorString(e));
    return 0;
}

static int read_overflow(int* d_overflow, const char* what) {
    int h = 0;
    CU_TRY(cudaMemcpy(&h, d_overflow, sizeof(int), cudaMemcpyDeviceToHost));
    if (h) return fail("%s: an index does not fit 32 bits", what);
    return 0;
}

extern "C" int spmvb200_csr_upload(uint64_t M, uint64_t N, const uint64_t* irp, const uint64_t* ja, const double* as,
                                   uint64_t row_begin, uint64_t row_end, spmvb200_matrix** out) {
    if (!out) return fail("csr_upload: null output");
    *out = nullptr;
    if (!irp || (M && irp[M] && (!ja || !as)) || row_begin > row_end || row_end > M) return fail("csr_upload: bad arguments (null IRP, or null JA / AS with NZ > 0)");
    int ndev = 0;
    if (spmvb200_device_count(&ndev) || ndev == 0) return fail("csr_upload: no CUDA device (no CPU fallback)");
    const uint64_t rows = row_end - row_begin, n0 = irp[row_begin], nz = irp[row_end] - n0;
    if (check_dims(rows, N, nz)) return 1;
    spmvb200_matrix* m = new spmvb200_matrix();
    m->format = SPMVB200_FMT_CSR;
    m->M = rows;
    m->N = N;
    m->NZ = nz;
    m->own = 1;
    int* d_of = nullptr;
    int rc = 0;
    do {
        if ((rc = (cudaMalloc(&d_of, sizeof(int)) != cudaSuccess))) break;
        cudaMemset(d_of, 0, sizeof(int));
        if ((rc = (cudaMalloc(&m->irp, (rows + 1) * 4) != cudaSuccess))) break;
        if ((rc = (cudaMalloc(&m->ja, (nz + PAD) * 4) != cudaSuccess))) break;
        if ((rc = (cudaMalloc(&m->as, (nz + PAD) * 8) != cudaSuccess))) break;
        cudaMemset(m->ja + nz, 0, PAD * 4);
        cudaMemset(m->as + nz, 0, PAD * 8);
        if ((rc = upload_narrow(irp + row_begin, rows + 1, n0, m->irp, d_of))) break;
        if ((rc = upload_narrow(ja + n0, nz, 0, m->ja, d_of))) break;
        if (nz && (rc = (cudaMemcpy(m->as, as + n0, nz * 8, cudaMemcpyHostToDevice) != cudaSuccess))) break;
        if ((rc = read_overflow(d_of, "csr_upload"))) break;
        if ((rc = validate_csr(m, "csr_upload"))) break;
        rc = finish_csr(m);
    } while (0);
    cudaFree(d_of);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("csr_upload: %s", cudaGetErrorString(cudaGetLastError()));
        free_arrays(m);
        delete m;
        return 1;
    }
    *out = m;
    return 0;
}

extern "C" int spmvb200_csr_adopt_device(uint64_t M, uint64_t N, uint64_t NZ, uint32_t* d_irp32, uint32_t* d_ja32,
                                         double* d_as, int own, spmvb200_matrix** out) {
    if (!out) return fail("csr_adopt_device: null output");
    *out = nullptr;
    if (check_dims(M, N, NZ)) return 1;
    if (!d_irp32 || (NZ && (!d_ja32 || !d_as))) return fail("csr_adopt_device: null array");
    spmvb200_matrix* m = new spmvb200_matrix();
    m->format = SPMVB200_FMT_CSR;
    m->M = M;
    m->N = N;
    m->NZ = NZ;
    m->irp = d_irp32;
    m->ja = d_ja32;
    m->as = d_as;
    m->own = own;
    if (validate_csr(m, "csr_adopt_device") || finish_csr(m)) {
        m->own = 0;  // the caller keeps ownership on failure
        free_arrays(m);
        delete m;
        return 1;
    }
    *out = m;
    return 0;
}

// ------------------------------------------------------------------------------------------------- ELL
static uint64_t ell_pitch(int format, uint64_t rows, uint64_t K) {
    return format == SPMVB200_FMT_ELL_COLMAJOR ? ((rows + 63) / 64) * 64 : ((std::max<uint64_t>(K, 1) + 3) / 4) * 4;
}
static int ell_alloc(spmvb200_matrix* m) {
    const uint64_t slots = m->format == SPMVB200_FMT_ELL_COLMAJOR ? m->pitch * std::max<uint64_t>(m->K, 1) : m->pitch * std::max<uint64_t>(m->M, 1);
    m->slots = slots;
    CU_TRY(cudaMalloc(&m->ja, (slots + PAD) * 4));
    CU_TRY(cudaMalloc(&m->as, (slots + PAD) * 8));
    CU_TRY(cudaMemset(m->ja, 0, (slots + PAD) * 4));
    CU_TRY(cudaMemset(m->as, 0, (slots + PAD) * 8));
    CU_TRY(cudaMalloc(&m->rl, std::max<uint64_t>(m->M, 1) * 4));
    return 0;
}
// column-major ELL: if every valid column id is within a 2^16 range of its row index, keep 16-bit offsets as well
static int ell_try_idx16(spmvb200_matrix* m) {
    if (m->format != SPMVB200_FMT_ELL_COLMAJOR || !m->M || !m->K || getenv("SPMVB200_ELL_NO_IDX16")) return 0;
    long long* d_rng = nullptr;
    CU_TRY(cudaMalloc(&d_rng, 16));
    const long long init[2] = {(1ll << 62), -(1ll << 62)};
    CU_TRY(cudaMemcpy(d_rng, init, 16, cudaMemcpyHostToDevice));
    ell_delta_range_kernel<<<(unsigned) ((m->M + 255) / 256), 256>>>(m->ja, m->rl, m->pitch, (uint32_t) m->M, d_rng, d_rng + 1);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, d_rng, 16, cudaMemcpyDeviceToHost);
    cudaFree(d_rng);
    if (e != cudaSuccess) return fail("ELL index range: %s", cudaGetErrorString(e));
    if (h[0] > h[1] || h[1] - h[0] > 65535 || h[0] < -(1ll << 30) || h[0] > (1ll << 30)) return 0;  // empty, or too wide
    m->ja16_base = (int32_t) h[0];
    CU_TRY(cudaMalloc(&m->ja16, (m->slots + PAD) * 2));
    CU_TRY(cudaMemset(m->ja16 + m->slots, 0, PAD * 2));
    ell_make_idx16_kernel<<<(unsigned) ((m->M + 255) / 256), 256>>>(m->ja, m->rl, m->pitch, (uint32_t) m->M, (uint32_t) m->K, m->ja16_base, m->ja16);
    CU_TRY(cudaDeviceSynchronize());
    return 0;
}

static void ell_pick_lanes(spmvb200_matrix* m) {
    int lanes = 1;
    while (lanes < 32 && (uint64_t) lanes * 4 <= m->K) lanes *= 2;  // ~2-4 slots per lane
    m->vec_lanes = lanes;
}

extern "C" int spmvb200_ell_upload(uint64_t M, uint64_t N, uint64_t K, const uint64_t* ja, const double* as,
                                   const uint64_t* rl, uint64_t row_begin, uint64_t row_end, int format,
                                   spmvb200_matrix** out) {
    if (!out) return fail("ell_upload: null output");
    *out = nullptr;
    if (format != SPMVB200_FMT_ELL_COLMAJOR && format != SPMVB200_FMT_ELL_ROWMAJOR) return fail("ell_upload: bad format %d", format);
    if (row_begin > row_end || row_end > M || (M && K && (!ja || !as))) return fail("ell_upload: bad arguments");
    int ndev = 0;
    if (spmvb200_device_count(&ndev) || ndev == 0) return fail("ell_upload: no CUDA device (no CPU fallback)");
    const uint64_t rows = row_end - row_begin;
    if (check_dims(rows, N, rows * K) || K > 0xffffffffull) return 1;
    spmvb200_matrix* m = new spmvb200_matrix();
    m->format = format;
    m->M = rows;
    m->N = N;
    m->K = K;
    m->own = 1;
    m->pitch = ell_pitch(format, rows, K);
    int* d_of = nullptr;
    uint64_t *st_ja = nullptr, *st_rl = nullptr;
    double* st_as = nullptr;
    int rc = 0;
    do {
        if ((rc = ell_alloc(m))) break;
        if ((rc = (cudaMalloc(&d_of, sizeof(int)) != cudaSuccess))) break;
        cudaMemset(d_of, 0, sizeof(int));
        if (rl) {
            if ((rc = upload_narrow(rl + row_begin, rows, 0, m->rl, d_of))) break;
        }
        // row chunks of the row-major host arrays -> staging -> layout kernel
        const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(rows, (1ull << 25) / std::max<uint64_t>(K, 1)));
        if (rows && K) {
            if ((rc = (cudaMalloc(&st_ja, chunk_rows * K * 8) != cudaSuccess))) break;
            if ((rc = (cudaMalloc(&st_as, chunk_rows * K * 8) != cudaSuccess))) break;
            for (uint64_t r = 0; r < rows && !rc; r += chunk_rows) {
                const uint64_t cr = std::min(chunk_rows, rows - r);
                const uint64_t ho = (row_begin + r) * K;
                if ((rc = (cudaMemcpy(st_ja, ja + ho, cr * K * 8, cudaMemcpyHostToDevice) != cudaSuccess))) break;
                if ((rc = (cudaMemcpy(st_as, as + ho, cr * K * 8, cudaMemcpyHostToDevice) != cudaSuccess))) break;
                if (format == SPMVB200_FMT_ELL_COLMAJOR) {
                    dim3 grid((unsigned) ((cr + 31) / 32), (unsigned) ((K + 31) / 32));
                    ell_transpose_kernel<<<grid, dim3(32, 8)>>>(st_ja, st_as, (uint32_t) cr, (uint32_t) K, (uint32_t) r, m->pitch, m->ja, m->as, d_of);
                } else {
                    ell_repitch_kernel<<<1184, 256>>>(st_ja, st_as, (uint32_t) cr, (uint32_t) K, (uint32_t) r, m->pitch, m->ja, m->as, d_of);
                }
                if (!rl) ell_derive_rl_kernel<<<(unsigned) ((cr + 255) / 256), 256>>>(st_as, (uint32_t) cr, (uint32_t) K, (uint32_t) r, m->rl);
                rc = cudaDeviceSynchronize() != cudaSuccess;
            }
            if (rc) break;
        } else if (rows) {
            cudaMemset(m->rl, 0, rows * 4);
        }
        if ((rc = read_overflow(d_of, "ell_upload"))) break;
        // NZ = sum of row lengths (host side sum of the narrowed vector; upload time only)
        std::vector<uint32_t> h_rl(rows);
        if (rows && (rc = (cudaMemcpy(h_rl.data(), m->rl, rows * 4, cudaMemcpyDeviceToHost) != cudaSuccess))) break;
        uint64_t nz = 0;
        for (uint64_t r = 0; r < rows; ++r) {
            if (h_rl[r] > K) { rc = fail("ell_upload: row length %u > K=%llu at row %llu", h_rl[r], (unsigned long long) K, (unsigned long long) r); break; }
            nz += h_rl[r];
        }
        if (rc) break;
        m->NZ = nz;
        if ((rc = validate_ell(m, "ell_upload"))) break;
        ell_pick_lanes(m);
        if ((rc = ell_try_idx16(m))) break;
    } while (0);
    cudaFree(d_of);
    cudaFree(st_ja);
    cudaFree(st_as);
    cudaFree(st_rl);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("ell_upload: %s", cudaGetErrorString(cudaGetLastError()));
        free_arrays(m);
        delete m;
        return 1;
    }
    *out = m;
    return 0;
}

extern "C" int spmvb200_ell_from_csr(const spmvb200_matrix* csr, int format, spmvb200_matrix** out) {
    if (!out) return fail("ell_from_csr: null output");
    *out = nullptr;
    if (!csr || csr->format != SPMVB200_FMT_CSR) return fail("ell_from_csr: source is not a CSR handle");
    if (format != SPMVB200_FMT_ELL_COLMAJOR && format != SPMVB200_FMT_ELL_ROWMAJOR) return fail("ell_from_csr: bad format %d", format);
    spmvb200_matrix* m = new spmvb200_matrix();
    m->format = format;
    m->M = csr->M;
    m->N = csr->N;
    m->NZ = csr->NZ;
    m->own = 1;
    uint32_t* d_kmax = nullptr;
    int rc = 0;
    do {
        if ((rc = (cudaMalloc(&d_kmax, 4) != cudaSuccess))) break;
        cudaMemset(d_kmax, 0, 4);
        if ((rc = (cudaMalloc(&m->rl, std::max<uint64_t>(m->M, 1) * 4) != cudaSuccess))) break;
        if (m->M) row_len_from_irp_kernel<<<(unsigned) ((m->M + 255) / 256), 256>>>(csr->irp, (uint32_t) m->M, m->rl, d_kmax);
        uint32_t kmax = 0;
        if ((rc = (cudaMemcpy(&kmax, d_kmax, 4, cudaMemcpyDeviceToHost) != cudaSuccess))) break;
        m->K = kmax;
        m->pitch = ell_pitch(format, m->M, m->K);
        uint32_t* keep_rl = m->rl;
        m->rl = nullptr;
        rc = ell_alloc(m);  // allocates a fresh rl too
        if (rc) { cudaFree(keep_rl); break; }
        cudaFree(m->rl);
        m->rl = keep_rl;
        if (m->M && m->K)
            csr_to_ell_kernel<<<(unsigned) ((m->M + 255) / 256), 256>>>(csr->irp, csr->ja, csr->as, (uint32_t) m->M, (uint32_t) m->K, m->pitch,
                                                                         format == SPMVB200_FMT_ELL_COLMAJOR, m->ja, m->as);
        if ((rc = (cudaDeviceSynchronize() != cudaSuccess))) break;
        ell_pick_lanes(m);
        if ((rc = ell_try_idx16(m))) break;
    } while (0);
    cudaFree(d_kmax);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("ell_from_csr: %s", cudaGetErrorString(cudaGetLastError()));
        free_arrays(m);
        delete m;
        return 1;
    }
    *out = m;
    return 0;
}

// ------------------------------------------------------------------------------------------------- SELL-32-sigma
static int sell_build(const spmvb200_matrix* csr, uint32_t sigma, uint32_t cap, spmvb200_matrix** out);
// A thread-per-row format must not walk a 10^5-entry row with one thread (R-MAT: 37 ms for one SpMV).  A stand-alone SELL handle
// therefore leaves rows longer than VEC_MID out of its slices and keeps them as a compact CSR handle of its own (`tail`): the
// per-row / per-segment kernels run them next to the slices and a scatter writes their y entries.
struct LongRowPred {
    const uint32_t* irp;
    uint32_t thr;
    __device__ bool operator()(uint32_t r) const { return irp[r + 1] - irp[r] > thr; }
};
__global__ void tail_len_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ map, uint32_t n, uint32_t* __restrict__ len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) len[i] = irp[map[i] + 1] - irp[map[i]];
    if (i == n) len[i] = 0;
}
__global__ void tail_copy_kernel(const uint32_t* __restrict__ irp, const uint32_t* __restrict__ ja, const double* __restrict__ as,
                                 const uint32_t* __restrict__ map, const uint32_t* __restrict__ irp_t, uint32_t n, uint32_t* __restrict__ ja_t,
                                 double* __restrict__ as_t) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint32_t s = irp[map[w]], len = irp[map[w] + 1] - s, d = irp_t[w];
    for (uint32_t k = lane; k < len; k += 32) {
        ja_t[d + k] = ja[s + k];
        as_t[d + k] = as[s + k];
    }
}
__global__ void tail_scatter_kernel(const double* __restrict__ yt, const uint32_t* __restrict__ map, uint32_t n, double* __restrict__ y) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[map[i]] = yt[i];
}
static int sell_attach_tail(const spmvb200_matrix* csr, spmvb200_matrix* sell) {
    const uint32_t M = (uint32_t) csr->M;
    uint32_t *map = nullptr, *d_num = nullptr, *len = nullptr, *irp_t = nullptr, *ja_t = nullptr;
    double* as_t = nullptr;
    void* tmp = nullptr;
    int rc = 0;
    do {
        if ((rc = cudaMalloc(&map, (size_t) M * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&d_num, 4) != cudaSuccess)) break;
        thrust::counting_iterator<uint32_t> rows_it(0);
        LongRowPred pred{csr->irp, (uint32_t) VEC_MID};
        size_t b_sel = 0, b_scan = 0;
        cub::DeviceSelect::If(nullptr, b_sel, rows_it, map, d_num, (int) M, pred);
        cub::DeviceScan::ExclusiveSum(nullptr, b_scan, len, irp_t, (int) M + 1);
        if ((rc = cudaMalloc(&tmp, std::max(b_sel, b_scan) + 16) != cudaSuccess)) break;
        if ((rc = cub::DeviceSelect::If(tmp, b_sel, rows_it, map, d_num, (int) M, pred) != cudaSuccess)) break;
        uint32_t n = 0;
        if ((rc = cudaMemcpy(&n, d_num, 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        if (!n) { rc = fail("sell tail: no long rows found"); break; }
        if ((rc = cudaMalloc(&len, ((size_t) n + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&irp_t, ((size_t) n + 1) * 4) != cudaSuccess)) break;
        tail_len_kernel<<<(n + 1 + 255) / 256, 256>>>(csr->irp, map, n, len);
        if ((rc = cub::DeviceScan::ExclusiveSum(tmp, b_scan, len, irp_t, (int) n + 1) != cudaSuccess)) break;
        uint32_t nz = 0;
        if ((rc = cudaMemcpy(&nz, irp_t + n, 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&ja_t, ((size_t) nz + PAD) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&as_t, ((size_t) nz + PAD) * 8) != cudaSuccess)) break;
        cudaMemset(ja_t + nz, 0, PAD * 4);
        cudaMemset(as_t + nz, 0, PAD * 8);
        tail_copy_kernel<<<(unsigned) (((uint64_t) n * 32 + 255) / 256), 256>>>(csr->irp, csr->ja, csr->as, map, irp_t, n, ja_t, as_t);
        if ((rc = cudaDeviceSynchronize() != cudaSuccess)) break;
        if ((rc = spmvb200_csr_adopt_device(n, csr->N, nz, irp_t, ja_t, as_t, 1, &sell->tail))) break;
        irp_t = nullptr; ja_t = nullptr; as_t = nullptr;  // owned by the tail handle now
        if ((rc = cudaMalloc(&sell->tail_y, (size_t) n * 8) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&sell->tail_map, (size_t) n * 4) != cudaSuccess)) break;
        if ((rc = cudaMemcpy(sell->tail_map, map, (size_t) n * 4, cudaMemcpyDeviceToDevice) != cudaSuccess)) break;
    } while (0);
    cudaFree(map);
    cudaFree(d_num);
    cudaFree(len);
    cudaFree(irp_t);
    cudaFree(ja_t);
    cudaFree(as_t);
    cudaFree(tmp);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("sell tail: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}
extern "C" int spmvb200_sell_from_csr(const spmvb200_matrix* csr, uint32_t sigma, spmvb200_matrix** out) {
    const bool skewed = csr && csr->format == SPMVB200_FMT_CSR && csr->lmax > (uint32_t) VEC_MID;
    if (sell_build(csr, sigma, skewed ? (uint32_t) VEC_MID : 0xffffffffu, out)) return 1;
    if (skewed && sell_attach_tail(csr, *out)) {
        spmvb200_free(*out);
        *out = nullptr;
        return 1;
    }
    return 0;
}
// cap < 2^32-1: rows longer than cap are left empty (hybrid of the adaptive mode: they go to the per-row / per-segment CTAs)
static int sell_build(const spmvb200_matrix* csr, uint32_t sigma, uint32_t cap, spmvb200_matrix** out) {
    if (!out) return fail("sell_from_csr: null output");
    *out = nullptr;
    if (!csr || (csr->format != SPMVB200_FMT_CSR && csr->format != SPMVB200_FMT_ELL_COLMAJOR))
        return fail("sell_from_csr: source is neither a CSR nor a column-major ELL handle");
    const RowSrc src = {csr->irp, csr->rl, csr->ja, csr->as, csr->format == SPMVB200_FMT_CSR ? 0ull : csr->pitch};
    if (sigma == 0) sigma = 16384;
    if (sigma % 32) return fail("sell_from_csr: sigma must be a multiple of 32");
    spmvb200_matrix* m = new spmvb200_matrix();
    m->format = SPMVB200_FMT_SELL;
    m->M = csr->M;
    m->N = csr->N;
    m->NZ = csr->NZ;
    m->own = 1;
    m->Mpad = ((csr->M + 31) / 32) * 32;
    const uint32_t Mpad = (uint32_t) m->Mpad, nsl = Mpad / 32;
    uint64_t *k0 = nullptr, *k1 = nullptr, *slots = nullptr, *slots_scan = nullptr;
    uint32_t* v0 = nullptr;
    void* tmp = nullptr;
    int rc = 0;
    do {
        if (Mpad == 0) { rc = fail("sell_from_csr: empty matrix"); break; }
        if ((rc = cudaMalloc(&k0, (size_t) Mpad * 8) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&k1, (size_t) Mpad * 8) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&v0, (size_t) Mpad * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->perm, (size_t) Mpad * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->rl, (size_t) Mpad * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->irp, ((size_t) nsl + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&slots, ((size_t) nsl + 1) * 8) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&slots_scan, ((size_t) nsl + 1) * 8) != cudaSuccess)) break;
        sell_keys_kernel<<<(Mpad + 255) / 256, 256>>>(src, (uint32_t) csr->M, Mpad, sigma, cap, k0, v0);
        size_t b1 = 0, b2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, k0, k1, v0, m->perm, (int) Mpad);
        cub::DeviceScan::ExclusiveSum(nullptr, b2, slots, slots_scan, (int) nsl + 1);
        if ((rc = cudaMalloc(&tmp, std::max(b1, b2) + 16) != cudaSuccess)) break;
        if ((rc = cub::DeviceRadixSort::SortPairs(tmp, b1, k0, k1, v0, m->perm, (int) Mpad) != cudaSuccess)) break;
        cudaMemset(slots, 0, ((size_t) nsl + 1) * 8);
        sell_slices_kernel<<<(Mpad + 255) / 256, 256>>>(k1, Mpad, m->rl, slots);
        if ((rc = cub::DeviceScan::ExclusiveSum(tmp, b2, slots, slots_scan, (int) nsl + 1) != cudaSuccess)) break;
        uint64_t total = 0;
        if ((rc = cudaMemcpy(&total, slots_scan + nsl, 8, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        if (total >= 0xfffffff0ull) { rc = fail("sell_from_csr: %llu slots do not fit 32-bit offsets", (unsigned long long) total); break; }
        m->slots = total;
        narrow_u64_kernel<<<592, 256>>>(slots_scan, m->irp, (uint64_t) nsl + 1, 0, (int*) slots);  // slots[] reused as overflow flag sink
        if ((rc = cudaMalloc(&m->ja, (total + PAD) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->as, (total + PAD) * 8) != cudaSuccess)) break;
        cudaMemset(m->ja + total, 0, PAD * 4);
        cudaMemset(m->as + total, 0, PAD * 8);
        sell_fill_kernel<<<(Mpad + 255) / 256, 256>>>(src, m->perm, m->irp, Mpad, cap, m->ja, m->as);
        if ((rc = cudaDeviceSynchronize() != cudaSuccess)) break;
        m->K = sigma;  // reported through spmvb200_dims as K
    } while (0);
    cudaFree(k0);
    cudaFree(k1);
    cudaFree(v0);
    cudaFree(slots);
    cudaFree(slots_scan);
    cudaFree(tmp);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("sell_from_csr: %s", cudaGetErrorString(cudaGetLastError()));
        free_arrays(m);
        delete m;
        return 1;
    }
    *out = m;
    return 0;
}


// ------------------------------------------------------------------------------------------------- hot-x hybrid (hotx.cuh)
// Column histogram -> H hottest columns -> remapped ids -> SELL copy of the short rows.  Fails (quietly, under g_quiet) when the hot
// columns cover less than min_cover of the non-zeros: uniform column distributions gain nothing from the cache.
static void hotx_drop(spmvb200_matrix* m);
static int hotx_build(spmvb200_matrix* m, uint32_t H, double min_cover) {
    if (m->format != SPMVB200_FMT_CSR) return fail("hotx: source is not a CSR handle");
    if (m->hot_sell && m->hot_H == H) return 0;
    if (m->hot_sell) hotx_drop(m);
    if (m->N < 4ull * H || m->NZ < (1u << 20) || m->N + H > 0xffffffffull) return fail("hotx: matrix too small (or too wide) for a hot-column cache");
    const uint32_t N = (uint32_t) m->N;
    uint32_t *cnt = nullptr, *cnt_s = nullptr, *cols = nullptr, *cols_s = nullptr, *remap = nullptr;
    void* tmp = nullptr;
    int rc = 0;
    do {
        if ((rc = cudaMalloc(&cnt, (size_t) N * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&cnt_s, (size_t) N * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&cols, (size_t) N * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&cols_s, (size_t) N * 4) != cudaSuccess)) break;
        cudaMemset(cnt, 0, (size_t) N * 4);
        hotx_hist_kernel<<<1184, 256>>>(m->ja, m->NZ, cnt);
        hotx_iota_kernel<<<(N + 255) / 256, 256>>>(cols, N);
        size_t b1 = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, b1, cnt, cnt_s, cols, cols_s, (int) N);
        if ((rc = cudaMalloc(&tmp, b1 + 16) != cudaSuccess)) break;
        if ((rc = cub::DeviceRadixSort::SortPairsDescending(tmp, b1, cnt, cnt_s, cols, cols_s, (int) N) != cudaSuccess)) break;
        std::vector<uint32_t> h(H);
        if ((rc = cudaMemcpy(h.data(), cnt_s, (size_t) H * 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        uint64_t covered = 0;
        for (uint32_t i = 0; i < H; ++i) covered += h[i];
        m->hot_cover = (float) ((double) covered / (double) std::max<uint64_t>(m->NZ, 1));
        if ((double) m->hot_cover < min_cover) { rc = fail("hotx: the %u hottest columns cover only %.1f %% of the non-zeros", H, 100.0 * m->hot_cover); break; }
        if ((rc = cudaMalloc(&m->hot_cols, (size_t) H * 4) != cudaSuccess)) break;
        if ((rc = cudaMemcpy(m->hot_cols, cols_s, (size_t) H * 4, cudaMemcpyDeviceToDevice) != cudaSuccess)) break;
        remap = cnt;  // the histogram is not needed any more
        hotx_remap_init_kernel<<<(N + 255) / 256, 256>>>(remap, N, H);
        hotx_remap_hot_kernel<<<(H + 255) / 256, 256>>>(remap, m->hot_cols, H);
        if ((rc = cudaMalloc(&m->ja_hot, (m->NZ + PAD) * 4) != cudaSuccess)) break;
        cudaMemset(m->ja_hot + m->NZ, 0, PAD * 4);
        hotx_apply_kernel<<<1184, 256>>>(m->ja, remap, m->NZ, m->ja_hot);
        if ((rc = cudaDeviceSynchronize() != cudaSuccess)) break;
        // SELL copy of the rows of at most VEC_MID entries, from the remapped ids (a shallow view of the parent with ja swapped)
        spmvb200_matrix view;
        view.format = SPMVB200_FMT_CSR;
        view.M = m->M;
        view.N = m->N;
        view.NZ = m->NZ;
        view.irp = m->irp;
        view.ja = m->ja_hot;
        view.as = m->as;
        if ((rc = sell_build(&view, 0, m->lmax <= (uint32_t) VEC_MID ? 0xffffffffu : (uint32_t) VEC_MID, &m->hot_sell))) break;
        {   // slices by decreasing length
            const uint32_t nsl = (uint32_t) (m->hot_sell->Mpad / 32);
            uint32_t *k0 = nullptr, *k1 = nullptr, *v0 = nullptr;
            void* t2 = nullptr;
            size_t b2 = 0;
            if ((rc = cudaMalloc(&k0, (size_t) nsl * 4) != cudaSuccess)) break;
            if (!rc) rc = cudaMalloc(&k1, (size_t) nsl * 4) != cudaSuccess;
            if (!rc) rc = cudaMalloc(&v0, (size_t) nsl * 4) != cudaSuccess;
            if (!rc) rc = cudaMalloc(&m->hot_slice_order, (size_t) nsl * 4) != cudaSuccess;
            if (!rc) {
                hotx_slice_keys_kernel<<<(nsl + 255) / 256, 256>>>(m->hot_sell->irp, nsl, k0, v0);
                cub::DeviceRadixSort::SortPairs(nullptr, b2, k0, k1, v0, m->hot_slice_order, (int) nsl);
                rc = cudaMalloc(&t2, b2 + 16) != cudaSuccess;
                if (!rc) rc = cub::DeviceRadixSort::SortPairs(t2, b2, k0, k1, v0, m->hot_slice_order, (int) nsl) != cudaSuccess;
                if (!rc) rc = cudaDeviceSynchronize() != cudaSuccess;
            }
            cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(t2);
            if (rc) break;
        }
        m->hot_H = H;
    } while (0);
    cudaFree(cnt);
    cudaFree(cnt_s);
    cudaFree(cols);
    cudaFree(cols_s);
    cudaFree(tmp);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("hotx: %s", cudaGetErrorString(cudaGetLastError()));
        if (m->hot_sell) { spmvb200_free(m->hot_sell); m->hot_sell = nullptr; }
        cudaFree(m->hot_cols);
        cudaFree(m->ja_hot);
        cudaFree(m->hot_slice_order);
        m->hot_cols = nullptr;
        m->ja_hot = nullptr;
        m->hot_slice_order = nullptr;
        m->hot_H = 0;
        return 1;
    }
    return 0;
}
static void hotx_drop(spmvb200_matrix* m) {
    if (m->hot_sell) { spmvb200_free(m->hot_sell); m->hot_sell = nullptr; }
    cudaFree(m->hot_cols);
    cudaFree(m->ja_hot);
    cudaFree(m->hot_slice_order);
    m->hot_cols = nullptr;
    m->ja_hot = nullptr;
    m->hot_slice_order = nullptr;
    m->hot_H = 0;
}

// ------------------------------------------------------------------------------------------------- x-window CSR
// (xwin.cuh) built on the device from a CSR handle: mark (row block, window) pairs -> scan -> tile list ->
// per-row counts -> scan -> jagged slot-major fill.
// max_window_ratio > 0: give up right after the tile census (before anything big is allocated) when the x windows one SpMV
// would stage exceed that many bytes per non-zero -- the adaptive mode's way of asking "is this matrix local enough?"
static int xwin_build(const spmvb200_matrix* csr, uint32_t rows_per_block, uint32_t window_cols, double max_window_ratio,
                      spmvb200_matrix** out) {
    if (!out) return fail("xwin_from_csr: null output");
    *out = nullptr;
    if (!csr || csr->format != SPMVB200_FMT_CSR) return fail("xwin_from_csr: source is not a CSR handle");
    uint32_t R = rows_per_block ? rows_per_block : 2048, W = window_cols ? window_cols : 8192;
    if (const char* e = getenv("SPMVB200_XW_R")) R = (uint32_t) atoi(e);  // developer knobs
    if (const char* e = getenv("SPMVB200_XW_W")) W = (uint32_t) atoi(e);
    if (R < 512 || R > 4096 || (R & (R - 1))) return fail("xwin_from_csr: rows_per_block must be a power of two in [512, 4096] (got %u)", R);
    if (W < 64 || W > 65536 || (W & 1)) return fail("xwin_from_csr: window_cols must be even and in [64, 65536] (got %u)", W);
    if (csr->M == 0) return fail("xwin_from_csr: empty matrix");
    const uint32_t M = (uint32_t) csr->M, G = R / 32;
    const uint32_t nrb = (M + R - 1) / R;
    const uint64_t nwin = (std::max<uint64_t>(csr->N, 1) + W - 1) / W;
    const uint32_t nwords = (uint32_t) ((nwin + 31) / 32);
    const uint64_t nbits_words = (uint64_t) nrb * nwords;
    if (nbits_words > (1ull << 28)) return fail("xwin_from_csr: %u row blocks x %llu windows is too sparse a tiling for this format", nrb, (unsigned long long) nwin);
    spmvb200_matrix* m = new spmvb200_matrix();
    m->format = SPMVB200_FMT_XWIN;
    m->M = csr->M;
    m->N = csr->N;
    m->NZ = csr->NZ;
    m->own = 1;
    m->xw_R = R;
    m->xw_W = W;
    m->xw_nrb = nrb;
    m->K = W;
    uint32_t *bitmap = nullptr, *pc = nullptr, *scan = nullptr, *tile_rb = nullptr, *grp_cnt = nullptr;
    int* d_flags = nullptr;  // [0] unsorted rows seen, [1] more than 255 entries of one row in one window
    void* tmp = nullptr;
    int rc = 0;
    do {
        if ((rc = cudaMalloc(&bitmap, nbits_words * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&pc, (nbits_words + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&scan, (nbits_words + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&d_flags, 8) != cudaSuccess)) break;
        cudaMemset(bitmap, 0, nbits_words * 4);
        cudaMemset(d_flags, 0, 8);
        xw_mark_kernel<<<(M + 255) / 256, 256>>>(csr->irp, csr->ja, M, R, W, nwords, bitmap, d_flags);
        xw_popc_kernel<<<(unsigned) ((nbits_words + 1 + 255) / 256), 256>>>(bitmap, nbits_words, pc);
        size_t b1 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b1, pc, scan, (int) (nbits_words + 1));
        if ((rc = cudaMalloc(&tmp, b1 + 16) != cudaSuccess)) break;
        if ((rc = cub::DeviceScan::ExclusiveSum(tmp, b1, pc, scan, (int) (nbits_words + 1)) != cudaSuccess)) break;
        uint32_t ntiles = 0;
        int h_flags[2] = {0, 0};
        if ((rc = cudaMemcpy(&ntiles, scan + nbits_words, 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        if ((rc = cudaMemcpy(h_flags, d_flags, 8, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        m->xw_ntiles = ntiles;
        m->xw_sorted = !h_flags[0];
        if (max_window_ratio > 0 && (double) ntiles * W * 8 > max_window_ratio * (double) std::max<uint64_t>(csr->NZ, 1)) {
            rc = fail("xwin_from_csr: %u tiles of %u columns: %.1f bytes of x windows per non-zero, not local enough", ntiles, W,
                      (double) ntiles * W * 8 / (double) std::max<uint64_t>(csr->NZ, 1));
            break;
        }
        const uint64_t ngroups = (uint64_t) ntiles * G;
        if (ngroups >= 0x7fffffffull) { rc = fail("xwin_from_csr: %llu (tile, group) pairs: tiling too fine", (unsigned long long) ngroups); break; }
        if ((rc = cudaMalloc(&m->xw_rb_tile0, ((size_t) nrb + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->xw_tile_win, std::max<size_t>(1, ntiles) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&tile_rb, std::max<size_t>(1, ntiles) * 4) != cudaSuccess)) break;
        xw_tiles_kernel<<<(unsigned) ((nbits_words + 1 + 255) / 256), 256>>>(bitmap, scan, nrb, nwords, m->xw_rb_tile0, m->xw_tile_win, tile_rb);
        if ((rc = cudaMalloc(&m->xw_cnt, std::max<size_t>(1, (size_t) ntiles * R) * 2) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&grp_cnt, (ngroups + 1) * 4) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->xw_grp_off, (ngroups + 1) * 4) != cudaSuccess)) break;
        // warps per CTA: fixed before the counts are stored, the layout of the per-row words follows the kernel shape (xw_cp_index)
        m->xw_nw = R >= 1024 ? 32 : 16;
        if (const char* e = getenv("SPMVB200_XW_NW")) m->xw_nw = (uint32_t) atoi(e);
        if ((m->xw_nw != 16 && m->xw_nw != 32) || R / (32 * m->xw_nw) < 1 || R / (32 * m->xw_nw) > (m->xw_nw == 32 ? 4u : 8u)) { rc = fail("xwin_from_csr: no kernel for R=%u with %u warps", R, m->xw_nw); break; }
        const unsigned cblocks = (unsigned) (((ngroups + 1) * 32 + 255) / 256);
        if (m->xw_sorted)
            xw_count_kernel<true><<<cblocks, 256>>>(csr->irp, csr->ja, M, R, W, m->xw_tile_win, tile_rb, ngroups, m->xw_cnt, grp_cnt, d_flags + 1, m->xw_nw);
        else
            xw_count_kernel<false><<<cblocks, 256>>>(csr->irp, csr->ja, M, R, W, m->xw_tile_win, tile_rb, ngroups, m->xw_cnt, grp_cnt, d_flags + 1, m->xw_nw);
        size_t b2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b2, grp_cnt, m->xw_grp_off, (int) (ngroups + 1));
        if (b2 > b1) {
            cudaFree(tmp);
            tmp = nullptr;
            if ((rc = cudaMalloc(&tmp, b2 + 16) != cudaSuccess)) break;
        }
        if ((rc = cub::DeviceScan::ExclusiveSum(tmp, b2, grp_cnt, m->xw_grp_off, (int) (ngroups + 1)) != cudaSuccess)) break;
        uint32_t total = 0;
        if ((rc = cudaMemcpy(&total, m->xw_grp_off + ngroups, 4, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        if ((rc = cudaMemcpy(h_flags, d_flags, 8, cudaMemcpyDeviceToHost) != cudaSuccess)) break;
        if (h_flags[1]) { rc = fail("xwin_from_csr: a row has more than 255 non-zeros inside one %u-column window (format limit)", W); break; }
        if (total != csr->NZ) { rc = fail("xwin_from_csr: internal count mismatch (%u entries placed, NZ=%llu)", total, (unsigned long long) csr->NZ); break; }
        if ((rc = cudaMalloc(&m->xw_col, (m->NZ + PAD) * 2) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&m->as, (m->NZ + PAD) * 8) != cudaSuccess)) break;
        cudaMemset(m->xw_col + m->NZ, 0, PAD * 2);
        cudaMemset(m->as + m->NZ, 0, PAD * 8);
        if (ngroups) {
            const unsigned fblocks = (unsigned) ((ngroups * 32 + 255) / 256);
            if (m->xw_sorted)
                xw_fill_kernel<true><<<fblocks, 256>>>(csr->irp, csr->ja, csr->as, M, R, W, m->xw_tile_win, tile_rb, ngroups, m->xw_cnt, m->xw_grp_off, m->xw_col, m->as, m->xw_nw);
            else
                xw_fill_kernel<false><<<fblocks, 256>>>(csr->irp, csr->ja, csr->as, M, R, W, m->xw_tile_win, tile_rb, ngroups, m->xw_cnt, m->xw_grp_off, m->xw_col, m->as, m->xw_nw);
        }
        {   // persistent CTAs: one per SM (or per row block if there are fewer), contiguous row blocks balanced by non-zeros
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            m->xw_ncta = std::min<uint32_t>(nrb, (uint32_t) sms);
            if (const char* e = getenv("SPMVB200_XW_NCTA")) {  // developer knob: force persistent CTAs of this count
                m->xw_ncta = std::min<uint32_t>(nrb, (uint32_t) std::max(1, atoi(e)));
                m->xw_mode = 1;
            }
            if (const char* e = getenv("SPMVB200_XW_MODE")) m->xw_mode = atoi(e);
            if ((rc = cudaMalloc(&m->xw_cta_rb, ((size_t) m->xw_ncta + 1) * 4) != cudaSuccess)) break;
            xw_cta_split_kernel<<<(m->xw_ncta + 1 + 255) / 256, 256>>>(m->xw_rb_tile0, m->xw_grp_off, G, nrb, m->NZ, m->xw_ncta, m->xw_cta_rb);
        }
        if ((rc = cudaDeviceSynchronize() != cudaSuccess)) break;
        // ring depth: as many windows as fit next to the barriers in the 227 KB a CTA may use
        uint32_t nbuf = (uint32_t) std::min<uint64_t>(XW_MAX_NBUF, (232448 - 256) / (((uint64_t) W + 2) * 8));
        if (const char* e = getenv("SPMVB200_XW_NBUF")) nbuf = std::min<uint32_t>(nbuf, (uint32_t) std::max(1, atoi(e)));
        if (nbuf < 2) { rc = fail("xwin_from_csr: window of %u columns leaves no room for double buffering", W); break; }
        uint32_t nbuf_cap = 4;
        if (const char* e = getenv("SPMVB200_XW_NBUF_CAP")) nbuf_cap = (uint32_t) std::min(std::max(atoi(e), 2), (int) XW_MAX_NBUF);  // developer knob
        m->xw_nbuf = std::min<uint32_t>(nbuf, nbuf_cap);
    } while (0);
    cudaFree(bitmap);
    cudaFree(pc);
    cudaFree(scan);
    cudaFree(tile_rb);
    cudaFree(grp_cnt);
    cudaFree(d_flags);
    cudaFree(tmp);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("xwin_from_csr: %s", cudaGetErrorString(cudaGetLastError()));
        free_arrays(m);
        delete m;
        return 1;
    }
    *out = m;
    return 0;
}

extern "C" int spmvb200_xwin_from_csr(const spmvb200_matrix* csr, uint32_t rows_per_block, uint32_t window_cols,
                                      spmvb200_matrix** out) {
    return xwin_build(csr, rows_per_block, window_cols, 0.0, out);
}

extern "C" int spmvb200_xwin_info(const spmvb200_matrix* m, uint32_t* rows_per_block, uint32_t* window_cols, uint32_t* ntiles,
                                  uint32_t* ring, uint64_t* moved_bytes) {
    if (!m || m->format != SPMVB200_FMT_XWIN) return fail("xwin_info: not an x-window handle");
    if (rows_per_block) *rows_per_block = m->xw_R;
    if (window_cols) *window_cols = m->xw_W;
    if (ntiles) *ntiles = m->xw_ntiles;
    if (ring) *ring = m->xw_nbuf;
    // what one SpMV reads and writes: entries, per-row counts, group offsets, tile list, y -- and the x windows (from L2)
    if (moved_bytes)
        *moved_bytes = 10 * m->NZ + (uint64_t) m->xw_ntiles * m->xw_R * 2 + (uint64_t) m->xw_ntiles * (m->xw_R / 32) * 4 + (uint64_t) m->xw_ntiles * 4 +
                       8 * m->M + (uint64_t) m->xw_ntiles * m->xw_W * 8;
    return 0;
}
