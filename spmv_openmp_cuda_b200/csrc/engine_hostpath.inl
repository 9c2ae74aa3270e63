// engine_hostpath.inl -- part of engine.cu (textually included there: one translation unit, file-local helpers stay static).
// host-buffer path: x pieces up, row chunks, y chunks down (spmvb200_spmv_host).

// ---- pipelined host path --------------------------------------------------------------------------
// the x-window handle a (kind, handle) pair runs, if any: the handle itself or the copy its first-use pick kept
static const spmvb200_matrix* xwin_of(const spmvb200_matrix* m, int kind) {
    if (kind == SPMVB200_XWIN_ROWS && m->format == SPMVB200_FMT_XWIN) return m;
    if (kind == SPMVB200_CSR_ROWS && m->tuned_x == CAND_XWIN) return m->x_child;
    if (kind == SPMVB200_CSR_ADAPTIVE && m->tuned == CAND_XWIN) return m->xw_child;
    return nullptr;
}
// which kernel a (kind, handle) pair runs chunk-wise: 0/1 stream variants (+10 exact), 2.. vector lanes, 100 ELL column-major,
// 200 x-window (chunks of whole row blocks, one CTA per row block)
static int pipe_candidate(const spmvb200_matrix* m, int kind) {
    if (xwin_of(m, kind)) return 200;
    switch (kind) {
        case SPMVB200_CSR_ROWS: return m->tuned_x == 0 ? 10 : -1;  // the SELL copy runs as one launch
        case SPMVB200_CSR_ADAPTIVE: return ((m->tuned >= 2 && m->nseg) || m->tuned >= 7) ? -1 : m->tuned;  // long rows / spans: one launch
        case SPMVB200_CSR_ROWS_WARP: return m->nseg ? -1 : 20 + m->vec_lanes;
        case SPMVB200_ELL_ROWS: return m->tuned_x == 0 ? 100 : -1;  // the SELL copy runs as one launch
        default: return -1;
    }
}
static void launch_chunk(spmvb200_matrix* m, const HostPipe* p, int k, const double* x, double* y) {
    const uint64_t r0 = p->row_b[k], r1 = p->row_b[k + 1];
    const int c = p->cand;
    if (c == 200) {
        const spmvb200_matrix* xw = xwin_of(m, p->kind);
        launch_xwin(xw, x, y, p->s_comp, nullptr, (uint32_t) (r0 / xw->xw_R), (uint32_t) ((r1 + xw->xw_R - 1) / xw->xw_R));
    } else if (c == 100) launch_ell_colmajor(m, x, y, p->s_comp, r0, r1);
    else if (c == 10) launch_csr_stream<false, 0>(m, x, y, p->s_comp, p->tile_b[k], p->tile_b[k + 1]);
    else if (c == 0) launch_csr_stream<true, 0>(m, x, y, p->s_comp, p->tile_b[k], p->tile_b[k + 1]);
    else if (c == 1) launch_csr_stream<true, 1>(m, x, y, p->s_comp, p->tile_b[k], p->tile_b[k + 1]);
    else if (c >= 20) launch_csr_vector(m, c - 20, x, y, p->s_comp, r0, r1);
    else launch_csr_vector(m, 2 << (c - 2), x, y, p->s_comp, r0, r1);
}
static int host_chunks_wanted(const spmvb200_matrix* m) {
    // measured on B200/PCIe5, cfg2 (16.8 MB vectors): 1 chunk 0.74 ms, 2: 0.63, 4: 0.54, 8: 0.59, 16: 0.70 per call -- a chunk
    // should stay above ~4 MB; longer vectors (cfg4: 268 MB) take more chunks so that the serial head (first x piece) and tail
    // (last y chunk) of the pipeline stay short
    int nch = (int) std::min<uint64_t>(16, std::max<uint64_t>(4, (std::max(m->M, m->N) * 8) >> 24));
    if (const char* e = getenv("SPMVB200_HOST_CHUNKS")) nch = std::max(1, atoi(e));  // developer knob
    return nch;
}
// Cumulative chunk bounds as fractions of the rows.  Short vectors: equal chunks.  Long vectors (>= 64 MB): a RAMP -- the first
// chunks double in size (1, 2, 4, 8 parts), the middle is made of equal large chunks, the last ones halve again -- so that the serial
// head of the pipeline (first x piece up + first kernel chunk, nothing coming down yet) and its serial tail (last kernel chunk + last
// y chunk down, nothing going up any more) are a sixteenth of what equal chunks make them, while the bulk moves in few large copies
// (every copy costs ~20 us of link idle time: tools/duplex_probe.py, 64 equal chunks 6.46 ms against 5.36 ms for one).
// SPMVB200_HOST_SCHED=uniform|ramp overrides (developer knob).
static std::vector<double> host_chunk_fractions(const spmvb200_matrix* m, int& nch) {
    const char* e = getenv("SPMVB200_HOST_SCHED");
    const uint64_t bytes = std::max(m->M, m->N) * 8;
    const bool ramp = e ? !strcmp(e, "ramp") : bytes >= (64ull << 20);
    std::vector<double> w;
    if (ramp && nch >= 10) {
        const int mid = nch - 8;  // 4 doubling chunks at either end
        for (int k = 0; k < 4; ++k) w.push_back((double) (1 << k));
        for (int k = 0; k < mid; ++k) w.push_back(16.0);
        for (int k = 3; k >= 0; --k) w.push_back((double) (1 << k));
    } else {
        w.assign(nch, 1.0);
    }
    nch = (int) w.size();
    double tot = 0;
    for (double v : w) tot += v;
    std::vector<double> cum(nch + 1, 0.0);
    for (int k = 0; k < nch; ++k) cum[k + 1] = cum[k] + w[k] / tot;
    cum[nch] = 1.0;
    return cum;
}
static int build_pipe(spmvb200_matrix* m, int kind, int cand) {
    destroy_pipe(m->pipe);  // rebuilt only when another kind is used through the host path
    m->pipe = nullptr;
    HostPipe* p = new HostPipe();
    m->pipe = p;
    p->kind = kind;
    p->cand = cand;
    int nch = host_chunks_wanted(m);
    p->nch_req = nch;
    if (m->M < 65536 || cand < 0) nch = 1;
    const bool stream = (cand == 0 || cand == 1 || cand == 10);
    const spmvb200_matrix* xw = cand == 200 ? xwin_of(m, kind) : nullptr;
    if (stream) nch = (int) std::max<uint32_t>(1, std::min<uint32_t>(nch, m->ntiles));
    if (xw) nch = (int) std::max<uint32_t>(1, std::min<uint32_t>(nch, xw->xw_nrb));
    const std::vector<double> cum = host_chunk_fractions(m, nch);
    p->nch = nch;
    p->row_b.assign(nch + 1, m->M);
    p->tile_b.assign(nch + 1, m->ntiles);
    p->row_b[0] = 0;
    p->tile_b[0] = 0;
    for (int k = 1; k < nch; ++k) {
        if (stream) {
            uint32_t t = (uint32_t) ((double) m->ntiles * cum[k]);
            // never cut inside the segment run of a long row: its y entry is written by whichever segment finishes last
            while (t < m->ntiles && t > 0 && (m->h_tile_row0[t] & SEG_FLAG) && (m->h_tile_row0[t - 1] & SEG_FLAG) &&
                   (m->h_tile_row0[t] == m->h_tile_row0[t - 1]))
                ++t;
            t = std::max(t, p->tile_b[k - 1]);
            p->tile_b[k] = t;
            p->row_b[k] = m->h_tile_row0[t] & ~SEG_FLAG;
        } else if (xw) {
            p->row_b[k] = std::max<uint64_t>(p->row_b[k - 1], std::min<uint64_t>(m->M, (uint64_t) ((double) xw->xw_nrb * cum[k]) * xw->xw_R));  // whole row blocks
        } else {
            p->row_b[k] = std::max<uint64_t>(p->row_b[k - 1], (uint64_t) ((double) m->M * cum[k]) & ~255ull);
        }
    }
    // x pieces: piece k ends right after the largest column id row chunks 0..k read, so that chunk k can start
    // the moment "its" piece has landed (banded / stencil matrices: pieces ~ equal; unstructured: piece 0 = all of x)
    p->x_b.assign(nch + 1, m->N);
    p->x_b[0] = 0;
    if (nch > 1 && xw) {
        // x-window copy: the columns a chunk reads end with the last window of its row blocks' tile lists (ascending per row block)
        std::vector<uint32_t> t0(xw->xw_nrb + 1), win(std::max<uint32_t>(xw->xw_ntiles, 1));
        CU_TRY(cudaMemcpy(t0.data(), xw->xw_rb_tile0, t0.size() * 4, cudaMemcpyDeviceToHost));
        if (xw->xw_ntiles) CU_TRY(cudaMemcpy(win.data(), xw->xw_tile_win, (size_t) xw->xw_ntiles * 4, cudaMemcpyDeviceToHost));
        for (int k = 0; k < nch; ++k) {
            uint64_t need = 0;
            for (uint64_t rb = p->row_b[k] / xw->xw_R; rb < (p->row_b[k + 1] + xw->xw_R - 1) / xw->xw_R; ++rb)
                if (t0[rb + 1] > t0[rb]) need = std::max<uint64_t>(need, ((uint64_t) win[t0[rb + 1] - 1] + 1) * xw->xw_W);
            p->x_b[k + 1] = std::max(p->x_b[k], k + 1 == nch ? m->N : std::min<uint64_t>(m->N, need));
        }
    } else if (nch > 1) {
        uint32_t* d_cm = nullptr;
        CU_TRY(cudaMalloc(&d_cm, nch * 4));
        CU_TRY(cudaMemset(d_cm, 0, nch * 4));
        for (int k = 0; k < nch; ++k) {
            if (m->format == SPMVB200_FMT_CSR) {
                uint64_t n0, n1;
                if (stream) { n0 = m->h_tile_nnz0[p->tile_b[k]]; n1 = m->h_tile_nnz0[p->tile_b[k + 1]]; }
                else {
                    uint32_t h[2] = {0, 0};
                    CU_TRY(cudaMemcpy(&h[0], m->irp + p->row_b[k], 4, cudaMemcpyDeviceToHost));
                    CU_TRY(cudaMemcpy(&h[1], m->irp + p->row_b[k + 1], 4, cudaMemcpyDeviceToHost));
                    n0 = h[0]; n1 = h[1];
                }
                if (n1 > n0) colmax_flat_kernel<<<592, 256>>>(m->ja, n0, n1, d_cm + k);
            } else if (p->row_b[k + 1] > p->row_b[k]) {
                const uint32_t r0 = (uint32_t) p->row_b[k], r1 = (uint32_t) p->row_b[k + 1];
                colmax_ell_cm_kernel<<<(r1 - r0 + 255) / 256, 256>>>(m->ja, m->rl, m->pitch, r0, r1, d_cm + k);
            }
        }
        std::vector<uint32_t> h_cm(nch);
        cudaError_t e = cudaMemcpy(h_cm.data(), d_cm, nch * 4, cudaMemcpyDeviceToHost);
        cudaFree(d_cm);
        if (e != cudaSuccess) return fail("host-path plan: %s", cudaGetErrorString(e));
        for (int k = 0; k < nch; ++k) {
            const uint64_t need = std::min<uint64_t>(m->N, ((uint64_t) h_cm[k] + 1 + 511) & ~511ull);  // 4 KB granules
            p->x_b[k + 1] = std::max(p->x_b[k], k + 1 == nch ? m->N : need);
        }
    }
    p->npieces = nch;
    CU_TRY(cudaStreamCreateWithFlags(&p->s_up, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&p->s_down, cudaStreamNonBlocking));
    if (const char* e = getenv("SPMVB200_HOST_COPY_STREAMS"))  // developer knob: copies of one direction alternate between two streams
        if (atoi(e) >= 2) {
            CU_TRY(cudaStreamCreateWithFlags(&p->s_up2, cudaStreamNonBlocking));
            CU_TRY(cudaStreamCreateWithFlags(&p->s_down2, cudaStreamNonBlocking));
        }
    p->x_ready.resize(nch);
    p->k_start.resize(nch);
    p->k_end.resize(nch);
    p->k_done.resize(nch);
    for (auto& ev : p->x_ready) CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : p->k_start) CU_TRY(cudaEventCreate(&ev));
    for (auto& ev : p->k_end) CU_TRY(cudaEventCreate(&ev));
    for (auto& ev : p->k_done) CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    return 0;
}

// If the caller's y is page-locked (cudaHostAlloc / cudaHostRegister) a kernel can store its rows straight into it: the
// device->host transfer then rides along with the kernel as posted PCIe writes instead of following it as a copy-engine job.
// Measured on cfg2 (16.8 MB of y, tools/e2e_probe.py, profiles/r01d_e2e_probe.log): a single x-window launch 0.737 -> 0.671 ms per
// call; the sub-warp kernels (one 8-byte store per row and sub-warp) 0.769 -> 0.890 ms; inside the chunked pipeline 0.539 -> 0.628 ms
// (the stores compete with the x pieces still coming up and run at 32-40 GB/s).  So by default only single x-window launches do it.
// SPMVB200_HOST_DIRECT_Y (developer knob): 0 never, 1 default, 2 every single launch, 3 the chunked pipeline as well.
// Returns the device-side alias of y, or nullptr (then y goes through d_y + cudaMemcpyAsync; always so for pageable memory).
static double* mapped_alias(double* y, bool chunked, bool xwin_launch) {
    const char* e = getenv("SPMVB200_HOST_DIRECT_Y");
    const int mode = e ? atoi(e) : 1;
    if (mode < (chunked ? 3 : (xwin_launch ? 1 : 2))) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, y) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? static_cast<double*>(a.devicePointer) : nullptr;
}

// ---- pageable caller buffers --------------------------------------------------------------------
// The reference driver hands malloc'ed x and y (src/main.cu:155,181): pageable memory goes through the driver's staging copies at
// roughly half the pinned rate and cannot overlap with anything.  spmvb200_host_register(p, bytes) page-locks such a buffer IN PLACE
// (cudaHostRegister) until spmvb200_host_unregister(p) -- the one call a driver adds next to its malloc.  It is explicit on purpose:
// a buffer that is free()d while still registered leaves a stale registration behind, and the next allocation that lands on the same
// address makes unrelated CUDA calls fail with "invalid argument" (seen in this repo's own tests with numpy buffers), so the library
// never registers memory whose lifetime it does not know -- unless SPMVB200_HOST_REGISTER=auto asks for it: then a buffer that comes
// back a SECOND time with the same address and size is registered (a harness that reuses its vectors for every repetition and calls
// spmvb200_host_unregister(NULL) / spmvb200_cache_drop(NULL) before freeing them).  At most 16 buffers are tracked.
namespace {
struct HostReg {
    size_t bytes = 0;
    int seen = 0;
    bool registered = false;
};
std::map<const void*, HostReg> g_hostreg;
std::mutex g_hostreg_mu;
}  // namespace
// Page-locked host vectors from the library (cudaHostAlloc): what a driver swaps its malloc / free of x and y for.  Measured on the B200
// box with 268 MB vectors (tools/hostmem_lab.cu): duplex copy 5.36 ms, against 6.35 ms for malloc'ed memory page-locked in place (4 KB
// pages; with transparent huge pages behind the buffer the in-place registration reaches 5.37 ms too) and 37 ms for pageable memory.
extern "C" void* spmvb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, std::max<size_t>(bytes, 8), cudaHostAllocDefault) != cudaSuccess) {
        fail("host_alloc: cudaHostAlloc(%zu) -> %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
extern "C" int spmvb200_host_free(void* p) {
    if (p) CU_TRY(cudaFreeHost(p));
    return 0;
}
extern "C" int spmvb200_host_register(const void* p, size_t bytes) {
    if (!p || !bytes) return fail("host_register: null buffer");
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return fail("host_register: cannot query the pointer"); }
    if (a.type != cudaMemoryTypeUnregistered) return 0;  // already page-locked (cudaHostAlloc / registered by the caller)
    std::lock_guard<std::mutex> lk(g_hostreg_mu);
    if (g_hostreg.size() >= 16 && !g_hostreg.count(p)) return fail("host_register: 16 buffers are registered already");
    cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail("host_register: cudaHostRegister -> %s", cudaGetErrorString(e));
    }
    HostReg& r = g_hostreg[p];
    r.bytes = bytes;
    r.seen = 2;
    r.registered = true;
    return 0;
}
static void host_buffer_seen(const void* p, size_t bytes) {
    static const bool automatic = getenv("SPMVB200_HOST_REGISTER") && !strcmp(getenv("SPMVB200_HOST_REGISTER"), "auto");
    if (!automatic || bytes < (1u << 20)) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (a.type != cudaMemoryTypeUnregistered) return;  // already page-locked (or not host memory at all)
    std::lock_guard<std::mutex> lk(g_hostreg_mu);
    if (g_hostreg.size() >= 16 && !g_hostreg.count(p)) return;
    HostReg& r = g_hostreg[p];
    if (r.bytes != bytes) { r.bytes = bytes; r.seen = 0; r.registered = false; }
    if (r.seen < 0 || ++r.seen < 2) return;
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) r.registered = true;
    else { cudaGetLastError(); r.seen = -1; }  // e.g. read-only mapping: stay on the pageable path for this buffer
}
extern "C" int spmvb200_host_registered(const void* p) {
    std::lock_guard<std::mutex> lk(g_hostreg_mu);
    auto it = g_hostreg.find(p);
    return it != g_hostreg.end() && it->second.registered ? 1 : 0;
}
extern "C" int spmvb200_host_unregister(const void* p) {
    std::lock_guard<std::mutex> lk(g_hostreg_mu);
    for (auto it = g_hostreg.begin(); it != g_hostreg.end();) {
        if (!p || it->first == p) {
            if (it->second.registered && cudaHostUnregister(const_cast<void*>(it->first)) != cudaSuccess) cudaGetLastError();
            it = g_hostreg.erase(it);
        } else {
            ++it;
        }
    }
    return 0;
}

extern "C" int spmvb200_spmv_host(spmvb200_matrix* m, int kind, const double* x, double* y, float* kernel_ms) {
    if (!m || !x || !y) return fail("spmv_host: null argument");
    host_buffer_seen(x, m->N * 8);
    host_buffer_seen(y, m->M * 8);
    if (!spmvb200_kind_supported(m, kind)) return fail("kind %d (%s) cannot run on format %d", kind, spmvb200_kind_name(kind), m->format);
    if (prefer_smem_once() || ensure_events(m)) return 1;
    if (!m->d_x) CU_TRY(cudaMalloc(&m->d_x, std::max<uint64_t>(m->N, 1) * 8));
    if (!m->d_y) CU_TRY(cudaMalloc(&m->d_y, std::max<uint64_t>(m->M, 1) * 8));
    const bool untuned = kind != SPMVB200_XWIN_ROWS && needs_tuning(m, kind);
    int cand = untuned ? -1 : pipe_candidate(m, kind);
    if (!untuned && (!m->pipe || m->pipe->kind != kind || m->pipe->cand != cand || m->pipe->nch_req != host_chunks_wanted(m)))
        if (build_pipe(m, kind, cand)) return 1;
    if (untuned || m->pipe->nch <= 1) {  // plain path: x up, one launch, y down (also the adaptive mode's tuning call)
        const bool xw = (kind == SPMVB200_XWIN_ROWS && m->xw_mode >= 0) || (kind == SPMVB200_CSR_ROWS && m->tuned_x == CAND_XWIN) ||
                        (kind == SPMVB200_CSR_ADAPTIVE && m->tuned == CAND_XWIN);
        const bool tuning = untuned || (kind == SPMVB200_XWIN_ROWS && m->xw_mode < 0);  // tuning launches time candidates: those stay on device memory
        double* const y_one = tuning ? nullptr : mapped_alias(y, false, xw);
        double* const out = y_one ? y_one : m->d_y;
        CU_TRY(cudaMemcpyAsync(m->d_x, x, m->N * 8, cudaMemcpyHostToDevice, 0));
        CU_TRY(cudaEventRecord(m->ev0, 0));
        if (launch(m, kind, m->d_x, out, 0)) return 1;
        CU_TRY(cudaEventRecord(m->ev1, 0));
        if (out == m->d_y) CU_TRY(cudaMemcpyAsync(y, m->d_y, m->M * 8, cudaMemcpyDeviceToHost, 0));
        CU_TRY(cudaStreamSynchronize(0));
        if (kernel_ms) CU_TRY(cudaEventElapsedTime(kernel_ms, m->ev0, m->ev1));
        return 0;
    }
    HostPipe* p = m->pipe;
    double* const y_map = mapped_alias(y, true, false);
    const bool dbg = getenv("SPMVB200_PIPE_DEBUG") != nullptr;
    const bool timed = kernel_ms != nullptr || dbg;  // time-stamped events between the chunks slow the pipeline down by ~10 %: only on request
    std::vector<cudaEvent_t> dbg_up, dbg_down;
    cudaEvent_t dbg0 = nullptr;
    if (dbg) {
        cudaEventCreate(&dbg0);
        cudaEventRecord(dbg0, p->s_up);
    }
    for (int j = 0; j < p->nch; ++j) {
        const uint64_t o = p->x_b[j], n = p->x_b[j + 1] - o;
        cudaStream_t su = (p->s_up2 && (j & 1)) ? p->s_up2 : p->s_up;
        if (n) CU_TRY(cudaMemcpyAsync(m->d_x + o, x + o, n * 8, cudaMemcpyHostToDevice, su));
        CU_TRY(cudaEventRecord(p->x_ready[j], su));
        if (dbg) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, su); dbg_up.push_back(e); }
    }
    for (int k = 0; k < p->nch; ++k) {
        CU_TRY(cudaStreamWaitEvent(p->s_comp, p->x_ready[k], 0));
        if (timed) CU_TRY(cudaEventRecord(p->k_start[k], p->s_comp));
        static const bool no_kernel = getenv("SPMVB200_PIPE_NO_KERNEL") != nullptr;  // developer knob: copies only (timeline experiments)
        if (!no_kernel) launch_chunk(m, p, k, m->d_x, y_map ? y_map : m->d_y);
        cudaEvent_t const kdone = timed ? p->k_end[k] : p->k_done[k];
        CU_TRY(cudaEventRecord(kdone, p->s_comp));
        const uint64_t r0 = p->row_b[k], r1 = p->row_b[k + 1];
        if (r1 > r0 && !y_map) {
            cudaStream_t sd = (p->s_down2 && (k & 1)) ? p->s_down2 : p->s_down;
            CU_TRY(cudaStreamWaitEvent(sd, kdone, 0));
            CU_TRY(cudaMemcpyAsync(y + r0, m->d_y + r0, (r1 - r0) * 8, cudaMemcpyDeviceToHost, sd));
            if (dbg) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, sd); dbg_down.push_back(e); }
        }
    }
    CU_TRY(cudaPeekAtLastError());
    CU_TRY(cudaStreamSynchronize(p->s_comp));
    CU_TRY(cudaStreamSynchronize(p->s_down));
    CU_TRY(cudaStreamSynchronize(p->s_up));
    if (p->s_up2) {
        CU_TRY(cudaStreamSynchronize(p->s_down2));
        CU_TRY(cudaStreamSynchronize(p->s_up2));
    }
    if (dbg) {
        float ms;
        fprintf(stderr, "pipe: x piece bounds=");
        for (int k = 0; k <= p->nch; ++k) fprintf(stderr, "%llu ", (unsigned long long) p->x_b[k]);
        fprintf(stderr, "\n  up done at:");
        for (auto e : dbg_up) { cudaEventElapsedTime(&ms, dbg0, e); fprintf(stderr, " %.3f", ms); cudaEventDestroy(e); }
        fprintf(stderr, "\n  kernels [start,end]:");
        for (int k = 0; k < p->nch; ++k) { float a, b; cudaEventElapsedTime(&a, dbg0, p->k_start[k]); cudaEventElapsedTime(&b, dbg0, p->k_end[k]); fprintf(stderr, " [%.3f,%.3f]", a, b); }
        fprintf(stderr, "\n  down done at:");
        for (auto e : dbg_down) { cudaEventElapsedTime(&ms, dbg0, e); fprintf(stderr, " %.3f", ms); cudaEventDestroy(e); }
        fprintf(stderr, "\n");
        cudaEventDestroy(dbg0);
    }
    if (kernel_ms) {
        float tot = 0;
        for (int k = 0; k < p->nch; ++k) {
            float ms = 0;
            CU_TRY(cudaEventElapsedTime(&ms, p->k_start[k], p->k_end[k]));
            tot += ms;
        }
        *kernel_ms = tot;
    }
    return 0;
}
