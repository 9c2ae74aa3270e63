// engine_launch.inl -- part of engine.cu (textually included there: one translation unit, file-local helpers stay static).
// launch: kernel launchers, first-use tuning of the self-tuning kinds, the kind -> kernel switch.

// ------------------------------------------------------------------------------------------------- launch
// Per-launch state handed down to the launchers (no globals: two host threads may drive two handles at once).
//   push  : fused output delivery of spmvb200_spmv_device_push (n == 0: none)
//   fused : set by a launcher whose kernel delivered the rows from its own epilogue
struct LaunchCtx {
    PushArgs push = {};
    bool fused = false;
};
// L1 policy of the gather-bound kernels (common.cuh ld_mat / ld_xp): 0 default loads, 1 stream = L1::no_allocate + x = L1::evict_last.
// Per handle (m->l1pol, picked at first use on skewed matrices); SPMVB200_L1_POLICY forces it process-wide (developer knob).
static int l1pol_of(const spmvb200_matrix* m) {
    static const int env = getenv("SPMVB200_L1_POLICY") ? atoi(getenv("SPMVB200_L1_POLICY")) : -1;
    return env >= 0 ? env : m->l1pol;
}
static const PushArgs NO_PUSH = {};
static inline const PushArgs& push_of(const LaunchCtx* lc) { return lc ? lc->push : NO_PUSH; }

// two timing events that are destroyed on every exit path
struct ScopedEvents {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int init() {
        CU_TRY(cudaEventCreate(&e0));
        CU_TRY(cudaEventCreate(&e1));
        return 0;
    }
    ~ScopedEvents() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};
// a child handle that is freed unless it is released to its new owner
struct ScopedChild {
    spmvb200_matrix* m = nullptr;
    ~ScopedChild() { if (m) spmvb200_free(m); }
    spmvb200_matrix* release() { spmvb200_matrix* r = m; m = nullptr; return r; }
};

// How the self-tuning kinds pick their kernel at first use (spmvb200_set_tuning_mode / SPMVB200_TUNE):
//   0 timed         : candidates are timed on the caller's vectors (best of two); fastest pick, may differ between runs
//   1 deterministic : the pick is a pure function of the matrix structure (rules below) -- tolerance kinds then return the same
//                     bits in every process, and nothing is timed (no event synchronisation inside the first launch)
static int g_tune_mode = -1;
static int tune_mode() {
    if (g_tune_mode < 0) {
        const char* e = getenv("SPMVB200_TUNE");
        g_tune_mode = (e && (e[0] == 'd' || e[0] == '1')) ? 1 : 0;
    }
    return g_tune_mode;
}
// Rows the vector kernels skip: medium rows (one CTA each) and rows longer than a tile (one CTA per segment).  They touch other rows
// of y than the main kernel, so they run NEXT to it on two side streams -- forked before the main launch, joined after it (events:
// legal inside a graph capture too).  On R-MAT (cfg3) the three kernels are 189 + 97 + 52 us back to back.
// exact: the medium rows go through the warp-per-row kernel that adds in the serial order (the bit-exact kind's SELL hybrid).
static void tail_fork(spmvb200_matrix* m, const double* x, double* y, cudaStream_t st, bool exact = false) {
    if (!m->nmid && !m->nseg) return;
    static const bool serial = getenv("SPMVB200_SERIAL_TAIL") != nullptr;  // developer knob: the old back-to-back order
    if (!serial && !m->e_fork) {
        bool ok = cudaEventCreateWithFlags(&m->e_fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 2 && ok; ++i)
            ok = cudaStreamCreateWithFlags(&m->s_tail[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&m->e_tail[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); m->e_fork = nullptr; }
    }
    const bool fork = !serial && m->e_fork && m->s_tail[0] && m->s_tail[1] && m->e_tail[0] && m->e_tail[1];
    if (fork) cudaEventRecord(m->e_fork, st);
    if (m->nmid) {
        cudaStream_t s = fork ? m->s_tail[0] : st;
        if (fork) cudaStreamWaitEvent(s, m->e_fork, 0);
        static const bool warp_mid = getenv("SPMVB200_NO_WARP_MID") == nullptr;  // developer knob
        const uint32_t lo = warp_mid ? (uint32_t) MIDW_MAX : 0u;  // rows up to MIDW_MAX: a warp each; longer: a CTA each
        if (exact) {
            csr_midrow_exact_kernel<256><<<(m->nmid + 7) / 8, 256, 0, s>>>(m->mid_rows, m->nmid, m->irp, m->ja, m->as, x, y);
            ++g_launches;
        } else if (warp_mid) {
            if (l1pol_of(m)) csr_midrow_warp_kernel<256, 1><<<(m->nmid + 7) / 8, 256, 0, s>>>(m->mid_rows, m->nmid, m->irp, m->ja, m->as, x, y, 0u, (uint32_t) MIDW_MAX);
            else csr_midrow_warp_kernel<256, 0><<<(m->nmid + 7) / 8, 256, 0, s>>>(m->mid_rows, m->nmid, m->irp, m->ja, m->as, x, y, 0u, (uint32_t) MIDW_MAX);
            ++g_launches;
        }
        if (!exact && (!warp_mid || m->lmax > (uint32_t) MIDW_MAX)) {
            if (l1pol_of(m)) csr_midrow_kernel<128, 1><<<m->nmid, 128, 0, s>>>(m->mid_rows, m->irp, m->ja, m->as, x, y, lo);
            else csr_midrow_kernel<128, 0><<<m->nmid, 128, 0, s>>>(m->mid_rows, m->irp, m->ja, m->as, x, y, lo);
            ++g_launches;
        }
        if (fork) cudaEventRecord(m->e_tail[0], s);
    }
    if (m->nseg) {
        cudaStream_t s = fork ? m->s_tail[1] : st;
        if (fork) cudaStreamWaitEvent(s, m->e_fork, 0);
        if (l1pol_of(m)) csr_longrow_kernel<128, 1><<<m->nseg, 128, 0, s>>>(m->seg_tiles, m->desc, m->longrec, m->ja, m->as, x, y, m->partial, m->ticket);
        else csr_longrow_kernel<128, 0><<<m->nseg, 128, 0, s>>>(m->seg_tiles, m->desc, m->longrec, m->ja, m->as, x, y, m->partial, m->ticket);
        if (fork) cudaEventRecord(m->e_tail[1], s);
        ++g_launches;
    }
}
static void tail_join(spmvb200_matrix* m, cudaStream_t st) {
    static const bool serial = getenv("SPMVB200_SERIAL_TAIL") != nullptr;
    if (!m->e_fork || serial) return;
    if (m->nmid) cudaStreamWaitEvent(st, m->e_tail[0], 0);
    if (m->nseg) cudaStreamWaitEvent(st, m->e_tail[1], 0);
}
// rows [r0, r1) (whole matrix: 0, M); y is always indexed by the handle's row number
template <int LANES>
static void launch_csr_vector_t(spmvb200_matrix* m, const double* x, double* y, cudaStream_t st, uint64_t r0, uint64_t r1) {
    constexpr int BLOCK = 256;
    const uint64_t threads = (r1 - r0) * LANES;
    if (!threads) return;
    // whole-matrix launches leave rows longer than VEC_MID to the per-row CTAs below; row-chunk launches keep them
    const uint32_t maxlen = (uint32_t) ((r0 == 0 && r1 == m->M) ? VEC_MID : STREAM_TILE);
    if (l1pol_of(m)) csr_vector_kernel<LANES, BLOCK, 1><<<(unsigned) ((threads + BLOCK - 1) / BLOCK), BLOCK, 0, st>>>(m->irp, m->ja, m->as, x, y, (uint32_t) r0, (uint32_t) r1, maxlen);
    else csr_vector_kernel<LANES, BLOCK, 0><<<(unsigned) ((threads + BLOCK - 1) / BLOCK), BLOCK, 0, st>>>(m->irp, m->ja, m->as, x, y, (uint32_t) r0, (uint32_t) r1, maxlen);
    ++g_launches;
}
// vector kernel for rows up to one tile + the long-row kernel for the rest (same stream, back to back)
static void launch_csr_vector(spmvb200_matrix* m, int lanes, const double* x, double* y, cudaStream_t st, uint64_t r0, uint64_t r1) {
    const bool whole = r0 == 0 && r1 == m->M;
    if (whole) tail_fork(m, x, y, st);
    switch (lanes) {
        case 2: launch_csr_vector_t<2>(m, x, y, st, r0, r1); break;
        case 4: launch_csr_vector_t<4>(m, x, y, st, r0, r1); break;
        case 8: launch_csr_vector_t<8>(m, x, y, st, r0, r1); break;
        case 16: launch_csr_vector_t<16>(m, x, y, st, r0, r1); break;
        default: launch_csr_vector_t<32>(m, x, y, st, r0, r1); break;
    }
    if (whole) tail_join(m, st);
}
template <int LANES>
static void launch_csr_vspan_t(spmvb200_matrix* m, const double* x, double* y, cudaStream_t st) {
    csr_vector_span_kernel<LANES, 1024><<<m->nspans, 1024, 0, st>>>(m->span_b, m->irp, m->ja, m->as, x, y, (uint32_t) VEC_MID);
    ++g_launches;
}
static void launch_csr_vspan(spmvb200_matrix* m, int lanes, const double* x, double* y, cudaStream_t st) {
    tail_fork(m, x, y, st);
    switch (lanes) {
        case 2: launch_csr_vspan_t<2>(m, x, y, st); break;
        case 4: launch_csr_vspan_t<4>(m, x, y, st); break;
        case 8: launch_csr_vspan_t<8>(m, x, y, st); break;
        case 16: launch_csr_vspan_t<16>(m, x, y, st); break;
        default: launch_csr_vspan_t<32>(m, x, y, st); break;
    }
    tail_join(m, st);
}
// tiles [t0, t1) (whole matrix: 0, ntiles)
template <bool ADAPT, int VARIANT>
static void launch_csr_stream(const spmvb200_matrix* m, const double* x, double* y, cudaStream_t st, uint32_t t0, uint32_t t1) {
    if (t1 <= t0) return;
    static const uint32_t pre_t = getenv("SPMVB200_PRE_T") ? (uint32_t) atoi(getenv("SPMVB200_PRE_T")) : (uint32_t) STREAM_PRE_T;  // developer knob
    csr_stream_kernel<STREAM_TILE, STREAM_BLOCK, STREAM_TILE_ROWS, ADAPT, VARIANT>
        <<<t1 - t0, STREAM_BLOCK, 0, st>>>(m->desc, m->longrec, m->irp, m->ja, m->as, x, y, m->partial, m->ticket, t0, pre_t);
    ++g_launches;
}
static void launch_ell_colmajor(const spmvb200_matrix* m, const double* x, double* y, cudaStream_t st, uint64_t r0, uint64_t r1, LaunchCtx* lc = nullptr) {
    constexpr int BLOCK = 256;
    if (r1 <= r0) return;
    const PushArgs& g_push = push_of(lc);
    static const bool no_exit = getenv("SPMVB200_ELL_NO_EARLY_EXIT") != nullptr;  // developer knob: walk all K slots like the reference
    // slots in flight per thread: 4, or 3 when that leaves a shorter tail of one-at-a-time slots (K = 27: 3 x 9 exactly; measured
    // 103.3 vs 105.2 us on cfg2; 5..9 lose more to occupancy than they gain)
    static const int unroll_env = getenv("SPMVB200_ELL_UNROLL") ? atoi(getenv("SPMVB200_ELL_UNROLL")) : 0;  // developer knob
    const int unroll16 = unroll_env ? unroll_env : ((m->K % 3) < (m->K % 4) ? 3 : 4);
    const unsigned grid = (unsigned) ((r1 - r0 + BLOCK - 1) / BLOCK);
#define ELL16(U) ell_colmajor_kernel<U, BLOCK, true><<<grid, BLOCK, 0, st>>>(m->as, m->ja16, m->rl, m->pitch, (uint32_t) r0, (uint32_t) r1, (uint32_t) m->K, m->ja16_base, x, y, g_push)
    // 16-bit ids: two rows per thread, 3 slots in flight (measured on cfg2: 96.9 us; one row per thread 103.0; pair x2 100.7, pair x4 107.1)
    // Short rows (K <= 6: 5-point / 7-point stencils) take ALL their slots in one batch -- one DRAM round trip per thread instead of two is
    // what an isolated launch of a small matrix is made of: cfg1 (K = 5, L2 flushed) 20.6 -> 18.5 us, 0.62 -> 0.69 of the roofline
    // (profiles/r02g_cfg1_ell_variants.log).
    static const int pair_knob = getenv("SPMVB200_ELL_PAIR") ? atoi(getenv("SPMVB200_ELL_PAIR")) : -1;  // developer knob: 0 = one row per thread
    const int pair_env = pair_knob >= 0 ? pair_knob : (m->K >= 2 && m->K <= 6 ? (int) m->K : 3);
    if (m->ja16 && !no_exit && pair_env && (r0 % 2) == 0) {
        const unsigned g2 = (unsigned) (((r1 - r0 + 1) / 2 + BLOCK - 1) / BLOCK);
#define ELLP(U, S) ell_colmajor_pair_kernel<U, BLOCK, S><<<g2, BLOCK, 0, st>>>(m->as, m->ja16, m->rl, m->pitch, (uint32_t) r0, (uint32_t) r1, m->ja16_base, x, y, g_push)
        // speculative first batch (fetched before the row lengths arrive) when the rectangle is nearly full: see the kernel
        static const bool no_spec = getenv("SPMVB200_ELL_NO_SPEC") != nullptr;  // developer knob
        const bool spec = !no_spec && m->K >= (uint64_t) pair_env && (double) m->K * (double) m->M <= 1.15 * (double) m->NZ;
        switch (pair_env) {
            case 2: if (spec) ELLP(2, true); else ELLP(2, false); break;
            case 3: if (spec) ELLP(3, true); else ELLP(3, false); break;
            case 5: if (spec) ELLP(5, true); else ELLP(5, false); break;
            case 6: if (spec) ELLP(6, true); else ELLP(6, false); break;
            default: if (spec) ELLP(4, true); else ELLP(4, false); break;
        }
#undef ELLP
    } else if (m->ja16 && !no_exit) {
        switch (unroll16) {
            case 3: ELL16(3); break;
            case 5: ELL16(5); break;
            case 6: ELL16(6); break;
            case 8: ELL16(8); break;
            case 9: ELL16(9); break;
            default: ELL16(4); break;
        }
    }
#undef ELL16
    else if (unroll16 == 3)
        ell_colmajor_kernel<3, BLOCK, false><<<grid, BLOCK, 0, st>>>(m->as, m->ja, no_exit ? nullptr : m->rl, m->pitch, (uint32_t) r0, (uint32_t) r1, (uint32_t) m->K, 0, x, y, g_push);
    else
        ell_colmajor_kernel<4, BLOCK, false><<<grid, BLOCK, 0, st>>>(m->as, m->ja, no_exit ? nullptr : m->rl, m->pitch, (uint32_t) r0, (uint32_t) r1, (uint32_t) m->K, 0, x, y, g_push);
    if (lc) lc->fused = true;
    ++g_launches;
}


// ---- x-window CSR: one CTA per row block, NW warps, ring of x windows in shared memory
template <int NW, int ACC, int UMAX>
static int launch_xwin_t(const spmvb200_matrix* m, const double* x, double* y, cudaStream_t st, LaunchCtx* lc, uint32_t rb0, uint32_t rb1) {
    const size_t smem = (size_t) m->xw_nbuf * (m->xw_W + 2) * 8 + XW_MAX_NBUF * 12;
    static size_t configured[64] = {0};  // per device: function attributes belong to the device's context
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    if (smem > configured[dev & 63]) {
        CU_TRY(cudaFuncSetAttribute(xwin_kernel<NW, ACC, UMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        configured[dev & 63] = smem;
    }
    // a row-block sub-range (chunked host path) always runs one CTA per row block
    const bool whole = rb0 == 0 && rb1 >= m->xw_nrb;
    const bool persist = m->xw_mode == 1 && whole;
    if (rb1 > m->xw_nrb) rb1 = m->xw_nrb;
    if (rb1 <= rb0) return 0;
    xwin_kernel<NW, ACC, UMAX><<<persist ? m->xw_ncta : rb1 - rb0, 32 * NW, smem, st>>>(persist ? m->xw_cta_rb : nullptr, m->xw_rb_tile0, m->xw_tile_win, m->xw_grp_off, m->xw_cnt, m->xw_col, m->as, x, y,
                                                                 (uint32_t) m->M, (uint32_t) m->N, m->xw_W, m->xw_nbuf, ((uintptr_t) x & 15) == 0, rb0, push_of(lc));
    if (lc) lc->fused = true;
    ++g_launches;
    return 0;
}
// (warps, row groups per warp, largest batch in slots): up to 2*UMAX*ACC loads in flight per lane
static int launch_xwin(const spmvb200_matrix* m, const double* x, double* y, cudaStream_t st, LaunchCtx* lc = nullptr, uint32_t rb0 = 0, uint32_t rb1 = 0xffffffffu) {
    const uint32_t acc = m->xw_R / (32 * m->xw_nw);
    static const int u_env = getenv("SPMVB200_XW_U") ? atoi(getenv("SPMVB200_XW_U")) : 0;  // developer knob
#define XW_CASE(NW, ACC, UDEF, UALT, UALT2)                                                  \
    if (m->xw_nw == NW && acc == ACC) {                                                      \
        if (u_env == UALT) return launch_xwin_t<NW, ACC, UALT>(m, x, y, st, lc, rb0, rb1);   \
        if (u_env == UALT2) return launch_xwin_t<NW, ACC, UALT2>(m, x, y, st, lc, rb0, rb1); \
        return launch_xwin_t<NW, ACC, UDEF>(m, x, y, st, lc, rb0, rb1);                      \
    }
    XW_CASE(32, 1, 8, 6, 4)
    XW_CASE(32, 2, 5, 4, 6)
    XW_CASE(32, 4, 3, 2, 4)
    XW_CASE(16, 1, 8, 6, 4)
    XW_CASE(16, 2, 8, 6, 4)
    XW_CASE(16, 4, 6, 4, 8)
    XW_CASE(16, 8, 3, 4, 2)
#undef XW_CASE
    return fail("x-window kernel: no instantiation for R=%u, %u warps", m->xw_R, m->xw_nw);
}

// first use of an x-window handle: one CTA per row block, or persistent CTAs?  (Persistent wins when a row block has few
// tiles -- no pipeline refill per row block; per-row-block wins on wide bands, where concurrent CTAs then share windows.)
// Deterministic rule (measured: cfg2 as CSR, 5 tiles per row block, and cfg4 with w = 2^12, 3 tiles: persistent +8 %;
// cfg4 with w = 2^15, 9 tiles: per-row-block).
static int xwin_mode_rule(const spmvb200_matrix* m) { return (uint64_t) m->xw_ntiles <= 6ull * std::max<uint32_t>(m->xw_nrb, 1) ? 1 : 0; }
static int tune_xwin(spmvb200_matrix* m, const double* x, double* y, cudaStream_t st, float* best_ms_out) {
    if (tune_mode() == 1) {
        m->xw_mode = xwin_mode_rule(m);
        if (best_ms_out) *best_ms_out = 0.f;
        return 0;
    }
    ScopedEvents ev;
    if (ev.init()) return 1;
    float best_ms = 1e30f;
    int best = 0;
    for (int mode = 0; mode < 2; ++mode) {
        m->xw_mode = mode;
        if (launch_xwin(m, x, y, st)) return 1;
        float ms_min = 1e30f;
        for (int rep = 0; rep < 2; ++rep) {
            CU_TRY(cudaEventRecord(ev.e0, st));
            if (launch_xwin(m, x, y, st)) return 1;
            CU_TRY(cudaEventRecord(ev.e1, st));
            CU_TRY(cudaEventSynchronize(ev.e1));
            float ms = 0;
            CU_TRY(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
            ms_min = std::min(ms_min, ms);
        }
        if (ms_min < best_ms) { best_ms = ms_min; best = mode; }
    }
    m->xw_mode = best;
    if (best_ms_out) *best_ms_out = best_ms;
    return 0;
}

// SELL-32-sigma, thread per row.  A stand-alone SELL handle built from a CSR parent (spmvb200_sell_from_csr) leaves rows longer than
// VEC_MID out of the slices and keeps the parent's arrays for them: those rows run on the per-row kernels next to the slices.
static void launch_sell(const spmvb200_matrix* m, const double* x, double* y, cudaStream_t st) {
    if (l1pol_of(m)) sell_kernel<4, 256, 1><<<(unsigned) ((m->Mpad + 255) / 256), 256, 0, st>>>(m->irp, m->perm, m->rl, m->as, m->ja, (uint32_t) m->Mpad, x, y);
    else sell_kernel<4, 256, 0><<<(unsigned) ((m->Mpad + 255) / 256), 256, 0, st>>>(m->irp, m->perm, m->rl, m->as, m->ja, (uint32_t) m->Mpad, x, y);
    ++g_launches;
}

// ---- SPMVB200_CSR_ADAPTIVE: candidates; the fastest on this matrix is picked at first use
static const int N_CAND = 15;
static const int CAND_HOTX = 14;  // hot-x hybrid (hotx.cuh): power-law column popularity, one persistent kernel with the hot x in shared memory
static const int CAND_XWIN = 12;  // x-window copy (built during tuning when the tile census says it can pay off)
static const int CAND_SELL = 13;  // SELL-32-sigma copy (matrices without long rows: thread per row, coalesced, no shuffles)
static const char* CAND_NAME[N_CAND] = {"stream/8cta", "stream/bigL1", "vector/2", "vector/4", "vector/8", "vector/16", "vector/32",
                                        "vspan/2", "vspan/4", "vspan/8", "vspan/16", "vspan/32", "xwindow", "sell", "hotx"};
static int cand_lanes(int c) { return c < 2 ? 0 : 2 << ((c - 2) % 5); }
static int cand_of_lanes(int lanes) { int c = 2; while (c < 6 && cand_lanes(c) < lanes) ++c; return c; }
static int launch_hotx(spmvb200_matrix* m, const double* x, double* y, cudaStream_t st) {
    const spmvb200_matrix* sl = m->hot_sell;
    HotxArgs a = {};
    a.slice_ptr = sl->irp;
    a.perm = sl->perm;
    a.rl = sl->rl;
    a.sas = sl->as;
    a.sja = sl->ja;
    a.nslices = (uint32_t) (sl->Mpad / 32);
    a.irp = m->irp;
    a.ja = m->ja_hot;
    a.as = m->as;
    a.mid_rows = m->mid_rows;
    a.nmid = m->nmid;
    a.seg_tiles = m->seg_tiles;
    a.nseg = m->nseg;
    a.desc = m->desc;
    a.longrec = m->longrec;
    a.partial = m->partial;
    a.ticket = m->ticket;
    a.hot_cols = m->hot_cols;
    a.slice_order = m->hot_slice_order;
    static int sms[64] = {0};
    static bool configured[64] = {false};
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        CU_TRY(cudaFuncSetAttribute(hotx_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        CU_TRY(cudaFuncSetAttribute(hotx_kernel<512, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
        CU_TRY(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
        configured[dev & 63] = true;
    }
    // the hot columns are stored by decreasing popularity: a smaller cache is a prefix of a bigger one, ids >= H_used but < hot_H
    // ... are NOT cold ids, so the kernel is always given the H the ids were remapped with and a shape whose cache holds it
    if (m->hot_shape == 1 && m->hot_H <= 8192) {
        a.H = m->hot_H;
        hotx_kernel<512, 3><<<3 * sms[dev & 63], 512, (size_t) a.H * 8, st>>>(a, x, y);
    } else {
        a.H = m->hot_H;
        hotx_kernel<1024, 1><<<sms[dev & 63], 1024, (size_t) a.H * 8, st>>>(a, x, y);
    }
    ++g_launches;
    return 0;
}
static void launch_candidate(spmvb200_matrix* m, int c, const double* x, double* y, cudaStream_t st, LaunchCtx* lc = nullptr) {
    if (c == CAND_HOTX) { launch_hotx(m, x, y, st); return; }
    if (c == 0) launch_csr_stream<true, 0>(m, x, y, st, 0, m->ntiles);
    else if (c == 1) launch_csr_stream<true, 1>(m, x, y, st, 0, m->ntiles);
    else if (c < 7) launch_csr_vector(m, cand_lanes(c), x, y, st, 0, m->M);
    else if (c < CAND_XWIN) launch_csr_vspan(m, cand_lanes(c), x, y, st);
    else if (c == CAND_XWIN) launch_xwin(m->xw_child, x, y, st, lc);
    else {
        const bool hybrid = m->lmax > (uint32_t) VEC_MID;  // rows the capped SELL copy left out (it does not write their y)
        if (hybrid) tail_fork(m, x, y, st);
        launch_sell(m->xw_child, x, y, st);
        if (hybrid) tail_join(m, st);
    }
}

// Re-tiled copies a CSR handle may keep next to its arrays.  exact: the copy must sum in the serial order (x-window: column-sorted
// rows only).  Returns 0 and the child, or non-zero when the matrix does not fit the format ("not local enough", too much padding).
static int build_child(spmvb200_matrix* m, int cand, bool exact, spmvb200_matrix** out) {
    *out = nullptr;
    ScopedChild c;
    const int quiet0 = g_quiet;
    g_quiet = 1;  // "does not fit this format" is an expected answer here, not an error to print
    int rc = 1;
    if (cand == CAND_XWIN && !getenv("SPMVB200_NO_XWINDOW")) {
        // window traffic (L2 -> shared memory) above ~1.5x the matrix stream cannot win: stop at the tile census
        rc = xwin_build(m, 0, 0, 15.0, &c.m);
        if (!rc && exact && !c.m->xw_sorted) rc = 1;
    } else if (cand == CAND_SELL && !getenv("SPMVB200_NO_SELL")) {
        // rows longer than VEC_MID are left out of the copy: per-row / per-segment CTAs take them (hybrid for skewed matrices);
        // kept only while the slices stay nearly padding-free
        rc = sell_build(m, 0, m->lmax <= (uint32_t) VEC_MID ? 0xffffffffu : (uint32_t) VEC_MID, &c.m);
        if (!rc && c.m->slots > m->NZ + m->NZ / 4) rc = 1;
    }
    g_quiet = quiet0;
    g_err[0] = 0;
    if (rc) return 1;
    *out = c.release();
    return 0;
}

template <typename F>
static int time_best_of_2(F&& run, cudaStream_t st, float* ms_out) {
    ScopedEvents ev;
    if (ev.init()) return 1;
    run();
    float best = 1e30f;
    for (int rep = 0; rep < 2; ++rep) {
        CU_TRY(cudaEventRecord(ev.e0, st));
        run();
        CU_TRY(cudaEventRecord(ev.e1, st));
        CU_TRY(cudaEventSynchronize(ev.e1));
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
        best = std::min(best, ms);
    }
    *ms_out = best;
    return 0;
}

// shape 0: 16384 hot columns, 1024 threads x 1 CTA per SM; shape 1: 8192 hot columns, 512 threads x 3 CTAs per SM
static int hotx_build_quiet(spmvb200_matrix* m, int shape = -1) {
    if (shape < 0) shape = getenv("SPMVB200_HOTX_SHAPE") ? atoi(getenv("SPMVB200_HOTX_SHAPE")) : 1;  // developer knob; default: the measured winner
    const int quiet0 = g_quiet;
    g_quiet = 1;
    const int rc = hotx_build(m, shape == 1 ? 8192 : 16384, 0.25);
    g_quiet = quiet0;
    g_err[0] = 0;
    if (!rc) m->hot_shape = shape == 1 ? 1 : 0;
    return rc;
}
static int tune_adaptive(spmvb200_matrix* m, const double* d_x, double* d_y, cudaStream_t st) {
    for (int c = 0; c < N_CAND; ++c) m->tuned_ms[c] = -1.f;
    const bool big = m->NZ >= (1u << 20);  // a second copy of the matrix (10.x B per non-zero) only when the matrix is big enough to matter
    if (tune_mode() == 1) {
        // deterministic: x-window copy if the matrix is local enough, else a nearly padding-free SELL copy, else the sub-warp kernel
        // with the width the mean row length suggests
        spmvb200_matrix* ch = nullptr;
        if (big && !build_child(m, CAND_XWIN, false, &ch)) {
            ch->xw_mode = xwin_mode_rule(ch);
            m->xw_child = ch;
            m->tuned = CAND_XWIN;
        } else if (big && !build_child(m, CAND_SELL, false, &ch)) {
            m->xw_child = ch;
            m->tuned = CAND_SELL;
        } else {
            m->tuned = cand_of_lanes(m->vec_lanes);
        }
        return 0;
    }
    // d_x / d_y are the caller's vectors: y is overwritten by every candidate with the same result
    int best = 0;
    float best_ms = 1e30f;
    const double mean = m->M ? (double) m->NZ / (double) m->M : 0.0;
    const char* force = getenv("SPMVB200_FORCE_CAND");  // developer knob
    for (int c = 0; c < CAND_XWIN; ++c) {
        if (c >= 2 && (cand_lanes(c) > 4 * mean + 2 || (double) cand_lanes(c) * 64 < mean)) continue;  // hopeless widths
        if (force && atoi(force) != c) continue;
        float ms = 0;
        if (time_best_of_2([&] { launch_candidate(m, c, d_x, d_y, st); }, st, &ms)) return 1;
        m->tuned_ms[c] = ms;
        if (ms < best_ms) { best_ms = ms; best = c; }
    }
    // x-window copy: kept only if it beats everything else by 5 %
    if (big && (!force || atoi(force) == CAND_XWIN)) {
        ScopedChild xw;
        if (!build_child(m, CAND_XWIN, false, &xw.m)) {
            float ms = -1.f;
            if (!tune_xwin(xw.m, d_x, d_y, st, &ms)) {
                m->tuned_ms[CAND_XWIN] = ms;
                if (ms < 0.95f * best_ms || force) { best_ms = ms; best = CAND_XWIN; m->xw_child = xw.release(); }
            }
        }
    }
    // SELL-32-sigma copy (a thread walks its whole row, coalesced, no shuffles)
    if (big && (!force || atoi(force) == CAND_SELL)) {
        ScopedChild sell;
        if (!build_child(m, CAND_SELL, false, &sell.m)) {
            const bool hybrid = m->lmax > (uint32_t) VEC_MID;
            float ms = 1e30f;
            if (time_best_of_2([&] {
                    if (hybrid) tail_fork(m, d_x, d_y, st);
                    launch_sell(sell.m, d_x, d_y, st);
                    if (hybrid) tail_join(m, st);
                }, st, &ms)) return 1;
            m->tuned_ms[CAND_SELL] = ms;
            if (ms < 0.95f * best_ms || force) {
                if (m->xw_child) spmvb200_free(m->xw_child);
                m->xw_child = sell.release();
                best_ms = ms;
                best = CAND_SELL;
            }
        }
    }
    // hot-x hybrid: only where the column popularity is skewed enough for a 64-128 KB cache of x to matter; both launch shapes are timed
    // Measured on R-MAT scale 22 (profiles/r02f_hotx_cfg3.log): 0.357 ms against 0.252 ms for the SELL hybrid -- 25 % fewer L2 sectors, but
    // one persistent kernel with 32-48 warps per SM is latency-bound where three concurrent kernels at full occupancy are not.  So the
    // candidate is only built and timed on request (SPMVB200_HOTX=1, or forced); DESIGN.md 4.11.
    if (big && (getenv("SPMVB200_HOTX") || (force && atoi(force) == CAND_HOTX)) && (!force || atoi(force) == CAND_HOTX)) {
        float shape_ms[2] = {1e30f, 1e30f};
        const char* only = getenv("SPMVB200_HOTX_SHAPE");
        for (int shape = 0; shape < 2; ++shape) {
            if (only && atoi(only) != shape) continue;
            if (hotx_build_quiet(m, shape)) break;  // not skewed enough: neither shape qualifies
            if (time_best_of_2([&] { launch_hotx(m, d_x, d_y, st); }, st, &shape_ms[shape])) return 1;
            hotx_drop(m);
        }
        const int shape = shape_ms[1] < shape_ms[0] ? 1 : 0;
        const float ms = shape_ms[shape];
        if (ms < 1e30f) {
            m->tuned_ms[CAND_HOTX] = ms;
            if (getenv("SPMVB200_VERBOSE")) fprintf(stderr, "spmv_b200: hot-x shapes: 16384 x 1 CTA %.3f ms, 8192 x 3 CTAs %.3f ms\n", shape_ms[0], shape_ms[1]);
            if ((ms < 0.95f * best_ms || force) && !hotx_build_quiet(m, shape)) {
                if (m->xw_child) { spmvb200_free(m->xw_child); m->xw_child = nullptr; }
                best_ms = ms;
                best = CAND_HOTX;
            }
        }
    }
    m->tuned = best;
    if (getenv("SPMVB200_VERBOSE")) {
        fprintf(stderr, "spmv_b200: adaptive tuning M=%llu NZ=%llu ->", (unsigned long long) m->M, (unsigned long long) m->NZ);
        for (int c = 0; c < N_CAND; ++c) fprintf(stderr, " %s=%.3fms%s", CAND_NAME[c], m->tuned_ms[c], c == best ? "*" : "");
        fprintf(stderr, "\n");
    }
    return 0;
}

// ---- SPMVB200_CSR_ROWS, the kind that must reproduce sgemvSerial bit for bit: the stream kernel, or -- picked at first use, for
// matrices of at least 2^20 non-zeros -- an x-window copy (column-sorted rows only) or a SELL copy (no row longer than VEC_MID);
// all three add a row's products left to right with separate mul / add roundings.
static void launch_exact_sell(spmvb200_matrix* m, const spmvb200_matrix* sell, const double* x, double* y, cudaStream_t st) {
    const bool hybrid = m->lmax > (uint32_t) VEC_MID;  // rows the capped SELL copy left out (it does not write their y)
    if (hybrid) tail_fork(m, x, y, st, true);
    launch_sell(sell, x, y, st);
    if (hybrid) tail_join(m, st);
}
static int tune_exact(spmvb200_matrix* m, const double* d_x, double* d_y, cudaStream_t st) {
    const bool big = m->NZ >= (1u << 20) && !getenv("SPMVB200_EXACT_ONLY_STREAM");  // developer knob
    m->tuned_x_ms[0] = m->tuned_x_ms[1] = m->tuned_x_ms[2] = -1.f;
    if (tune_mode() == 1) {
        spmvb200_matrix* ch = nullptr;
        if (big && !build_child(m, CAND_XWIN, true, &ch)) {
            ch->xw_mode = xwin_mode_rule(ch);
            m->x_child = ch;
            m->tuned_x = CAND_XWIN;
        } else if (big && !build_child(m, CAND_SELL, true, &ch)) {
            m->x_child = ch;
            m->tuned_x = CAND_SELL;
        } else {
            m->tuned_x = 0;
        }
        return 0;
    }
    float best_ms = 0;
    if (time_best_of_2([&] { launch_csr_stream<false, 0>(m, d_x, d_y, st, 0, m->ntiles); }, st, &best_ms)) return 1;
    int best = 0;
    m->tuned_x_ms[0] = best_ms;
    if (big) {
        ScopedChild xw;
        if (!build_child(m, CAND_XWIN, true, &xw.m)) {
            float ms = 1e30f;
            if (!tune_xwin(xw.m, d_x, d_y, st, &ms)) m->tuned_x_ms[1] = ms;
            if (ms < 0.95f * best_ms) { best_ms = ms; best = CAND_XWIN; m->x_child = xw.release(); }
        }
        // rows longer than VEC_MID are left out of the SELL copy: a warp each adds them in the serial order next to it (exact hybrid);
        // rows longer than a tile are walked by one CTA with the running sum carried through (csr_longrow_exact_kernel)
        ScopedChild sell;
        if (!build_child(m, CAND_SELL, true, &sell.m)) {
            float ms = 1e30f;
            if (!time_best_of_2([&] { launch_exact_sell(m, sell.m, d_x, d_y, st); }, st, &ms)) m->tuned_x_ms[2] = ms;
            const char* f = getenv("SPMVB200_FORCE_EXACT");  // developer knob (tests): 13 = keep the SELL copy whatever the timing says
            if (ms < 0.95f * best_ms || (f && atoi(f) == CAND_SELL && ms < 1e30f)) {
                if (m->x_child) spmvb200_free(m->x_child);
                best_ms = ms; best = CAND_SELL; m->x_child = sell.release();
            }
        }
    }
    m->tuned_x = best;
    if (getenv("SPMVB200_VERBOSE"))
        fprintf(stderr, "spmv_b200: exact-kind tuning M=%llu NZ=%llu -> stream=%.3fms xwindow=%.3fms sell=%.3fms, picked %s\n", (unsigned long long) m->M,
                (unsigned long long) m->NZ, m->tuned_x_ms[0], m->tuned_x_ms[1], m->tuned_x_ms[2], best == 0 ? "stream" : best == CAND_XWIN ? "xwindow" : "sell");
    return 0;
}

// ---- SPMVB200_ELL_ROWS on a padded matrix: the column-major kernel exits early per WARP (a warp runs to the longest of its rows), so one
// long row among 64 short ones keeps the warp's slot busy for K dependent round trips.  When the ELL rectangle is at least 1.25 x the
// non-zeros, a SELL-32-sigma copy (rows sorted by length inside windows: warps see equal lengths) is built from the ELL arrays on the
// device and timed against the column-major kernel at first use; it is kept only if it wins by 5 % (deterministic mode: always kept --
// it won every padded case of the cfg5 sweep).  Both add a row's products left to right with separate mul / add roundings:
// bit-identical results either way.  tuned_x: 0 = column-major ELL, CAND_SELL = the copy.
static int tune_ell(spmvb200_matrix* m, const double* d_x, double* d_y, cudaStream_t st) {
    m->tuned_x = 0;
    if (getenv("SPMVB200_NO_SELL") || getenv("SPMVB200_ELL_NO_SELL") || getenv("SPMVB200_ELL_NO_EARLY_EXIT") || m->NZ < (1u << 20) || !m->rl) return 0;
    if ((double) m->K * (double) m->M < 1.25 * (double) m->NZ) return 0;
    m->tuned_x_ms[0] = m->tuned_x_ms[1] = m->tuned_x_ms[2] = -1.f;
    ScopedChild sell;
    const int quiet0 = g_quiet;
    g_quiet = 1;
    const int rc = sell_build(m, 0, 0xffffffffu, &sell.m);
    g_quiet = quiet0;
    g_err[0] = 0;
    if (rc) return 0;
    if (tune_mode() == 1) {
        m->tuned_x = CAND_SELL;
        m->x_child = sell.release();
        return 0;
    }
    float ell_ms = 0, ms = 1e30f;
    if (time_best_of_2([&] { launch_ell_colmajor(m, d_x, d_y, st, 0, m->M); }, st, &ell_ms)) return 1;
    m->tuned_x_ms[0] = ell_ms;
    if (!time_best_of_2([&] { launch_sell(sell.m, d_x, d_y, st); }, st, &ms)) m->tuned_x_ms[2] = ms;
    if (ms < 0.95f * ell_ms) { m->tuned_x = CAND_SELL; m->x_child = sell.release(); }
    if (getenv("SPMVB200_VERBOSE"))
        fprintf(stderr, "spmv_b200: ELL tuning M=%llu K=%llu NZ=%llu -> ell=%.3fms sell=%.3fms, picked %s\n", (unsigned long long) m->M,
                (unsigned long long) m->K, (unsigned long long) m->NZ, m->tuned_x_ms[0], m->tuned_x_ms[2], m->tuned_x ? "sell" : "ell");
    return 0;
}

static int tune_warp(spmvb200_matrix* m, const double* d_x, double* d_y, cudaStream_t st) {
    // first use: the sub-warp width guessed from the mean row length against its two neighbours (the x gather pattern
    // decides, not the mean: 27-point stencil, mean 26.6 -> guess 16 lanes, 0.203 ms; 4 lanes: 0.141 ms)
    m->vec_tuned = 1;
    if (tune_mode() == 1 || getenv("SPMVB200_VEC_LANES") || m->NZ < (1u << 18)) return 0;
    int best = m->vec_lanes;
    float best_ms = 1e30f;
    const int guess = m->vec_lanes;
    for (int lanes = 2; lanes <= 32; lanes *= 2) {
        if (lanes > 4 * guess || 4 * lanes < guess) continue;
        float ms = 0;
        if (time_best_of_2([&] { launch_csr_vector(m, lanes, d_x, d_y, st, 0, m->M); }, st, &ms)) return 1;
        if (ms < best_ms) { best_ms = ms; best = lanes; }
    }
    m->vec_lanes = best;
    return 0;
}

template <int LANES>
static void launch_ell_rowmajor(const spmvb200_matrix* m, const double* x, double* y, cudaStream_t st) {
    constexpr int BLOCK = 256;
    const uint64_t threads = m->M * LANES;
    ell_rowmajor_kernel<LANES, BLOCK><<<(unsigned) ((threads + BLOCK - 1) / BLOCK), BLOCK, 0, st>>>(m->as, m->ja, m->rl, m->pitch, (uint32_t) m->M,
                                                                                                       (uint32_t) m->K, x, y);
    ++g_launches;
}

// does (handle, kind) still have its first-use pick ahead of it?
static bool needs_tuning(const spmvb200_matrix* m, int kind) {
    switch (kind) {
        case SPMVB200_CSR_ROWS:
        case SPMVB200_ELL_ROWS: return m->tuned_x < 0;
        case SPMVB200_CSR_ADAPTIVE: return m->tuned < 0;
        case SPMVB200_CSR_ROWS_WARP: return !m->vec_tuned && m->format == SPMVB200_FMT_CSR;
        case SPMVB200_XWIN_ROWS: return m->xw_mode < 0;
        default: return false;
    }
}
// Picking allocates, synchronises and (timed mode) reads events: none of that is legal on a capturing stream.
static bool stream_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {  // e.g. the legacy stream while another stream captures in global mode
        cudaGetLastError();
        return true;
    }
    return cs != cudaStreamCaptureStatusNone;
}

// L2 PERSISTENCE FOR x (SPMVB200_X_PERSIST=1, or spmvb200_set_x_persistence): the x vector of the launch is declared a persisting
// access-policy window on the launch stream (everything else the stream touches -- the matrix -- is "streaming"), with the device's
// persisting L2 carve-out at its maximum.  ncu shows why one would want it: on cfg5 (uniform random columns, x = 64 MB < L2) the SELL
// kernel reads 1.63 GB from DRAM for 0.93 GB of matrix -- the 0.83 GB matrix stream keeps evicting x, which is then re-fetched ~11 times.
static int g_x_persist = -1;
static int x_persist_on() {
    if (g_x_persist < 0) g_x_persist = getenv("SPMVB200_X_PERSIST") ? atoi(getenv("SPMVB200_X_PERSIST")) : 0;
    return g_x_persist;
}
static void x_window(cudaStream_t st, const double* d_x, uint64_t n) {
    static thread_local const void* last_ptr = nullptr;
    static thread_local cudaStream_t last_st = nullptr;
    static thread_local uint64_t last_n = 0;
    static bool limit_set[64] = {false};
    if (last_ptr == d_x && last_st == st && last_n == n) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return; }
    if (!limit_set[dev & 63]) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t) p.persistingL2CacheMaxSize);
        limit_set[dev & 63] = true;
    }
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    const size_t bytes = (size_t) std::min<uint64_t>(n * 8, (uint64_t) p.accessPolicyMaxWindowSize);
    a.accessPolicyWindow.base_ptr = const_cast<double*>(d_x);
    a.accessPolicyWindow.num_bytes = bytes;
    a.accessPolicyWindow.hitRatio = (float) std::min(1.0, (double) p.persistingL2CacheMaxSize / (double) std::max<size_t>(bytes, 1));
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) cudaGetLastError();
    last_ptr = d_x;
    last_st = st;
    last_n = n;
}

static int launch(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, cudaStream_t st, LaunchCtx* lc = nullptr) {
    if (!spmvb200_kind_supported(m, kind)) return fail("kind %d (%s) cannot run on format %d", kind, spmvb200_kind_name(kind), m->format);
    if (m->M == 0) return 0;
    if (x_persist_on() && !stream_capturing(st)) x_window(st, d_x, m->N);
    // A handle whose pick is still ahead of it, launched into a stream capture: run the kind's plain kernel (no allocation, no
    // synchronisation, the handle stays untuned) -- call spmvb200_tune before capturing to get the tuned kernel into the graph.
    const bool plain = needs_tuning(m, kind) && stream_capturing(st);
    switch (kind) {
        case SPMVB200_CSR_ROWS:
            if (!plain && m->tuned_x < 0 && tune_exact(m, d_x, d_y, st)) return 1;
            if (!plain && m->tuned_x == CAND_XWIN) { if (launch_xwin(m->x_child, d_x, d_y, st, lc)) return 1; }
            else if (!plain && m->tuned_x == CAND_SELL) launch_exact_sell(m, m->x_child, d_x, d_y, st);
            else launch_csr_stream<false, 0>(m, d_x, d_y, st, 0, m->ntiles);
            break;
        case SPMVB200_CSR_ADAPTIVE:
            if (!plain && m->tuned < 0 && tune_adaptive(m, d_x, d_y, st)) return 1;
            if (plain) launch_csr_vector(m, m->vec_lanes, d_x, d_y, st, 0, m->M);
            else launch_candidate(m, m->tuned, d_x, d_y, st, lc);
            break;
        case SPMVB200_CSR_ROWS_WARP:
            if (!plain && !m->vec_tuned && m->format == SPMVB200_FMT_CSR && tune_warp(m, d_x, d_y, st)) return 1;
            launch_csr_vector(m, m->vec_lanes, d_x, d_y, st, 0, m->M);
            break;
        case SPMVB200_ELL_ROWS:
            if (!plain && m->tuned_x < 0 && tune_ell(m, d_x, d_y, st)) return 1;
            if (!plain && m->tuned_x == CAND_SELL) launch_sell(m->x_child, d_x, d_y, st);
            else launch_ell_colmajor(m, d_x, d_y, st, 0, m->M, lc);
            break;
        case SPMVB200_SELL_ROWS:
            // skewed matrix: the rows the slices leave out run as a compact CSR handle on the serial-order per-row kernels next to them
            if (m->tail) tail_fork(m->tail, d_x, m->tail_y, st, true);
            launch_sell(m, d_x, d_y, st);
            if (m->tail) {
                tail_join(m->tail, st);
                tail_scatter_kernel<<<(unsigned) ((m->tail->M + 255) / 256), 256, 0, st>>>(m->tail_y, m->tail_map, (uint32_t) m->tail->M, d_y);
                ++g_launches;
            }
            break;
        case SPMVB200_XWIN_ROWS:
            if (!plain && m->xw_mode < 0 && tune_xwin(m, d_x, d_y, st, nullptr)) return 1;
            if (launch_xwin(m, d_x, d_y, st, lc)) return 1;
            break;
        case SPMVB200_ELL_ROWS_NT:
            switch (m->vec_lanes) {
                case 1: launch_ell_rowmajor<1>(m, d_x, d_y, st); break;
                case 2: launch_ell_rowmajor<2>(m, d_x, d_y, st); break;
                case 4: launch_ell_rowmajor<4>(m, d_x, d_y, st); break;
                case 8: launch_ell_rowmajor<8>(m, d_x, d_y, st); break;
                case 16: launch_ell_rowmajor<16>(m, d_x, d_y, st); break;
                default: launch_ell_rowmajor<32>(m, d_x, d_y, st); break;
            }
            break;
        case SPMVB200_ELL_ROWS_WARP_NT: {  // warp per row, four rows of a warp in flight
            static const bool plain_warp = getenv("SPMVB200_ELL_WARP_PLAIN") != nullptr;  // developer knob: one row per warp at a time
            if (plain_warp) { launch_ell_rowmajor<32>(m, d_x, d_y, st); break; }
            constexpr int BLOCK = 256, ROWS = 4;
            const uint64_t warps = (m->M + ROWS - 1) / ROWS;
            ell_rowmajor_warp_kernel<ROWS, BLOCK><<<(unsigned) ((warps * 32 + BLOCK - 1) / BLOCK), BLOCK, 0, st>>>(m->as, m->ja, m->rl, m->pitch, (uint32_t) m->M,
                                                                                                               (uint32_t) m->K, d_x, d_y);
            ++g_launches;
            break;
        }
    }
    CU_TRY(cudaPeekAtLastError());
    return 0;
}

static int prefer_smem_once() {
    static bool done_dev[64] = {false};  // per device: function attributes belong to the device's context
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    bool& done = done_dev[dev & 63];
    if (done) return 0;
    // streaming variant: 8 CTAs x 27 KB of shared memory per SM => largest carve-out.
    // gather variant (VARIANT=1): half the shared memory, the rest stays L1 for the x gathers.
    int carve1 = 50;
    if (const char* e = getenv("SPMVB200_CARVEOUT")) carve1 = atoi(e);  // developer knob
    CU_TRY(cudaFuncSetAttribute(csr_stream_kernel<STREAM_TILE, STREAM_BLOCK, STREAM_TILE_ROWS, false, 0>,
                                cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU_TRY(cudaFuncSetAttribute(csr_stream_kernel<STREAM_TILE, STREAM_BLOCK, STREAM_TILE_ROWS, true, 0>,
                                cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU_TRY(cudaFuncSetAttribute(csr_stream_kernel<STREAM_TILE, STREAM_BLOCK, STREAM_TILE_ROWS, true, 1>,
                                cudaFuncAttributePreferredSharedMemoryCarveout, carve1));
    done = true;
    return 0;
}

static int ensure_events(spmvb200_matrix* m) {
    if (!m->ev0) CU_TRY(cudaEventCreate(&m->ev0));
    if (!m->ev1) CU_TRY(cudaEventCreate(&m->ev1));
    return 0;
}
