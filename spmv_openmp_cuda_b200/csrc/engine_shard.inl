// engine_shard.inl -- part of engine.cu (textually included there: one translation unit, file-local helpers stay static).
// One GPU's share of a ROW-BLOCK PARTITIONED SpMV over the GPUs of one box, one process per GPU (the reference's own
// CPU decomposition, spmvRowsBlocksCSR, src/SpMV_CSR_OMP.c:65-99 with the block arithmetic of src/include/macros.h:33-36, one block
// per GPU).  The whole step runs inside the library -- no host-language code between the copies, the exchange and the kernel:
//
//   spmvb200_shard_step      : device-resident x <- A x.  The SpMV kernel stores each finished row into the next x of every GPU
//                              whose columns reference it (posted NVLink stores from the kernel's epilogue: a halo for banded
//                              matrices, the whole block for unstructured ones), then a one-block flag barrier.  No collective
//                              library call in the data path.
//   spmvb200_shard_spmv_host : host x slice -> device over this GPU's own PCIe link (rows the peers read go up first and are
//                              delivered to them by peer stores + barrier while the rest is still uploading), row chunks launch as
//                              their x pieces land, y chunks go down while later chunks compute.
//
// The rendezvous (exchanging the 64-byte CUDA IPC handles) is the caller's: spmvb200_shard_export gives this rank's blob,
// spmvb200_shard_connect takes everybody's (torch.distributed all_gather_object in bench.py / distributed.py -- plumbing only).

struct spmvb200_shard {
    spmvb200_matrix* m = nullptr;
    int kind = 0, rank = 0, world = 1, nbuf = 2;
    uint64_t r0 = 0, r1 = 0, N = 0;
    std::vector<uint64_t> splits;
    double* x[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* flags = nullptr;
    double* peer_x[8][4] = {};
    uint32_t* peer_flags[8] = {};
    uint64_t col_lo = 1, col_hi = 0;  // columns this rank's rows reference (lo > hi: none)
    int npeer = 0;                    // peers that read rows of mine
    int peer_id[8] = {};
    uint64_t need_lo[8] = {}, need_hi[8] = {};
    std::vector<std::pair<uint64_t, uint64_t>> halo;  // merged need ranges (global rows), uploaded first by the host path
    bool halo_first = false;
    uint32_t epoch = 0;
    uint32_t* ticket = nullptr;  // finished-CTA counter of the fused neighbour synchronisation
    int nsync = 0;               // neighbours: ranks I deliver rows to or receive rows from
    int sync_rank[8] = {};
    int nb_mode = -2;            // x-window launch shape the boundary CTAs were counted for (-2: not counted yet)
    uint32_t nboundary = 0;
    uint8_t* cta_boundary = nullptr;  // [grid of that launch shape] 1 = boundary CTA
    uint32_t rb_rot = 0;              // rotation of the row-block order (PushArgs::rb_rot)
    double* d_y = nullptr;
    bool connected = false;
    int host_cur = 0;
    cudaEvent_t e_halo = nullptr;
};

static const size_t SHARD_BLOB_FIXED = 16;  // col_lo, col_hi
extern "C" size_t spmvb200_shard_blob_bytes(const spmvb200_shard* s) { return s ? (size_t) (s->nbuf + 1) * 64 + SHARD_BLOB_FIXED : 0; }

extern "C" int spmvb200_shard_free(spmvb200_shard* s) {
    if (!s) return 0;
    cudaDeviceSynchronize();
    for (int p = 0; p < s->world; ++p) {
        if (p == s->rank) continue;
        for (int b = 0; b < s->nbuf; ++b)
            if (s->peer_x[p][b]) cudaIpcCloseMemHandle(s->peer_x[p][b]);
        if (s->peer_flags[p]) cudaIpcCloseMemHandle(s->peer_flags[p]);
    }
    for (int b = 0; b < 4; ++b) cudaFree(s->x[b]);
    cudaFree(s->flags);
    cudaFree(s->ticket);
    cudaFree(s->cta_boundary);
    cudaFree(s->d_y);
    if (s->e_halo) cudaEventDestroy(s->e_halo);
    delete s;
    cudaGetLastError();
    return 0;
}

extern "C" int spmvb200_shard_create(spmvb200_matrix* m, int kind, int rank, int world, const uint64_t* splits, int nbuf,
                                     const uint64_t* col_range, spmvb200_shard** out) {
    if (!out) return fail("shard_create: null output");
    *out = nullptr;
    if (!m || !splits || world < 1 || world > 8 || rank < 0 || rank >= world || nbuf < 2 || nbuf > 4) return fail("shard_create: bad arguments");
    if (!spmvb200_kind_supported(m, kind)) return fail("shard_create: kind %d (%s) cannot run on format %d", kind, spmvb200_kind_name(kind), m->format);
    for (int g = 0; g < world; ++g)
        if (splits[g] > splits[g + 1]) return fail("shard_create: split points must be non-decreasing");
    if (splits[0] != 0 || splits[rank + 1] - splits[rank] != m->M) return fail("shard_create: handle has %llu rows, the partition gives rank %d %llu", (unsigned long long) m->M, rank, (unsigned long long) (splits[rank + 1] - splits[rank]));
    if (splits[world] != m->N) return fail("shard_create: x <- A x needs a square matrix (rows %llu, columns %llu)", (unsigned long long) splits[world], (unsigned long long) m->N);
    spmvb200_shard* s = new spmvb200_shard();
    s->m = m;
    s->kind = kind;
    s->rank = rank;
    s->world = world;
    s->nbuf = nbuf;
    s->N = m->N;
    s->splits.assign(splits, splits + world + 1);
    s->r0 = splits[rank];
    s->r1 = splits[rank + 1];
    int rc = 0;
    do {
        for (int b = 0; b < nbuf && !rc; ++b) {
            rc = cudaMalloc(&s->x[b], std::max<uint64_t>(s->N, 2) * 8) != cudaSuccess;
            if (!rc) rc = cudaMemset(s->x[b], 0, std::max<uint64_t>(s->N, 2) * 8) != cudaSuccess;
        }
        if (rc) break;
        if ((rc = cudaMalloc(&s->flags, 64) != cudaSuccess)) break;
        if ((rc = cudaMemset(s->flags, 0, 64) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&s->ticket, 64) != cudaSuccess)) break;
        if ((rc = cudaMemset(s->ticket, 0, 64) != cudaSuccess)) break;
        if ((rc = cudaMalloc(&s->d_y, std::max<uint64_t>(m->M, 2) * 8) != cudaSuccess)) break;
        if ((rc = cudaEventCreateWithFlags(&s->e_halo, cudaEventDisableTiming) != cudaSuccess)) break;
        if (col_range) {
            s->col_lo = col_range[0];
            s->col_hi = col_range[1];
        } else if (m->format == SPMVB200_FMT_CSR && m->NZ) {
            if ((rc = spmvb200_col_range(m, &s->col_lo, &s->col_hi))) break;
        } else if (m->NZ) {  // no cheap way to know: the rows may read all of x
            s->col_lo = 0;
            s->col_hi = s->N - 1;
        }
        if ((rc = cudaDeviceSynchronize() != cudaSuccess)) break;
    } while (0);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("shard_create: %s", cudaGetErrorString(cudaGetLastError()));
        spmvb200_shard_free(s);
        return 1;
    }
    s->peer_flags[rank] = s->flags;
    for (int b = 0; b < nbuf; ++b) s->peer_x[rank][b] = s->x[b];
    if (world == 1) s->connected = true;
    *out = s;
    return 0;
}

extern "C" int spmvb200_shard_export(spmvb200_shard* s, unsigned char* blob) {
    if (!s || !blob) return fail("shard_export: null argument");
    memset(blob, 0, spmvb200_shard_blob_bytes(s));
    if (s->world > 1) {
        for (int b = 0; b <= s->nbuf; ++b) {
            cudaIpcMemHandle_t h;
            CU_TRY(cudaIpcGetMemHandle(&h, b < s->nbuf ? (void*) s->x[b] : (void*) s->flags));
            memcpy(blob + (size_t) b * 64, &h, 64);
        }
    }
    memcpy(blob + (size_t) (s->nbuf + 1) * 64, &s->col_lo, 8);
    memcpy(blob + (size_t) (s->nbuf + 1) * 64 + 8, &s->col_hi, 8);
    return 0;
}

// blobs: world x spmvb200_shard_blob_bytes, in rank order (this rank's own entry is ignored for the mappings)
extern "C" int spmvb200_shard_connect(spmvb200_shard* s, const unsigned char* blobs) {
    if (!s || !blobs) return fail("shard_connect: null argument");
    const size_t bb = spmvb200_shard_blob_bytes(s);
    s->npeer = 0;
    s->nsync = 0;
    s->halo.clear();
    for (int p = 0; p < s->world; ++p) {
        const unsigned char* b = blobs + (size_t) p * bb;
        uint64_t lo, hi;
        memcpy(&lo, b + (size_t) (s->nbuf + 1) * 64, 8);
        memcpy(&hi, b + (size_t) (s->nbuf + 1) * 64 + 8, 8);
        if (p == s->rank) continue;
        if (!s->connected) {
            for (int i = 0; i <= s->nbuf; ++i) {
                cudaIpcMemHandle_t h;
                memcpy(&h, b + (size_t) i * 64, 64);
                void* ptr = nullptr;
                CU_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
                if (i < s->nbuf) s->peer_x[p][i] = (double*) ptr; else s->peer_flags[p] = (uint32_t*) ptr;
            }
        }
        // neighbour = a rank that reads rows of mine, or whose rows I read (my columns reach into its slice)
        const bool p_reads_me = lo <= hi && std::max(s->r0, lo) < std::min(s->r1, hi + 1);
        const bool i_read_p = s->col_lo <= s->col_hi && std::max(s->splits[p], s->col_lo) < std::min(s->splits[p + 1], s->col_hi + 1);
        if (p_reads_me || i_read_p) s->sync_rank[s->nsync++] = p;
        // rows of mine that rank p reads as columns of x: [r0, r1) intersected with its referenced column range
        if (lo > hi) continue;
        const uint64_t a = std::max(s->r0, lo), e = std::min(s->r1, hi + 1);
        if (a < e) {
            s->peer_id[s->npeer] = p;
            s->need_lo[s->npeer] = a;
            s->need_hi[s->npeer] = e;
            ++s->npeer;
            s->halo.emplace_back(a, e);
        }
    }
    std::sort(s->halo.begin(), s->halo.end());
    std::vector<std::pair<uint64_t, uint64_t>> merged;
    uint64_t covered = 0;
    for (auto& h : s->halo) {
        if (!merged.empty() && h.first <= merged.back().second) merged.back().second = std::max(merged.back().second, h.second);
        else merged.push_back(h);
    }
    for (auto& h : merged) covered += h.second - h.first;
    s->halo.swap(merged);
    // uploading the halo ahead of the rest pays only while it is a small part of the slice (banded / stencil matrices)
    s->halo_first = covered * 4 <= (s->r1 - s->r0);
    s->connected = true;
    return 0;
}

extern "C" double* spmvb200_shard_x(spmvb200_shard* s, int buf) { return (s && buf >= 0 && buf < s->nbuf) ? s->x[buf] : nullptr; }
extern "C" int spmvb200_shard_halo_rows(const spmvb200_shard* s, uint64_t* rows) {
    if (!s || !rows) return fail("shard_halo_rows: null argument");
    *rows = 0;
    for (int i = 0; i < s->npeer; ++i) *rows += s->need_hi[i] - s->need_lo[i];
    return 0;
}

static int shard_barrier(spmvb200_shard* s, cudaStream_t st) {
    if (s->world <= 1) return 0;
    BarrierArgs b = {};
    for (int p = 0; p < s->world; ++p) b.flags[p] = s->peer_flags[p];
    peer_barrier_kernel<<<1, 32, 0, st>>>(b, s->world, s->rank, ++s->epoch);
    ++g_launches;
    CU_TRY(cudaPeekAtLastError());
    return 0;
}
// The cross-GPU flag barrier on its own (asynchronous on `stream`): every rank's earlier work on its stream, including stores to
// peer memory, is visible to the others' later work.  Every rank must make the call.
extern "C" int spmvb200_shard_barrier(spmvb200_shard* s, void* stream) {
    if (!s || !s->connected) return fail("shard_barrier: shard not connected");
    return shard_barrier(s, (cudaStream_t) stream);
}
static void shard_push_args(const spmvb200_shard* s, int buf, spmvb200_push* p) {
    memset(p, 0, sizeof(*p));
    p->n = s->npeer;
    for (int i = 0; i < s->npeer; ++i) {
        p->dst[i] = s->peer_x[s->peer_id[i]][buf];
        p->lo[i] = s->need_lo[i];
        p->hi[i] = s->need_hi[i];
    }
    p->row_offset = s->r0;
}

// x[dst][r0..r1) = A_local * x[src]; the rows the peers read are stored into THEIR x[dst] as well.  Asynchronous on `stream`.
// Kernels with a delivery epilogue (x-window, column-major ELL) also carry the step's synchronisation: they wait for their neighbours'
// previous step before the first access and their last CTA publishes this step's completion (PushArgs, common.cuh) -- one launch per
// step.  Other kinds deliver with one pass over y and pass a flag barrier across all GPUs.  Epochs advance identically on every rank.
extern "C" int spmvb200_shard_step(spmvb200_shard* s, int src, int dst, void* stream) {
    if (!s || !s->connected) return fail("shard_step: shard not connected");
    if (src < 0 || src >= s->nbuf || dst < 0 || dst >= s->nbuf || src == dst) return fail("shard_step: bad buffers %d -> %d", src, dst);
    spmvb200_matrix* m = s->m;
    cudaStream_t st = (cudaStream_t) stream;
    if (prefer_smem_once()) return 1;
    if (needs_tuning(m, s->kind) && !stream_capturing(st))  // first use: pick without deliveries
        if (launch(m, s->kind, s->x[src], s->x[dst] + s->r0, st)) return 1;
    // Synchronisation of the step: by default a one-block flag-barrier kernel after the SpMV.  SPMVB200_SHARD_FUSED_SYNC=1 moves it into
    // the boundary CTAs of the x-window kernel instead (PushArgs, common.cuh).  Both were measured on cfg4: 2 GPUs
    // 1.227 ms (barrier kernel) vs 1.234-1.239 ms (fused) on a 1.215 ms kernel; 8 GPUs 0.3321 vs 0.3329 ms on 0.320 ms -- no gain from
    // the fused form, so the simpler one is the default.
    static const bool no_fused_sync = getenv("SPMVB200_SHARD_FUSED_SYNC") == nullptr || getenv("SPMVB200_SHARD_BARRIER") != nullptr;
    LaunchCtx lc;
    PushArgs& a = lc.push;
    a.n = s->npeer;
    for (int i = 0; i < s->npeer; ++i) {
        a.dst[i] = s->peer_x[s->peer_id[i]][dst];
        a.lo[i] = (uint32_t) s->need_lo[i];
        a.hi[i] = (uint32_t) s->need_hi[i];
    }
    a.row_offset = (uint32_t) s->r0;
    // the x-window kernel can carry the step's synchronisation in its boundary CTAs
    const spmvb200_matrix* xw = xwin_of(m, s->kind);
    if (xw && s->world > 1 && s->nsync && !no_fused_sync && !stream_capturing(st)) {
        a.nsync = s->nsync;
        a.my_flags = s->flags;
        for (int i = 0; i < s->nsync; ++i) {
            a.peer_rank[i] = (uint8_t) s->sync_rank[i];
            a.peer_cell[i] = s->peer_flags[s->sync_rank[i]] + s->rank;
        }
        a.ticket = s->ticket;
        a.own_lo = (uint32_t) s->r0;
        a.own_hi = (uint32_t) s->r1;
        const int mode = xw->xw_mode == 1 ? 1 : 0;
        if (s->nb_mode != mode) {  // count the boundary CTAs of this launch shape once
            uint32_t* d_cnt = nullptr;
            CU_TRY(cudaMalloc(&d_cnt, 4));
            const uint32_t ncta = mode ? xw->xw_ncta : xw->xw_nrb;
            cudaFree(s->cta_boundary);
            s->cta_boundary = nullptr;
            CU_TRY(cudaMalloc(&s->cta_boundary, std::max<uint32_t>(ncta, 1)));
            static const bool no_rot = getenv("SPMVB200_SHARD_NO_ROTATION") != nullptr;  // developer knob
            s->rb_rot = 0;
            cudaError_t e = cudaSuccess;
            for (int pass = 0; pass < 2 && e == cudaSuccess; ++pass) {
                // pass 0: flags in natural order -> how many boundary CTAs lead the grid; pass 1 (if any do): flags for the rotated order
                a.rb_rot = s->rb_rot;
                CU_TRY(cudaMemsetAsync(d_cnt, 0, 4, st));
                xw_count_boundary_kernel<<<(ncta + 255) / 256, 256, 0, st>>>(mode ? xw->xw_cta_rb : nullptr, xw->xw_rb_tile0, xw->xw_tile_win, ncta, xw->xw_R, xw->xw_W,
                                                                            (uint32_t) xw->M, (uint32_t) xw->N, a, d_cnt, s->cta_boundary);
                e = cudaMemcpyAsync(&s->nboundary, d_cnt, 4, cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (pass == 0 && e == cudaSuccess && mode == 0 && !no_rot && s->nboundary && s->nboundary < ncta) {
                    std::vector<uint8_t> h(ncta);
                    e = cudaMemcpy(h.data(), s->cta_boundary, ncta, cudaMemcpyDeviceToHost);
                    uint32_t lead = 0;
                    while (lead < ncta && h[lead]) ++lead;
                    if (lead == 0 || lead == ncta) break;
                    s->rb_rot = lead;
                } else {
                    break;
                }
            }
            cudaFree(d_cnt);
            if (e != cudaSuccess) return fail("shard_step: boundary count: %s", cudaGetErrorString(e));
            s->nb_mode = mode;
        }
        a.nboundary = s->nboundary;
        a.cta_boundary = s->cta_boundary;
        a.rb_rot = s->rb_rot;
        a.wait_epoch = s->epoch;      // neighbours have finished the previous exchange step
        a.sig_epoch = s->epoch + 1;   // ... and this is what my completion looks like to them
        if (a.nboundary == 0) a.nsync = 0;  // cannot happen with neighbours, but never launch a kernel nobody would signal from
    }
    if (launch(m, s->kind, s->x[src], s->x[dst] + s->r0, st, &lc)) return 1;
    if (lc.fused && a.nsync) {
        ++s->epoch;  // the kernel carried the synchronisation
        return 0;
    }
    if (!lc.fused && a.n && m->M) {  // kinds without the delivery epilogue: one more pass over y
        PushArgs plain = a;
        plain.nsync = 0;
        push_rows_kernel<<<592, 256, 0, st>>>(s->x[dst] + s->r0, (uint32_t) m->M, plain);
        ++g_launches;
        CU_TRY(cudaPeekAtLastError());
    }
    return shard_barrier(s, st);
}

// Host-buffer step: x_slice = this rank's rows [r0, r1) of x (host), y_slice = the same rows of y = A x (host).  Returns when y_slice is
// complete.  Every rank of the job must make the call (it contains the cross-GPU barrier).  *kernel_ms: CUDA-event time of the SpMV.
extern "C" int spmvb200_shard_spmv_host(spmvb200_shard* s, const double* x_slice, double* y_slice, float* kernel_ms) {
    if (!s || !s->connected || !x_slice || !y_slice) return fail("shard_spmv_host: bad arguments / shard not connected");
    spmvb200_matrix* m = s->m;
    const int kind = s->kind;
    if (prefer_smem_once() || ensure_events(m)) return 1;
    const uint64_t rows = s->r1 - s->r0;
    host_buffer_seen(x_slice, rows * 8);
    host_buffer_seen(y_slice, rows * 8);
    if (needs_tuning(m, kind) && kind != SPMVB200_XWIN_ROWS)  // first use: pick on whatever the buffer holds (values do not matter)
        if (launch(m, kind, s->x[s->host_cur], s->d_y, 0) || cudaStreamSynchronize(0) != cudaSuccess) return fail("shard_spmv_host: first-use pick failed");
    const int cand = pipe_candidate(m, kind);
    if (!m->pipe || m->pipe->kind != kind || m->pipe->cand != cand || m->pipe->nch_req != host_chunks_wanted(m))
        if (build_pipe(m, kind, cand)) return 1;
    HostPipe* p = m->pipe;
    const int buf = (s->host_cur ^= 1);  // alternate: a peer may still be reading the other buffer (its SpMV of the previous step)
    double* const xd = s->x[buf];
    auto up = [&](uint64_t a, uint64_t e) -> cudaError_t {  // global rows [a, e) of my slice
        return e > a ? cudaMemcpyAsync(xd + a, x_slice + (a - s->r0), (e - a) * 8, cudaMemcpyHostToDevice, p->s_up) : cudaSuccess;
    };
    spmvb200_push push;
    shard_push_args(s, buf, &push);
    // 1. the rows the peers read go up first, are delivered by peer stores, and the barrier is passed while the rest uploads
    if (s->halo_first) {
        for (auto& h : s->halo) CU_TRY(up(h.first, h.second));
        CU_TRY(cudaEventRecord(s->e_halo, p->s_up));
        CU_TRY(cudaStreamWaitEvent(p->s_comp, s->e_halo, 0));
        for (auto& h : s->halo) {
            spmvb200_push sub = push;
            sub.row_offset = h.first;
            if (spmvb200_push_rows(xd + h.first, h.second - h.first, &sub, p->s_comp)) return 1;
        }
        if (shard_barrier(s, p->s_comp)) return 1;
    }
    // 2. x pieces of the pipeline plan, clipped to my slice (what lies outside comes from the peers)
    for (int j = 0; j < p->nch; ++j) {
        uint64_t a = std::max(p->x_b[j], s->r0), e = std::min(p->x_b[j + 1], s->r1);
        if (j + 1 == p->nch) e = s->r1;  // everything: rows nobody's kernel reads still belong to the replicated x
        if (j == 0) a = s->r0;
        if (s->halo_first) {  // skip what went up already (halo ranges sit at the edges of the slice: clip against each)
            uint64_t cur = a;
            for (auto& h : s->halo) {
                if (h.second <= cur || h.first >= e) continue;
                CU_TRY(up(cur, std::min(h.first, e)));
                cur = std::max(cur, h.second);
            }
            CU_TRY(up(cur, e));
        } else {
            CU_TRY(up(a, e));
        }
        CU_TRY(cudaEventRecord(p->x_ready[j], p->s_up));
    }
    if (!s->halo_first) {
        CU_TRY(cudaStreamWaitEvent(p->s_comp, p->x_ready[p->nch - 1], 0));
        if (push.n && spmvb200_push_rows(xd + s->r0, rows, &push, p->s_comp)) return 1;
        if (shard_barrier(s, p->s_comp)) return 1;
    }
    // 3. row chunks as their pieces land; y chunks go down while later chunks compute
    for (int k = 0; k < p->nch; ++k) {
        CU_TRY(cudaStreamWaitEvent(p->s_comp, p->x_ready[k], 0));
        if (kernel_ms) CU_TRY(cudaEventRecord(p->k_start[k], p->s_comp));  // time stamps only on request (they slow the pipeline down)
        if (p->nch <= 1 && cand < 0) { if (launch(m, kind, xd, s->d_y, p->s_comp)) return 1; }
        else launch_chunk(m, p, k, xd, s->d_y);
        cudaEvent_t const kdone = kernel_ms ? p->k_end[k] : p->k_done[k];
        CU_TRY(cudaEventRecord(kdone, p->s_comp));
        const uint64_t a = p->row_b[k], e = p->row_b[k + 1];
        if (e > a) {
            CU_TRY(cudaStreamWaitEvent(p->s_down, kdone, 0));
            CU_TRY(cudaMemcpyAsync(y_slice + a, s->d_y + a, (e - a) * 8, cudaMemcpyDeviceToHost, p->s_down));
        }
    }
    CU_TRY(cudaPeekAtLastError());
    CU_TRY(cudaStreamSynchronize(p->s_comp));
    CU_TRY(cudaStreamSynchronize(p->s_down));
    CU_TRY(cudaStreamSynchronize(p->s_up));
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
