// engine_multigpu.inl -- part of engine.cu (textually included there: one translation unit, file-local helpers stay static).
// multi-GPU: fused output delivery (peer stores), async copies, CUDA IPC, cross-GPU barrier, CUDA-graph iteration.

extern "C" int spmvb200_spmv_device_push(spmvb200_matrix* m, int kind, const double* d_x, double* d_y, const spmvb200_push* push,
                                         void* stream) {
    if (!m || !d_x || !d_y || !push) return fail("spmv_device_push: null argument");
    if (push->n < 0 || push->n > 8) return fail("spmv_device_push: %d destinations (at most 8)", push->n);
    if (prefer_smem_once()) return 1;
    // first use of a self-tuning kind: tune without deliveries (the tuning run launches every candidate)
    if (needs_tuning(m, kind) && !stream_capturing((cudaStream_t) stream))
        if (launch(m, kind, d_x, d_y, (cudaStream_t) stream)) return 1;
    LaunchCtx lc;
    PushArgs& a = lc.push;
    a.n = push->n;
    for (int i = 0; i < push->n; ++i) {
        if (!push->dst[i] || push->hi[i] > 0xffffffffull || push->lo[i] > push->hi[i]) return fail("spmv_device_push: bad destination %d", i);
        a.dst[i] = push->dst[i];
        a.lo[i] = (uint32_t) push->lo[i];
        a.hi[i] = (uint32_t) push->hi[i];
    }
    if (push->row_offset + m->M > 0xffffffffull) return fail("spmv_device_push: row offset too large");
    a.row_offset = (uint32_t) push->row_offset;
    const int rc = launch(m, kind, d_x, d_y, (cudaStream_t) stream, &lc);
    const bool fused = lc.fused;
    if (rc) return 1;
    if (!fused && a.n && m->M) {  // kernels without the fused epilogue: one more pass over y
        push_rows_kernel<<<592, 256, 0, (cudaStream_t) stream>>>(d_y, (uint32_t) m->M, a);
        ++g_launches;
        CU_TRY(cudaPeekAtLastError());
    }
    return 0;
}

extern "C" int spmvb200_h2d_async(void* d, const void* h, size_t bytes, void* stream) {
    CU_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream));
    return 0;
}
extern "C" int spmvb200_d2h_async(void* h, const void* d, size_t bytes, void* stream) {
    CU_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
    return 0;
}
extern "C" int spmvb200_stream_sync(void* stream) {
    CU_TRY(cudaStreamSynchronize((cudaStream_t) stream));
    return 0;
}
// deliver rows that already sit in device memory (e.g. this GPU's freshly uploaded slice of x) to the destinations that want them
extern "C" int spmvb200_push_rows(const double* d_rows, uint64_t nrows, const spmvb200_push* push, void* stream) {
    if (!d_rows || !push || push->n < 0 || push->n > 8 || nrows > 0xffffffffull) return fail("push_rows: bad arguments");
    PushArgs a = {};
    a.n = push->n;
    for (int i = 0; i < push->n; ++i) {
        if (!push->dst[i] || push->hi[i] > 0xffffffffull || push->lo[i] > push->hi[i]) return fail("push_rows: bad destination %d", i);
        a.dst[i] = push->dst[i];
        a.lo[i] = (uint32_t) push->lo[i];
        a.hi[i] = (uint32_t) push->hi[i];
    }
    a.row_offset = (uint32_t) push->row_offset;
    if (a.n && nrows) {
        push_rows_kernel<<<592, 256, 0, (cudaStream_t) stream>>>(d_rows, (uint32_t) nrows, a);
        ++g_launches;
        CU_TRY(cudaPeekAtLastError());
    }
    return 0;
}

// ---- peer memory plumbing for one-process-per-GPU jobs (CUDA IPC) and the cross-GPU barrier
extern "C" int spmvb200_ipc_export(void* d_ptr, unsigned char handle[64]) {
    if (!d_ptr || !handle) return fail("ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return 0;
}
extern "C" int spmvb200_ipc_open(const unsigned char handle[64], void** d_ptr) {
    if (!d_ptr || !handle) return fail("ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int spmvb200_ipc_close(void* d_ptr) {
    CU_TRY(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}
extern "C" int spmvb200_peer_barrier(uint32_t* const* d_flags, int n, int rank, uint32_t epoch, void* stream) {
    if (!d_flags || n < 1 || n > 8 || rank < 0 || rank >= n) return fail("peer_barrier: bad arguments");
    BarrierArgs b = {};
    for (int i = 0; i < n; ++i) {
        if (!d_flags[i]) return fail("peer_barrier: null flag array %d", i);
        b.flags[i] = d_flags[i];
    }
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t) stream>>>(b, n, rank, epoch);
    CU_TRY(cudaPeekAtLastError());
    return 0;
}

// ---- iterated SpMV on one GPU: x <- A x, `iters` times, ping-pong between two vectors; the launch pair is captured in a CUDA
// graph so that back-to-back SpMVs are not separated by launch latency
extern "C" int spmvb200_iterate_device(spmvb200_matrix* m, int kind, double* d_a, double* d_b, int iters, int use_graph, void* stream,
                                       float* total_ms) {
    if (!m || !d_a || !d_b || iters < 0) return fail("iterate_device: bad arguments");
    if (m->M != m->N) return fail("iterate_device: matrix is %llu x %llu, iteration needs a square matrix", (unsigned long long) m->M, (unsigned long long) m->N);
    if (!spmvb200_kind_supported(m, kind)) return fail("kind %d (%s) cannot run on format %d", kind, spmvb200_kind_name(kind), m->format);
    if (prefer_smem_once() || ensure_events(m)) return 1;
    cudaStream_t st = (cudaStream_t) stream;
    cudaStream_t own = nullptr;
    if (use_graph && (st == nullptr || st == cudaStreamLegacy)) {  // the legacy default stream cannot be captured
        CU_TRY(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
        CU_TRY(cudaDeviceSynchronize());
        st = own;
    }
    int rc = 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    do {
        // self-tuning kinds tune here (cannot happen inside a capture); d_b is scratch at this point
        if (needs_tuning(m, kind))
            if ((rc = launch(m, kind, d_a, d_b, st))) break;
        const int pairs = iters / 2;
        unsigned long long per_replay = 0;
        if (use_graph && pairs > 0) {
            const unsigned long long l0 = g_launches.load();
            if ((rc = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess)) break;
            int r1 = launch(m, kind, d_a, d_b, st);
            int r2 = r1 ? 1 : launch(m, kind, d_b, d_a, st);
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (r1 || r2 || ce != cudaSuccess) { rc = 1; if (ce != cudaSuccess) fail("iterate_device: capture failed: %s", cudaGetErrorString(ce)); break; }
            if ((rc = cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)) break;
            per_replay = g_launches.load() - l0;  // kernels inside the graph: counted per replay below
            g_launches -= per_replay;
        }
        if ((rc = cudaEventRecord(m->ev0, st) != cudaSuccess)) break;
        for (int i = 0; i < pairs && !rc; ++i) {
            if (exec) { rc = cudaGraphLaunch(exec, st) != cudaSuccess; g_launches += per_replay; }
            else rc = launch(m, kind, d_a, d_b, st) || launch(m, kind, d_b, d_a, st);
        }
        if (!rc && (iters & 1)) rc = launch(m, kind, d_a, d_b, st);
        if (rc) break;
        if ((rc = cudaEventRecord(m->ev1, st) != cudaSuccess)) break;
        if ((rc = cudaEventSynchronize(m->ev1) != cudaSuccess)) break;
        if (total_ms && (rc = cudaEventElapsedTime(total_ms, m->ev0, m->ev1) != cudaSuccess)) break;
    } while (0);
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (own) cudaStreamDestroy(own);
    if (rc) {
        if (!g_err[0] || cudaPeekAtLastError() != cudaSuccess) fail("iterate_device: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}
