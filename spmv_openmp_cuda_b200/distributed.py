"""Row-block partitioned SpMV over the GPUs of one box: one process per GPU, torch.distributed
(NCCL over NVLink/NVSwitch) for the plumbing.

The reference has no multi-device path; this follows its own CPU decomposition --
contiguous row blocks (spmvRowsBlocksCSR, src/SpMV_CSR_OMP.c:65-99, block arithmetic of
UNIF_REMINDER_DISTRI[_STARTIDX], src/include/macros.h:33-36) -- with one block per GPU:

  * partition: split points by nnz balance, r_g = lower_bound(IRP, g*nnz/G), or by row count;
  * every rank uploads only its rows (column ids stay global) with spMatCpyCSR(mat, r_g, r_{g+1});
  * x is replicated: root copies it host->device, then ONE collective, broadcast(x) -- the only
    exchange step the path has; y needs none: each rank owns a disjoint slice.  `spmv` gathers the
    slices on the root, `spmv_allgather` leaves the full y on every rank (x <- y iterations).

The local compute is always the CUDA engine; `local_spmv` exists so that the CPU test-suite can
drive the partition + collective logic on the gloo backend with an injected checker.
"""
import numpy as np

from . import engine


def row_partition_uniform(M, G):
    """Contiguous blocks, the first M % G get one extra row (UNIF_REMINDER_DISTRI_STARTIDX,
    src/include/macros.h:33-36).  Returns G+1 split points."""
    base, rem = divmod(int(M), int(G))
    return [g * base + min(g, rem) for g in range(G)] + [int(M)]


def row_partition_by_nnz(irp, G):
    """Split points r_0=0 <= ... <= r_G=M with r_g = first row whose start offset reaches g*nnz/G
    (SURVEY.md §8e): per-GPU matrix bytes are balanced even for skewed row lengths."""
    irp = np.asarray(irp)
    M, nnz = len(irp) - 1, int(irp[-1])
    pts = [0]
    for g in range(1, G):
        target = (g * nnz) // G
        r = int(np.searchsorted(irp, target, side="left"))
        pts.append(min(max(r, pts[-1]), M))
    return pts + [M]


class RowBlockSpmv:
    """y = A x with A row-block partitioned over the ranks of a torch.distributed process group."""

    def __init__(self, mat, kind=engine.CSR_ADAPTIVE, balance="nnz", group=None, device=None, local_spmv=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.M, self.N, self.kind = mat.M, mat.N, kind
        self.splits = row_partition_by_nnz(mat.IRP, self.world) if balance == "nnz" else row_partition_uniform(mat.M, self.world)
        self.r0, self.r1 = self.splits[self.rank], self.splits[self.rank + 1]
        self.rows = self.r1 - self.r0
        self.max_rows = max(b - a for a, b in zip(self.splits[:-1], self.splits[1:]))
        self.device = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                                          if torch.cuda.is_available() and local_spmv is None else torch.device("cpu"))
        if local_spmv is None:
            self.dmat = engine.spMatCpyCSR(mat, self.r0, self.r1)

            def local_spmv(x_dev, y_dev):
                stream = torch.cuda.current_stream().cuda_stream
                engine._launch(kind, self.dmat, x_dev, Config_none, y_dev, stream)
        self.local_spmv = local_spmv
        self.x = torch.empty(self.N, dtype=torch.float64, device=self.device)
        # padded to the largest slice so that gather / all_gather see equal shapes
        self.y = torch.zeros(self.max_rows, dtype=torch.float64, device=self.device)

    def _bcast(self, x_root):
        if self.rank == 0:
            self.x.copy_(x_root if self.torch.is_tensor(x_root) else self.torch.from_numpy(np.ascontiguousarray(x_root)), non_blocking=True)
        if self.world > 1:
            self.dist.broadcast(self.x, src=0, group=self.group)

    def spmv(self, x_root=None):
        """x lives on rank 0 (host array / tensor); returns the full y on rank 0 (None elsewhere)."""
        torch, dist = self.torch, self.dist
        self._bcast(x_root)
        self.local_spmv(self.x, self.y[:self.rows] if self.rows else self.y[:0])
        if self.world == 1:
            return self.y[:self.rows].cpu().numpy()
        parts = [torch.empty_like(self.y) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(self.y, parts, dst=0, group=self.group)
        if self.rank != 0:
            return None
        out = np.empty(self.M, dtype=np.float64)
        for g, p in enumerate(parts):
            a, b = self.splits[g], self.splits[g + 1]
            out[a:b] = p[:b - a].cpu().numpy()
        return out

    def spmv_allgather(self, x_dev=None):
        """x already replicated on every rank (device tensor, or the one from the last call);
        returns the full y as a device tensor on every rank (requires M == N to iterate x <- y)."""
        torch, dist = self.torch, self.dist
        if x_dev is not None:
            self.x.copy_(x_dev)
        self.local_spmv(self.x, self.y[:self.rows] if self.rows else self.y[:0])
        if self.world == 1:
            return self.y[:self.rows].clone()
        buf = torch.empty(self.world * self.max_rows, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(buf, self.y, group=self.group)
        out = torch.empty(self.M, dtype=torch.float64, device=self.device)
        for g in range(self.world):
            a, b = self.splits[g], self.splits[g + 1]
            out[a:b] = buf[g * self.max_rows: g * self.max_rows + (b - a)]
        return out


Config_none = engine.Config()


def needed_rows(splits, col_ranges, rank):
    """Which of `rank`'s rows [r0, r1) each other rank reads as columns of x: the intersection of [r0, r1) with that
    rank's referenced column range [cmin, cmax] (None = the rank has no non-zeros).  Returns {peer: (lo, hi)} with lo < hi:
    for a banded matrix this is the halo next to the block boundary, for an unstructured one the whole block."""
    r0, r1 = splits[rank], splits[rank + 1]
    out = {}
    for p, cr in enumerate(col_ranges):
        if p == rank or cr is None:
            continue
        lo, hi = max(r0, int(cr[0])), min(r1, int(cr[1]) + 1)
        if lo < hi:
            out[p] = (lo, hi)
    return out


class RowBlockIterate:
    """x <- A x, repeated, over the GPUs of one box: A is row-block partitioned (one process per GPU,
    `dmat` = this rank's rows [r0, r1) with global column ids), x is replicated.

    mode="push": the exchange is FUSED into the SpMV -- every rank maps the other ranks' x buffers (CUDA IPC) and its kernel
    stores each finished row straight into the next x of the ranks that read it (posted NVLink stores, only the rows a
    peer's columns reference), followed by a one-block flag barrier across the GPUs; no collective library call per iteration.
    mode="nccl": the plain way -- SpMV into the local y slice, then an NCCL all-gather of the slices (equal slices only).
    torch.distributed is used for the one-time rendezvous (handles, column ranges) in both modes."""

    def __init__(self, dmat, splits, kind, group=None, mode="push", halo=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.mode, self.kind, self.dmat = torch, dist, group, mode, kind, dmat
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.splits = [int(v) for v in splits]
        self.r0, self.r1 = self.splits[self.rank], self.splits[self.rank + 1]
        self.N = dmat.N
        assert dmat.M == self.r1 - self.r0 and self.splits[-1] == self.N, "square matrix, row-block partitioned"
        self.epoch = 0
        self.peer = {}
        lib, C = engine.lib(), engine.C
        if mode == "nccl":
            assert len({b - a for a, b in zip(self.splits[:-1], self.splits[1:])}) == 1, "all-gather needs equal slices"
            self.x = [torch.zeros(self.N, dtype=torch.float64, device="cuda") for _ in range(2)]
            self.y = torch.zeros(self.r1 - self.r0, dtype=torch.float64, device="cuda")
            return
        assert self.world <= 8
        self.bufs = [engine.DeviceVector(self.N), engine.DeviceVector(self.N)]
        self.flags = engine.DeviceVector(8)  # 64 bytes, used as uint32 flags
        for v in self.bufs + [self.flags]:
            v.fill_bytes(0)
        mine = []
        for v in self.bufs + [self.flags]:
            h = (C.c_ubyte * 64)()
            engine.check(lib.spmvb200_ipc_export(v.data_ptr(), h), "ipc_export")
            mine.append(bytes(h))
        cr = dmat.col_range if dmat.NZ else None
        allh = [None] * self.world
        dist.all_gather_object(allh, (mine, cr), group=group)
        self.col_ranges = [a[1] for a in allh]
        err = None
        for p in range(self.world):
            if p == self.rank:
                self.peer[p] = [v.data_ptr() for v in self.bufs + [self.flags]]
                continue
            ptrs = []
            for hb in allh[p][0]:
                out = C.c_void_p()
                try:
                    engine.check(lib.spmvb200_ipc_open((C.c_ubyte * 64).from_buffer_copy(hb), C.byref(out)), "ipc_open")
                    ptrs.append(out.value)
                except engine.SpmvB200Error as e:  # keep going: every rank must reach the agreement below
                    err = err or e
            self.peer[p] = ptrs
        # all ranks agree on success before anyone uses (or gives up on) the peer mappings
        oks = [None] * self.world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            raise engine.SpmvB200Error("peer mapping failed on rank(s) %s: %s" % ([i for i, o in enumerate(oks) if not o], err))
        # halo=False: deliver the whole block to everybody (a fused all-gather) even where only a halo is read
        self.need = needed_rows(self.splits, self.col_ranges if halo else [(0, self.N - 1)] * self.world, self.rank)
        self.flag_ptrs = (C.c_void_p * self.world)(*[self.peer[p][2] for p in range(self.world)])
        dist.barrier(group=group)

    # ---- x on entry (replicated: every rank fills the whole vector)
    cur = 0

    def set_x(self, x_host):
        if self.mode == "nccl":
            self.x[0].copy_(self.torch.from_numpy(np.ascontiguousarray(x_host)))
        else:
            engine.check(engine.lib().spmvb200_h2d(self.bufs[0].data_ptr(), engine.ptr(np.ascontiguousarray(x_host, dtype=np.float64)),
                                                   self.N * 8), "h2d")
        self.cur = 0

    def step(self, stream=None):
        lib = engine.lib()
        cur, nxt = self.cur, 1 - self.cur
        if self.mode == "nccl":
            engine._launch(self.kind, self.dmat, self.x[cur], Config_none, self.y, stream)
            self.dist.all_gather_into_tensor(self.x[nxt], self.y, group=self.group)
        else:
            peers = sorted(self.need)
            engine.spmv_push(self.kind, self.dmat, self.bufs[cur].data_ptr(), self.bufs[nxt].data_ptr() + self.r0 * 8,
                             [self.peer[p][nxt] for p in peers], [self.need[p][0] for p in peers], [self.need[p][1] for p in peers],
                             self.r0, stream)
            self.epoch += 1
            engine.check(lib.spmvb200_peer_barrier(self.flag_ptrs, self.world, self.rank, self.epoch, stream), "peer_barrier")
        self.cur = nxt

    # ---- end-to-end step from host buffers (bench.py's N>1 e2e): no collective library in the data path
    def load_x_slice(self, host_ptr, stream=None, after_h2d=None):
        """This rank's slice x[r0:r1) comes up from (pinned) host memory over its own PCIe link into the current x buffer and
        is delivered -- only the rows they read -- to the other ranks by peer stores; the flag barrier makes every rank's
        x complete where it is read.  host_ptr: address of element r0.  The two x buffers alternate from call to call: a
        peer may still be reading the previous one (its SpMV of the last step) while this step's rows arrive."""
        assert self.mode == "push"
        lib = engine.lib()
        self.cur = 1 - self.cur
        mine = self.bufs[self.cur].data_ptr() + self.r0 * 8
        engine.check(lib.spmvb200_h2d_async(mine, host_ptr, (self.r1 - self.r0) * 8, stream), "h2d_async")
        if after_h2d:
            after_h2d()
        peers = sorted(self.need)
        p = engine.capi.Push()
        p.n = len(peers)
        for i, q in enumerate(peers):
            p.dst[i], p.lo[i], p.hi[i] = self.peer[q][self.cur], self.need[q][0], self.need[q][1]
        p.row_offset = self.r0
        engine.check(lib.spmvb200_push_rows(mine, self.r1 - self.r0, engine.C.byref(p), stream), "push_rows")
        self.epoch += 1
        engine.check(lib.spmvb200_peer_barrier(self.flag_ptrs, self.world, self.rank, self.epoch, stream), "peer_barrier")

    def x_ptr(self):
        return self.bufs[self.cur].data_ptr()

    def my_slice(self):
        """this rank's rows of the current x, as a host array"""
        if self.mode == "nccl":
            return self.x[self.cur][self.r0:self.r1].cpu().numpy()
        out = np.empty(self.r1 - self.r0, dtype=np.float64)
        engine.check(engine.lib().spmvb200_sync(), "sync")
        engine.check(engine.lib().spmvb200_d2h(engine.ptr(out), self.bufs[self.cur].data_ptr() + self.r0 * 8, out.nbytes), "d2h")
        return out

    def close(self):
        if self.mode == "push":
            engine.check(engine.lib().spmvb200_sync(), "sync")
            self.dist.barrier(group=self.group)
            for p, ptrs in self.peer.items():
                if p != self.rank:
                    for q in ptrs:
                        engine.lib().spmvb200_ipc_close(q)
            self.peer = {}
            self.dist.barrier(group=self.group)
            for v in self.bufs + [self.flags]:
                v.free()


class RowBlockShard:
    """This rank's share of a row-block partitioned square matrix, driven entirely inside the C library (spmvb200_shard_*,
    include/spmv_b200.h): `dmat` holds rows [splits[rank], splits[rank+1]) with global column ids; x is replicated in `nbuf`
    library-owned device buffers that the peers map through CUDA IPC.  torch.distributed only carries the one-time rendezvous
    (the IPC handle blobs); no collective runs in the data path.

      step(src, dst)          x[dst] <- A x[src]: the SpMV kernel stores the rows the peers read into THEIR x[dst], flag barrier
      spmv_host(x_sl, y_sl)   host x slice up, halo delivered to the peers + barrier, row chunks, y slice down -- one C call
    """

    def __init__(self, dmat, splits, kind, group=None, nbuf=3, col_range=None):
        import torch.distributed as dist
        self.dist, self.group, self.kind, self.dmat = dist, group, kind, dmat
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        self.splits = [int(v) for v in splits]
        assert len(self.splits) == self.world + 1
        self.r0, self.r1, self.N, self.nbuf = self.splits[self.rank], self.splits[self.rank + 1], dmat.N, nbuf
        lib, C = engine.lib(), engine.C
        sp_arr = (C.c_uint64 * (self.world + 1))(*self.splits)
        cr = None if col_range is None else (C.c_uint64 * 2)(int(col_range[0]), int(col_range[1]))
        h = C.c_void_p()
        err = None
        try:
            engine.check(lib.spmvb200_shard_create(dmat.handle, kind, self.rank, self.world, sp_arr, nbuf, cr, C.byref(h)), "shard_create")
            self._h = h.value
            nb = lib.spmvb200_shard_blob_bytes(self._h)
            blob = (C.c_ubyte * nb)()
            engine.check(lib.spmvb200_shard_export(self._h, blob), "shard_export")
            mine = bytes(blob)
        except engine.SpmvB200Error as e:  # keep going: every rank must reach the agreement below
            err, mine, self._h = e, b"", None
        if self.world > 1:
            allb = [None] * self.world
            dist.all_gather_object(allb, mine, group=group)
            if err is None and all(len(b) == len(mine) for b in allb):
                try:
                    joined = b"".join(allb)
                    engine.check(lib.spmvb200_shard_connect(self._h, (C.c_ubyte * len(joined)).from_buffer_copy(joined)), "shard_connect")
                except engine.SpmvB200Error as e:
                    err = e
            elif err is None:
                err = engine.SpmvB200Error("another rank could not create its shard")
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=group)
            if not all(oks):
                self.close(collective=False)
                raise engine.SpmvB200Error("shard set-up failed on rank(s) %s: %s" % ([i for i, o in enumerate(oks) if not o], err))
        elif err is not None:
            raise err

    def x_ptr(self, buf):
        return engine.lib().spmvb200_shard_x(self._h, buf)

    def set_x(self, buf, x_host):
        """replicated x: every rank uploads the whole vector into buffer `buf`"""
        x_host = np.ascontiguousarray(x_host, dtype=np.float64)
        assert len(x_host) == self.N
        engine.check(engine.lib().spmvb200_h2d(self.x_ptr(buf), engine.ptr(x_host), self.N * 8), "h2d")

    @property
    def halo_rows(self):
        n = engine.C.c_uint64()
        engine.check(engine.lib().spmvb200_shard_halo_rows(self._h, engine.C.byref(n)), "shard_halo_rows")
        return n.value

    def step(self, src, dst, stream=None):
        engine.check(engine.lib().spmvb200_shard_step(self._h, src, dst, stream), "shard_step")

    def barrier(self, stream=None):
        """the cross-GPU flag barrier on its own, asynchronous on `stream` (every rank calls it)"""
        engine.check(engine.lib().spmvb200_shard_barrier(self._h, stream), "shard_barrier")

    def spmv_host(self, x_slice, y_slice, timed=True):
        """timed=False: NULL kernel_ms, no time-stamped events between the chunks (faster), returns None"""
        if not timed:
            engine.check(engine.lib().spmvb200_shard_spmv_host(self._h, engine.ptr(x_slice), engine.ptr(y_slice), None), "shard_spmv_host")
            return None
        ms = engine.C.c_float(0)
        engine.check(engine.lib().spmvb200_shard_spmv_host(self._h, engine.ptr(x_slice), engine.ptr(y_slice), engine.C.byref(ms)), "shard_spmv_host")
        return ms.value

    def rows_of(self, buf, a=None, b=None):
        """global rows [a, b) (default: my slice) of x[buf] as a host array"""
        a = self.r0 if a is None else a
        b = self.r1 if b is None else b
        out = np.empty(b - a, dtype=np.float64)
        engine.check(engine.lib().spmvb200_sync(), "sync")
        if b > a:
            engine.check(engine.lib().spmvb200_d2h(engine.ptr(out), self.x_ptr(buf) + a * 8, out.nbytes), "d2h")
        return out

    def close(self, collective=True):
        if getattr(self, "_h", None):
            engine.check(engine.lib().spmvb200_sync(), "sync")
            if collective and self.world > 1:
                self.dist.barrier(group=self.group)  # nobody unmaps while a peer may still store into its buffers
            h, self._h = self._h, None
            engine.lib().spmvb200_shard_free(h)
            if collective and self.world > 1:
                self.dist.barrier(group=self.group)
