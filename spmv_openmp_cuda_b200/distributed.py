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
