#!/usr/bin/env python
"""Iterated SpMV x <- A x over the GPUs of one box (run under torchrun, one rank per GPU): parity of the fused
peer-store exchange against the oracle, then time per iteration of push (fused) vs NCCL all-gather.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_iterate.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import synth  # noqa: E402
from spmv_openmp_cuda_b200.distributed import RowBlockIterate, RowBlockShard, row_partition_uniform  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    sp.capi.check(sp.capi.lib().spmvb200_set_device(local), "set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream().cuda_stream
    ok = True

    # ---- parity: small banded and unstructured matrices, 5 iterations, every rank's slice against the oracle
    import oracle
    for name, spec_or_mat in (("banded", synth.banded(60000 * world, 32, 5000)), ("rmat", None)):
        mat = synth.host_csr(spec_or_mat) if spec_or_mat is not None else synth.rmat_host_csr(14, 16)
        splits = row_partition_uniform(mat.M, world)
        x0 = synth.host_vector(mat.N) * 1e4
        ref = x0.copy()
        for _ in range(5):
            ref = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, ref)
        for kind_name in ("xwin", "csr_rows", "ell"):
            d_csr = sp.spMatCpyCSR(mat, splits[rank], splits[rank + 1])
            if kind_name == "xwin":
                if mat.MAX_ROW_NZ > 255:
                    continue
                dm, kind = d_csr.to_xwin(512, 1024), sp.XWIN_ROWS
            elif kind_name == "ell":
                if mat.MAX_ROW_NZ > 255:
                    continue
                dm, kind = d_csr.to_ell(sp.FMT_ELL_COLMAJOR), sp.ELL_ROWS
            else:
                dm, kind = d_csr, sp.CSR_ROWS
            for mode in ("push", "nccl"):
                if mode == "nccl" and len({b - a for a, b in zip(splits[:-1], splits[1:])}) != 1:
                    continue
                it = RowBlockIterate(_with_cols(dm, d_csr), splits, kind, mode=mode)
                it.set_x(x0)
                for _ in range(5):
                    it.step(stream)
                torch.cuda.synchronize()
                mine = it.my_slice()
                want = ref[splits[rank]:splits[rank + 1]]
                if mat.MAX_ROW_NZ <= 2048:  # every kind used here then sums in the serial order: bit-identical iterates
                    good = bool(np.array_equal(mine, want))
                else:                       # rows split across CTAs: deterministic but not the serial order -> norm-wise
                    good = bool(np.max(np.abs(mine - want)) <= 1e-9 * np.max(np.abs(ref)))
                flag = torch.tensor([1 if good else 0], device="cuda")
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if rank == 0:
                    print("parity %-7s %-8s %-5s world=%d halo=%s -> %s" % (name, kind_name, mode, world,
                          getattr(it, "need", None) if mat.M < 10 ** 6 else "", "OK" if flag.item() else "FAILED"), flush=True)
                ok = ok and bool(flag.item())
                it.close()

    # ---- the same through the C-level shard (spmvb200_shard_*): device iterates and the one-call host step, every rank's slice
    for name, spec_or_mat in (("banded", synth.banded(70000 * world, 32, 5000)), ("rmat", None)):
        mat = synth.host_csr(spec_or_mat) if spec_or_mat is not None else synth.rmat_host_csr(14, 16)
        splits = row_partition_uniform(mat.M, world)
        r0, r1 = splits[rank], splits[rank + 1]
        x0 = synth.host_vector(mat.N) * 1e4
        ref1 = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, x0)
        ref3 = ref1
        for _ in range(2):
            ref3 = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, ref3)
        exact = mat.MAX_ROW_NZ <= 2048
        for kind_name in ("csr_rows", "xwin", "ell", "adaptive"):
            d_csr = sp.spMatCpyCSR(mat, r0, r1)
            if kind_name in ("xwin", "ell") and mat.MAX_ROW_NZ > 255:
                continue
            dm, kind = {"csr_rows": (d_csr, sp.CSR_ROWS), "adaptive": (d_csr, sp.CSR_ADAPTIVE)}.get(kind_name) or \
                ((d_csr.to_xwin(512, 1024), sp.XWIN_ROWS) if kind_name == "xwin" else (d_csr.to_ell(sp.FMT_ELL_COLMAJOR), sp.ELL_ROWS))
            sh = RowBlockShard(dm, splits, kind, nbuf=3, col_range=d_csr.col_range if d_csr.NZ else (1, 0))
            sh.set_x(0, x0)
            sh.step(0, 1, stream)
            sh.step(1, 2, stream)
            sh.step(2, 1, stream)
            torch.cuda.synchronize()
            mine = sh.rows_of(1)
            tol = 1e-9 * np.max(np.abs(ref3))
            good = bool(np.array_equal(mine, ref3[r0:r1])) if (exact and kind_name != "adaptive") else bool(np.max(np.abs(mine - ref3[r0:r1]), initial=0.0) <= tol)
            y = np.full(r1 - r0, np.nan)
            for _ in range(3):
                sh.spmv_host(x0[r0:r1], y)
            good_h = bool(np.array_equal(y, ref1[r0:r1])) if (exact and kind_name != "adaptive") else bool(np.max(np.abs(y - ref1[r0:r1]), initial=0.0) <= 1e-9 * np.max(np.abs(ref1)))
            flag = torch.tensor([1 if (good and good_h) else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print("parity shard %-7s %-8s world=%d halo_rows=%d -> %s" % (name, kind_name, world, sh.halo_rows, "OK" if flag.item() else "FAILED"), flush=True)
            ok = ok and bool(flag.item())
            sh.close()
    sp.capi.lib().spmvb200_host_unregister(None)

    if "--parity-only" in sys.argv:
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("MULTI_GPU_ITERATE", "OK" if flag.item() else "FAILED", flush=True)
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if flag.item() == 1 else 1)

    # ---- timing: banded 2^22 rows per GPU, w = 2^15 (cfg4 row density), x-window kernel
    rows_per = 1 << 22
    M = rows_per * world
    spec = synth.banded(M, 32, 1 << 15)
    splits = [g * rows_per for g in range(world + 1)]
    d_csr = synth.device_csr(spec, splits[rank], splits[rank + 1])
    dm = d_csr.to_xwin()
    res = {}
    for mode in ("push", "push_full", "nccl"):
        it = RowBlockIterate(_with_cols(dm, d_csr), splits, sp.XWIN_ROWS, mode=mode.split("_")[0], halo=(mode != "push_full"))
        it.set_x(synth.host_vector(M) * 1e4)
        for _ in range(4):
            it.step(stream)
        iters = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(iters):
            it.step(stream)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = float(t.item())
        it.close()
    # kernel alone (x resident, no exchange)
    x = torch.zeros(M, dtype=torch.float64, device="cuda")
    y = torch.zeros(rows_per, dtype=torch.float64, device="cuda")
    tk = sp.time_kernel(sp.XWIN_ROWS, dm, x, y, reps=20)
    t = torch.tensor([float(tk.mean())], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        nnz = 32 * M
        out = {"what": "iterated SpMV x<-Ax, banded %d rows/GPU x 32 nnz/row, w=2^15, x-window kernel" % rows_per, "n_gpus": world,
               "ms_per_iter_push_fused_halo": res["push"], "ms_per_iter_push_fused_whole_block": res["push_full"],
               "ms_per_iter_nccl_allgather": res["nccl"], "ms_kernel_only": float(t.item()),
               "gflops_push_halo": 2 * nnz / res["push"] / 1e6, "gflops_push_whole_block": 2 * nnz / res["push_full"] / 1e6,
               "gflops_nccl": 2 * nnz / res["nccl"] / 1e6, "gflops_kernel_only": 2 * nnz / float(t.item()) / 1e6, "parity_ok": ok}
        print("ITERATE " + json.dumps(out), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


class _with_cols:
    """a device matrix in any format + the column range taken from its CSR source (RowBlockIterate asks the handle)"""

    def __init__(self, dm, d_csr):
        self._dm, self._cr = dm, d_csr.col_range if d_csr.NZ else None
        self.M, self.N, self.NZ, self.handle = dm.M, dm.N, dm.NZ, dm.handle

    @property
    def col_range(self):
        return self._cr


if __name__ == "__main__":
    main()
