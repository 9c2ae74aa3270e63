#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): row-block partitioned SpMV with
the CUDA engine, x broadcast over NCCL, y gathered on rank 0 and compared with the oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200.distributed import RowBlockSpmv  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    sp.capi.check(sp.capi.lib().spmvb200_set_device(local), "set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, mat in (("banded", sp.synth.host_csr(sp.synth.banded(1 << 18, 32, 4096))), ("rmat", sp.synth.rmat_host_csr(15, 16))):
        x = sp.synth.host_vector(mat.N)
        for kind in (sp.CSR_ROWS, sp.CSR_ROWS_WARP, sp.CSR_ADAPTIVE):
            op = RowBlockSpmv(mat, kind=kind, balance="nnz")
            y = op.spmv(x if rank == 0 else None)
            y_all = op.spmv_allgather().cpu().numpy()
            if rank == 0:
                import oracle
                y_ref = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)
                bad, worst = oracle.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y, tau=1e-12)
                bad2, _ = oracle.strict_diff_csr(mat.IRP, mat.JA, mat.AS, x, y_ref, y_all, tau=1e-12)
                exact = kind == sp.CSR_ROWS and bool(np.array_equal(y[np.diff(mat.IRP) <= 2048], y_ref[np.diff(mat.IRP) <= 2048]))
                print("%s kind=%d world=%d splits=%s bad=%d/%d worst=%.2e exact=%s" % (name, kind, op.world, op.splits, bad, bad2, worst, exact))
                ok = ok and bad == 0 and bad2 == 0 and (kind != sp.CSR_ROWS or exact)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
