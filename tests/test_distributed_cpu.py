"""Row-block partition + collective plumbing on CPU: world_size 2 and 3 over gloo.  The local
compute is injected (the oracle) -- the product's own local compute is CUDA only and is covered by
the -m gpu tests; here we check split points, the x broadcast and the y slice gather/all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, balance, ret):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import spmv_openmp_cuda_b200 as sp
    from spmv_openmp_cuda_b200.distributed import RowBlockSpmv

    mat = sp.synth.rmat_host_csr(10, 8)  # skewed rows: nnz balance differs from row balance
    x = sp.synth.host_vector(mat.N)
    holder = {}

    def local(x_t, y_t):  # checker-side compute on this rank's rows
        a, b = holder["op"].r0, holder["op"].r1
        irp = (mat.IRP[a:b + 1] - mat.IRP[a]).astype(np.uint64)
        lo, hi = int(mat.IRP[a]), int(mat.IRP[b])
        y_t.copy_(torch.from_numpy(oracle.sgemv_serial(irp, mat.JA[lo:hi], mat.AS[lo:hi], x_t.numpy())))

    op = RowBlockSpmv(mat, balance=balance, device=torch.device("cpu"), local_spmv=local)
    holder["op"] = op
    y = op.spmv(x if rank == 0 else None)
    y_all = op.spmv_allgather()
    y_ref = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, x)
    ok = np.array_equal(y_all.numpy(), y_ref)
    if rank == 0:
        ok = ok and np.array_equal(y, y_ref)
        ret["splits"] = op.splits
        ret["nnz"] = [int(mat.IRP[b] - mat.IRP[a]) for a, b in zip(op.splits[:-1], op.splits[1:])]
    else:
        ok = ok and y is None
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,balance", [(2, "nnz"), (3, "rows")])
def test_row_block_spmv_gloo(world, balance):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), balance, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
    assert ret["splits"][0] == 0 and ret["splits"][-1] == 1024 and len(ret["splits"]) == world + 1
    if balance == "nnz":
        nnz = ret["nnz"]
        assert max(nnz) - min(nnz) <= 0.1 * sum(nnz), nnz  # balanced within 10 %


def test_partition_functions():
    from spmv_openmp_cuda_b200.distributed import row_partition_by_nnz, row_partition_uniform
    assert row_partition_uniform(10, 4) == [0, 3, 6, 8, 10]  # first M % G blocks get one extra row
    assert row_partition_uniform(3, 5) == [0, 1, 2, 3, 3, 3]
    irp = np.array([0, 100, 100, 101, 102, 200], dtype=np.uint64)
    assert row_partition_by_nnz(irp, 2) == [0, 1, 5]
    assert row_partition_by_nnz(irp, 1) == [0, 5]
    pts = row_partition_by_nnz(np.arange(0, 33 * 32, 32, dtype=np.uint64), 8)
    assert pts == [0, 4, 8, 12, 16, 20, 24, 28, 32]
    assert row_partition_by_nnz(np.zeros(5, dtype=np.uint64), 3) == [0, 0, 0, 4]


def test_needed_rows_halo_planning():
    """Which of a rank's rows the other ranks read (fused delivery of the multi-GPU iteration): halos for a banded matrix,
    whole blocks for an unstructured one, nothing for a rank without non-zeros."""
    from spmv_openmp_cuda_b200.distributed import needed_rows
    splits = [0, 100, 200, 300]
    banded = [(0, 120), (80, 230), (170, 299)]
    assert needed_rows(splits, banded, 0) == {1: (80, 100)}
    assert needed_rows(splits, banded, 1) == {0: (100, 121), 2: (170, 200)}
    assert needed_rows(splits, banded, 2) == {1: (200, 231)}
    dense = [(0, 299)] * 3
    assert needed_rows(splits, dense, 1) == {0: (100, 200), 2: (100, 200)}
    assert needed_rows(splits, [(0, 50), None, (250, 299)], 1) == {}
