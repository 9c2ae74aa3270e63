"""Row-block partition on >= 2 real GPUs (skipped on a 1-GPU box): torchrun + NCCL + the CUDA engine."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_row_block_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_GPU_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_two_gpu_fused_iteration_parity():
    """x <- A x over 2 GPUs with the exchange fused into the SpMV epilogue (peer stores over CUDA IPC + flag barrier) and with an
    NCCL all-gather: every rank's slice equals the oracle's iterates (bit for bit where the kind sums in the serial order)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29518", os.path.join(ROOT, "tools", "multi_gpu_iterate.py"), "--parity-only"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_GPU_ITERATE OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_two_gpu_shard_fused_neighbour_sync_parity():
    """The same parity run with SPMVB200_SHARD_FUSED_SYNC=1: the x-window kernel's boundary CTAs carry the step's synchronisation
    (wait for the neighbours' previous step, last boundary CTA publishes completion) instead of the separate barrier kernel."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tools", "multi_gpu_iterate.py"), "--parity-only"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, SPMVB200_SHARD_FUSED_SYNC="1"))
    assert out.returncode == 0 and "MULTI_GPU_ITERATE OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
