#!/usr/bin/env python
"""Where does a multi-GPU end-to-end step spend its time?  (torchrun, one rank per GPU.)  Phases of bench.py's N>1 e2e step, each
bracketed by CUDA events: H2D of this rank's x slice from pinned memory, NCCL all-gather of x, SpMV, D2H of the y slice."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spmv_openmp_cuda_b200 as sp  # noqa: E402
from spmv_openmp_cuda_b200 import capi, synth  # noqa: E402


REPS = int(os.environ.get("PROBE_REPS", "20"))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    capi.check(capi.lib().spmvb200_set_device(local), "set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows_per = 128 ** 3
    spec = synth.stencil27(128, 128, 128 * world)
    d_csr = synth.device_csr(spec, rank * rows_per, (rank + 1) * rows_per)
    dm = d_csr.to_ell(sp.FMT_ELL_COLMAJOR)
    N = dm.N
    x = torch.zeros(N, dtype=torch.float64, device="cuda")
    y = torch.zeros(dm.M, dtype=torch.float64, device="cuda")
    hx = torch.zeros(N, dtype=torch.float64).pin_memory()
    hy = torch.zeros(dm.M, dtype=torch.float64).pin_memory()
    a, b = rank * N // world, (rank + 1) * N // world
    stream = torch.cuda.current_stream().cuda_stream
    lib = capi.lib()

    if os.environ.get("PROBE_FILL"):
        synth.device_vector_fill(x, N)
        hx.copy_(x.cpu())
    if os.environ.get("PROBE_PRELOOP"):
        for _ in range(100):
            capi.check(lib.spmvb200_spmv_device(dm.handle, sp.ELL_ROWS, x.data_ptr(), y.data_ptr(), stream), "spmv")
        torch.cuda.synchronize()
        dist.barrier()
    if os.environ.get("PROBE_FREE"):
        d_csr.free()
    only = os.environ.get("PROBE_ONLY")
    phases = {
        "h2d_slice": lambda: x[a:b].copy_(hx[a:b], non_blocking=True),
        "allgather": lambda: dist.all_gather_into_tensor(x, x[a:b]),
        "spmv": lambda: capi.check(lib.spmvb200_spmv_device(dm.handle, sp.ELL_ROWS, x.data_ptr(), y.data_ptr(), stream), "spmv"),
        "d2h_slice": lambda: hy.copy_(y, non_blocking=True),
    }

    four = list(phases.values())

    def all_of_them():
        for f in four:
            f()
    phases["whole_step"] = all_of_them

    def synced(f):
        def g():
            f()
            torch.cuda.current_stream().synchronize()
        return g
    phases["whole_step_synced"] = synced(all_of_them)
    phases["allgather_synced"] = synced(phases["allgather"])
    phases["h2d_synced"] = synced(phases["h2d_slice"])
    phases["h2d_allgather_synced"] = synced(lambda: (four[0](), four[1]()))
    phases["allgather_spmv_d2h_synced"] = synced(lambda: (four[1](), four[2](), four[3]()))
    out = {}
    for name, f in phases.items():
        if only and name not in only.split(","):
            continue
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(REPS):
            f()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / REPS * 1e3
        t = torch.tensor([e0.elapsed_time(e1) / REPS, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = [round(float(v), 4) for v in t.tolist()]
    if rank == 0:
        print("E2E_PROBE world=%d pinned=%s/%s  (ms: [cuda events, wall])" % (world, hx.is_pinned(), hy.is_pinned()), out, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
