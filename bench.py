#!/usr/bin/env python
"""Headline benchmark of the B200 SpMV engine (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one SpMV (y = A*x, fp64) over the workload:
  cfg2 (default): BASELINE.json configs[1] -- ELL, 27-point 3-D stencil on 128^3 PER GPU (2 097 152 rows,
        55 742 968 nnz per GPU).  With N GPUs the global grid is 128 x 128 x (128*N), row-block
        partitioned in z-slabs (weak scaling), x replicated on every GPU.
  cfg4: BASELINE.json configs[3] -- CSR, random banded 2^25 rows x 32 nnz/row, row-block partitioned over
        the N GPUs (strong scaling).
`value`  : whole-job GFLOP/s (2*nnz/t), matrix and x resident in HBM, CUDA-event timed on the launch
           stream, max over ranks.
`e2e`    : the same metric through the host-buffer call spmvb200_spmv_host (the SPMV_INTERF-shaped entry
           point of include/spmv_b200.h): per step x goes host(pinned)->device, [N>1: NCCL broadcast of x
           from rank 0 over NVLink], kernel, y slice device->host(pinned).
--impl reference times the reference's own OpenMP CPU implementation (oracle/_ref, compiled from the
unmodified sources; falls back to the oracle port) on the host cores, rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# libgomp reads OMP_SCHEDULE when it is loaded: set it before anything pulls libgomp in.  The reference's
# kernels use schedule(runtime); plain "static" crashes its ompGetRuntimeSchedule (ompGetICV.c:43)
os.environ.setdefault("OMP_SCHEDULE", "nonmonotonic:static")
if "LOCAL_RANK" not in os.environ or "reference" in sys.argv:
    # CPU baseline / reference arm only.  NEVER under torchrun for our own arm: with OMP_NUM_THREADS=1 (torchrun's default) a bound
    # libgomp pins every rank's main thread to the same core, and the ranks then time-slice on it -- measured: a 6.5 ms stall
    # at every cross-GPU synchronisation point of the end-to-end loop (4.1 ms per step instead of 0.8 at 2 GPUs)
    os.environ.setdefault("OMP_PROC_BIND", "close")
if "reference" in sys.argv and "LOCAL_RANK" in os.environ:
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU reference arm (rank 0 only) uses all host cores
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

METRIC = "SpMV GFLOP/s (2*nnz/t), fp64"
UNIT = "GFLOP/s"


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, False, None
        try:
            if os.environ.get("BENCH_NO_NVML"):
                raise RuntimeError("sampling switched off")
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop = True
        if self._t:
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(spec_fn, steps, warmup, label):
    """Time the reference's OpenMP CPU SpMV (best of its CSR/ELL row kernels) on the host cores.
    Returns (value GFLOP/s, ms_per_step, info dict)."""
    import oracle
    from spmv_openmp_cuda_b200 import synth

    t0 = time.time()
    mat = synth.host_csr(spec_fn())
    x = synth.host_vector(mat.N)
    cores = oracle.omp_max_threads()
    flops = 2.0 * mat.NZ
    results = {}
    if oracle.ref_available():
        kind = "reference"
        grid = int(min(65535, 4 * cores))  # gridRows caps the parallelism of *Blocks* (SpMV_CSR_OMP.c:70-76)
        rmat = oracle.ref_spmat(mat.M, mat.N, mat.NZ, mat.JA, mat.AS, irp=mat.IRP, rl=mat.RL)
        ell = synth.csr_to_ell_host(mat)
        emat = oracle.ref_spmat(ell.M, ell.N, ell.NZ, ell.JA, ell.AS, rl=ell.RL, max_row_nz=ell.MAX_ROW_NZ)
        # both settings of SIMD_ROWS_REDUCTION (src/include/config.h:92-94; TRUE is the reference's default): same unmodified sources
        variants = [v for v in ("default", "nosimd") if oracle.ref_available(v)]
        cfgs = {v: oracle.ref_config(grid_rows=grid, grid_cols=8, threads=cores, chunks=0, variant=v) for v in variants}
        cands = []
        for v in variants:
            tag = "" if v == "default" else "[SIMD_ROWS_REDUCTION=FALSE]"
            cands += [("spmvRowsBlocksCSR" + tag, (rmat, v)), ("spmvRowsBasicCSR" + tag, (rmat, v)),
                      ("spmvRowsBlocksELL" + tag, (emat, v)), ("spmvRowsBasicELL" + tag, (emat, v))]

        def run(name, m):
            return oracle.ref_call(name.split("[")[0], m[0], x, cfgs[m[1]], mat.M, variant=m[1])
    else:
        kind = "port"
        cands = [("oracle_spmv_rows_blocks_csr", None), ("oracle_spmv_rows_basic_csr", None)]

        def run(name, m):
            if name == "oracle_spmv_rows_blocks_csr":
                return oracle.spmv_rows_blocks_csr(mat.IRP, mat.JA, mat.AS, x, grid_rows=4 * cores)
            return oracle.spmv_rows_basic_csr(mat.IRP, mat.JA, mat.AS, x)
    y_ref = None
    for name, m in cands:  # pick the fastest variant on 3 probes each
        ts = []
        for _ in range(3):
            t = time.perf_counter()
            y = run(name, m)
            ts.append(time.perf_counter() - t)
        if y_ref is None:
            y_ref = oracle.sgemv_serial(mat.IRP, mat.JA, mat.AS, x) if mat.NZ < 2e8 else y
        assert not oracle.double_vectors_diff(y_ref, y)[0], name
        results[name] = min(ts)
    best = min(results, key=results.get)
    m = dict(cands)[best]
    for _ in range(warmup):
        run(best, m)
    t = time.perf_counter()
    for _ in range(steps):
        run(best, m)
    dt = (time.perf_counter() - t) / max(steps, 1)
    info = {"value": flops / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%s, full per-GPU matrix (M=%d, nnz=%d), %d timed calls of %s after %d warm-ups; probes (s): %s; setup %.1f s"
                      % (label, mat.M, mat.NZ, steps, best, warmup,
                         {k: round(v, 5) for k, v in results.items()}, time.time() - t0)}
    return flops / dt / 1e9, dt * 1e3, info


# ------------------------------------------------------------------------------------------ main
def emit(obj):
    """The one JSON line, on the real stdout (fd 1 is pointed at stderr while the benchmark runs so that
    library banners such as 'NCCL version ...' cannot pollute it)."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of one block of the end-to-end loop (default: min(steps, 200))")
    ap.add_argument("--e2e-blocks", type=int, default=5, help="blocks of the end-to-end loop; the median block is reported")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--half-width", type=int, default=1 << 15, help="cfg4: band half width w")
    ap.add_argument("--x-dist", default="push", choices=["push", "allgather", "bcast"],
                    help="N>1 e2e: every rank uploads its slice of x, then either delivers the rows the other ranks read by peer stores + "
                         "flag barrier (push, default), or NCCL all-gather; or rank 0 uploads all of x then NCCL broadcast")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N = args.gpus

    from spmv_openmp_cuda_b200 import synth

    if args.workload == "cfg2":
        wl_name = ("cfg2: ELL fp64 SpMV, 27-point 3-D stencil 128^3 per GPU (2097152 rows, 55742968 nnz per GPU), "
                   "column-major pitched ELL + row-length early exit, kernel cudaSpMVRowsELL")
        scaling = "weak"

        def slab_spec():
            return synth.stencil27(128, 128, 128)
    else:
        wl_name = ("cfg4: CSR fp64 SpMV, random banded 2^25 rows x 32 nnz/row (half width w=%d), row-block partitioned, "
                   "kernel: adaptive CSR mode" % args.half_width)
        scaling = "strong"

        def slab_spec():
            return synth.banded(1 << 21, 32, args.half_width)  # bounded CPU sample: 2^21 rows of the same generator

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = min(args.steps, 50)
        value, ms, info = cpu_reference(slab_spec, steps, min(args.warmup, 5), args.workload)
        emit(({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": steps,
                          "warmup": min(args.warmup, 5), "ms_per_step": ms, "higher_is_better": True, "scaling": scaling,
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": wl_name, "note": "reference OpenMP CPU implementation on the host cores"},
                          "cpu_baseline": info,
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist

    import spmv_openmp_cuda_b200 as sp
    from spmv_openmp_cuda_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    capi.check(capi.lib().spmvb200_set_device(local_rank), "set_device")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == N or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    nr = world  # ranks actually running

    # ---- build this rank's row block on its GPU
    if args.workload == "cfg2":
        spec = synth.stencil27(128, 128, 128 * nr)
        rows_per = 128 * 128 * 128
        r0, r1 = rank * rows_per, (rank + 1) * rows_per
        d_csr = synth.device_csr(spec, r0, r1)
        dm = d_csr.to_ell(sp.FMT_ELL_COLMAJOR)
        col_range = d_csr.col_range
        d_csr.free()
        # 16-bit column offsets: the two-rows-per-thread kernel; otherwise the one-row-per-thread kernel with 32-bit ids
        kind, kname = sp.ELL_ROWS, ("ell_colmajor_pair_kernel<idx16>" if dm.index_bits == 16 else "ell_colmajor_kernel")
    else:
        spec = synth.banded(1 << 25, 32, args.half_width)
        Mtot = 1 << 25
        r0, r1 = rank * Mtot // nr, (rank + 1) * Mtot // nr
        dm = synth.device_csr(spec, r0, r1)
        col_range = dm.col_range
        kind, kname = sp.CSR_ADAPTIVE, "csr_adaptive"
    Ncols = dm.N
    tot = torch.tensor([dm.NZ, dm.M], dtype=torch.int64, device="cuda")
    if world > 1 and "allreduce" not in os.environ.get("BENCH_SKIP", ""):
        dist.all_reduce(tot)
    elif world > 1:
        tot *= world
    nnz_total, rows_total = int(tot[0].item()), int(tot[1].item())
    # algorithmic bytes of the GLOBAL SpMV, split evenly: x is counted once for the whole job
    rowmeta = 4 * rows_total if args.workload == "cfg2" else 4 * (rows_total + 1)
    bytes_local = (12 * nnz_total + rowmeta + 8 * Ncols + 8 * rows_total) // nr

    x = torch.empty(Ncols, dtype=torch.float64, device="cuda")
    y = torch.empty(dm.M, dtype=torch.float64, device="cuda")
    synth.device_vector_fill(x, Ncols)
    stream = torch.cuda.current_stream().cuda_stream
    lib = capi.lib()

    def step():
        capi.check(lib.spmvb200_spmv_device(dm.handle, kind, x.data_ptr(), y.data_ptr(), stream), "spmv_device")

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    skip = os.environ.get("BENCH_SKIP", "").split(",")
    for _ in range(0 if "warmup" in skip else args.warmup):
        step()
    sync_all()
    if kind == sp.CSR_ADAPTIVE:
        kname = "csr_adaptive[%s]" % dm.adaptive_choice
    idx_bits = dm.index_bits if kind == sp.ELL_ROWS else (16 if "xwindow" in kname else 32)
    launches0 = lib.spmvb200_launch_count()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    ev0.record()
    for _ in range(1 if "kloop" in skip else args.steps):
        step()
    ev1.record()
    sync_all()
    clocks = sampler.stop()
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    launches = int(lib.spmvb200_launch_count() - launches0)
    gflops = 2.0 * nnz_total / (ms_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    achieved = bytes_local / (ms_step * 1e-3) / 1e9  # per GPU: the dominant kernel's launch on one GPU

    # ---- end to end through the host-buffer entry point
    e2e_steps = args.e2e_steps or min(args.steps, 200)
    hx = torch.empty(Ncols, dtype=torch.float64).pin_memory()
    hy = torch.empty(dm.M, dtype=torch.float64).pin_memory()
    hx.copy_(x.cpu())
    e2e_ms = None
    xs0, xs1 = rank * Ncols // nr, (rank + 1) * Ncols // nr  # this rank's slice of x (square matrix: its own rows)
    even = Ncols % nr == 0

    dbg = os.environ.get("BENCH_E2E_DEBUG")
    dbg_t = []

    def mark():
        if dbg:
            torch.cuda.synchronize()
            dbg_t.append(time.perf_counter())

    pusher, push_err = None, ""
    if world > 1 and args.x_dist == "push":
        from spmv_openmp_cuda_b200.distributed import RowBlockIterate

        class _Rows:  # the handle + the column range of its CSR source
            M, N, NZ, handle = dm.M, dm.N, dm.NZ, dm.handle
        _Rows.col_range = col_range
        allr = [None] * world
        dist.all_gather_object(allr, (r0, r1))
        try:
            pusher = RowBlockIterate(_Rows, [a_[0] for a_ in allr] + [allr[-1][1]], kind, mode="push")
            pusher.set_x(hx.numpy())
        except Exception as e:  # noqa: BLE001  (e.g. CUDA IPC not permitted in this container)
            pusher, push_err = None, repr(e)
        ok_all = torch.tensor([1 if pusher is not None else 0], device="cuda")
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        if not ok_all.item():  # every rank takes the same path: fall back to the NCCL all-gather of x
            pusher = None
            args.x_dist = "allgather"
            if rank == 0:
                sys.stderr.write("bench.py: peer-store exchange unavailable (%s); using --x-dist allgather\n" % (push_err or "another rank failed"))

    def e2e_step():
        mark()
        if world == 1:
            capi.check(lib.spmvb200_spmv_host(dm.handle, kind, hx.data_ptr(), hy.data_ptr(), None), "spmv_host")
        elif pusher is not None:
            pusher.load_x_slice(hx.data_ptr() + r0 * 8, stream, after_h2d=mark)  # own slice over own PCIe link, halo rows to the peers, barrier
            mark()
            capi.check(lib.spmvb200_spmv_device(dm.handle, kind, pusher.x_ptr(), y.data_ptr(), stream), "spmv_device")
            mark()
            capi.check(lib.spmvb200_d2h_async(hy.data_ptr(), y.data_ptr(), dm.M * 8, stream), "d2h_async")
            capi.check(lib.spmvb200_stream_sync(stream), "stream_sync")
            mark()
        else:
            if args.x_dist == "allgather" and even:
                x[xs0:xs1].copy_(hx[xs0:xs1], non_blocking=True)      # every rank: its slice over its own PCIe link
                mark()
                dist.all_gather_into_tensor(x, x[xs0:xs1])             # replicate over NVLink / NVSwitch
                mark()
            else:
                if rank == 0:
                    x.copy_(hx, non_blocking=True)
                dist.broadcast(x, src=0)
            step()
            mark()
            hy.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            mark()

    e2e_stream = None
    if world > 1 and not os.environ.get("BENCH_E2E_LEGACY_STREAM"):
        # the end-to-end loop runs on a stream of its own (not the legacy default stream, which synchronises implicitly with
        # every other blocking stream of the process)
        e2e_stream = torch.cuda.Stream()
        e2e_stream.wait_stream(torch.cuda.current_stream())
        torch.cuda.set_stream(e2e_stream)
        stream = e2e_stream.cuda_stream
    for _ in range(max(3, args.warmup)):
        e2e_step()
    # The same binary measured 0.53-0.88 ms per step in four back-to-back runs on one box (profiles/README.md, r01k).  It is a warm-up effect of
    # the host<->device legs, not of the buffers: in one process the first 150 calls ran at 0.83 ms and every later one at 0.54 ms, with
    # cudaHostAlloc'ed and with huge-page + cudaHostRegister'ed buffers alike (tools/hugepage_probe.py, profiles/r01m_e2e_warmup_probe.log).
    # So the loop is timed in blocks of e2e_steps steps and the MEDIAN block is reported (all blocks are listed next to it).
    block_ms, block_wall = [], []
    launches_e2e0 = int(lib.spmvb200_launch_count())
    for _ in range(max(1, args.e2e_blocks)):
        sync_all()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(e2e_steps):
            e2e_step()
        ev1.record()
        sync_all()
        block_wall.append((time.perf_counter() - t0) * 1e3)
        e2e_t = torch.tensor([max(ev0.elapsed_time(ev1), 0.0)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        block_ms.append(float(e2e_t.item()) / e2e_steps)
    mid = sorted(range(len(block_ms)), key=lambda i: block_ms[i])[len(block_ms) // 2]
    e2e_ms, wall = block_ms[mid], block_wall[mid]
    e2e_launch = int(lib.spmvb200_launch_count()) - launches_e2e0  # kernels launched inside the timed blocks (all of them)
    if dbg and world > 1 and len(dbg_t) >= 10:
        last = dbg_t[-16:]
        sys.stderr.write("rank %d e2e marks per step [start | h2d | allgather | spmv | d2h+sync], deltas in ms: %s\n" % (
            rank, " ".join("%.3f" % ((b - a) * 1e3) for a, b in zip(last[:-1], last[1:]))))
    y_check = hy.numpy().copy()

    # ---- parity spot check of what was just measured (rank 0, sampled rows, against the oracle)
    parity = None
    if rank == 0:
        import oracle
        a, b = 0, min(dm.M, 20000)
        h = synth.host_csr(spec, r0 + a, r0 + b)
        xs = hx.numpy()
        yr = oracle.sgemv_serial(h.IRP, h.JA, h.AS, xs)
        parity = {"rows_checked": int(b - a), "bit_identical": bool(np.array_equal(yr, y_check[a:b])),
                  "ref_check_failed": bool(oracle.double_vectors_diff(yr, y_check[a:b])[0])}

    out = {
        "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl_name, "rows_per_gpu": int(dm.M), "nnz_total": nnz_total, "parallelism": "row-block x%d" % nr,
                   "l2": "no flush: per-GPU inputs (%.0f MB) exceed the 126 MB L2" % (bytes_local / 1e6),
                   "x": "replicated on every GPU, resident for `value`; host->device (+NCCL broadcast for N>1) inside `e2e`"},
        "hbm_gbs": achieved * nr, "clocks": clocks,
        "e2e": {"value": 2.0 * nnz_total / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "wall_ms_per_step": wall / e2e_steps,
                "steps": e2e_steps, "blocks_ms_per_step": [round(b, 4) for b in block_ms], "statistic": "median block of %d steps" % e2e_steps,
                "h2d_bytes_per_step": int(Ncols * 8), "d2h_bytes_per_step": int(dm.M * 8 * nr),
                "path": "spmvb200_spmv_host (pinned host x -> device in pieces, row chunks, y chunks -> pinned host)" if world == 1 else
                        "per-rank H2D of its x slice, rows the other ranks read delivered by peer stores (CUDA IPC) + flag barrier, "
                        "spmvb200_spmv_device, per-rank D2H of its y slice" if pusher is not None else
                        ("per-rank H2D of its x slice, NCCL all-gather, spmvb200_spmv_device, per-rank D2H of its y slice"
                         if args.x_dist == "allgather" and even else
                         "rank0 H2D x, NCCL broadcast, spmvb200_spmv_device, per-rank D2H of its y slice")},
        "gpu_launches": launches, "gpu_launches_e2e": e2e_launch,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(bytes_local),
                     "index_bits": idx_bits, "moved_bytes_per_launch": int(bytes_local - (2 * nnz_total // nr if idx_bits == 16 else 0)),
                     "frac_moved": (bytes_local - (2 * nnz_total // nr if idx_bits == 16 else 0)) / (ms_step * 1e-3) / 1e9 / peak,
                     "note": "per GPU = global algorithmic bytes / n_gpus; global = 12*nnz + 4*M(+1 for CSR) + 8*N + 8*M (DESIGN.md). "
                             "With index_bits 16 the kernel reads 2-byte column offsets (10 B per non-zero): it moves fewer bytes than the "
                             "algorithmic figure it is scored against, so frac may exceed 1; frac_moved = the bytes it really moves / (t * peak)"},
        "parity": parity,
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file)).get(kname)
            # the captures are single-GPU launches of the named workloads: per-GPU traffic of a strong-scaled run is not in the file
            out["roofline"]["traffic"] = tr if (args.workload == "cfg2" or nr == 1) else None
        except Exception:  # noqa: BLE001
            pass
    if rank == 0 and N == 1 and not args.no_cpu:
        try:
            _, _, info = cpu_reference(slab_spec, args.cpu_steps, 2, args.workload)
            out["cpu_baseline"] = info
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        emit(out)
    if pusher is not None:
        pusher.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
