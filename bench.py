#!/usr/bin/env python
"""Headline benchmark of the B200 SpMV engine (contract: the task statement / DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is one SpMV (y = A*x, fp64) over the workload.

  cfg4 (default, every N): BASELINE.json configs[3] -- CSR, random banded 2^25 rows x 32 nnz/row (2^30 nnz, half width 2^15), the
        matrix the metric's "1/2/4/8 GPUs vs host OpenMP CSR" is quoted on; it fits one B200 (13.6 GB).  Row-block partitioned over
        the N GPUs (the reference's own CPU decomposition, spmvRowsBlocksCSR, one block per GPU): STRONG scaling.
  cfg2: BASELINE.json configs[1] -- ELL, 27-point 3-D stencil 128^3 per GPU (weak scaling); at N=1 its line (and a cfg1 line) is
        also measured after the cfg4 run and reported under the extra keys "cfg2" / "cfg1".

`value`  : whole-job GFLOP/s (2*nnz/t) of the iterated step x <- A x with x replicated on every GPU: the SpMV kernel itself
           stores the rows the other GPUs read into THEIR next x (posted NVLink stores from its epilogue) and a flag barrier
           separates iterations -- the exchange is INSIDE the timed region (at N=1 there is nobody to deliver to).  CUDA events on
           the launch stream, barrier + synchronize on both sides, max over ranks.  `kernel_only` = the same kernel without the
           exchange.
`e2e`    : the same metric through the host-buffer entry point, every step moving x up and y down: N=1 spmvb200_spmv_host, N>1
           spmvb200_shard_spmv_host (every rank: its x slice up over its own PCIe link, halo rows to the peers, kernel chunks, its y
           slice down).  Headline: page-locked vectors from spmvb200_host_alloc (the contract's "pinned host memory").  Beside it:
           `registered` = caller-allocated malloc / numpy vectors as the reference's driver has them (src/main.cu:155,181),
           page-locked in place by one spmvb200_host_register call each, `unregistered` = the same vectors left pageable, and the
           bare duplex-copy ceiling of the box for both kinds of page-locked memory.
--impl reference times the reference's own OpenMP CPU implementation (oracle/_ref, compiled from the unmodified sources; the
oracle port if that is absent) on the host cores over the FULL matrix of the same workload, rank 0 only.
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
CFG4_ROWS = 1 << 25
RESET_EVERY = 64  # x <- A x grows ~3.3x per step on cfg4 (32 U(-1,1) entries per row): restart from x0 before fp64 overflows


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload(name, nr, half_width):
    """name -> (description, scaling, synth spec of the GLOBAL matrix for nr ranks, split points, format)"""
    from spmv_openmp_cuda_b200 import synth
    if name == "cfg2":
        rows_per = 128 * 128 * 128
        return ("cfg2: ELL fp64 SpMV, 27-point 3-D stencil 128^3 per GPU (2097152 rows, 55742968 nnz per GPU), "
                "column-major pitched ELL + row-length early exit, kernel cudaSpMVRowsELL", "weak",
                synth.stencil27(128, 128, 128 * nr), [g * rows_per for g in range(nr + 1)], "ell")
    return ("cfg4: CSR fp64 SpMV, random banded 2^25 rows x 32 nnz/row (2^30 nnz, half width w=%d), row-block partitioned over the "
            "GPUs, kernel cudaSpMVRowsCSR (bit-exact kind)" % half_width, "strong",
            synth.banded(CFG4_ROWS, 32, half_width), [g * CFG4_ROWS // nr for g in range(nr + 1)], "csr")


def make_config(wl_name, rows_total, nnz_total, nr, fmt):
    """The `config` object of the JSON line -- built by ONE function for both arms so that they carry identical keys and values."""
    rowmeta = 4 * rows_total if fmt == "ell" else 4 * (rows_total + 1)
    bytes_local = (12 * nnz_total + rowmeta + 16 * rows_total) // nr
    return {"workload": wl_name, "rows_total": int(rows_total), "nnz_total": int(nnz_total), "rows_per_gpu": int(rows_total // nr),
            "parallelism": "row-block x%d" % nr,
            "l2": "no flush: per-GPU inputs (%.0f MB) exceed the 126 MB L2" % (bytes_local / 1e6)}


# ------------------------------------------------------------------------------------------ clocks / link
class ClockSampler:
    """Samples SM clock, throttle reasons and the PCIe link state through NVML while a timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, False, None
        self.link = []
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
                try:
                    self.link.append((nv.nvmlDeviceGetCurrPcieLinkGeneration(self.h), nv.nvmlDeviceGetCurrPcieLinkWidth(self.h),
                                      nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM)))
                except Exception:  # noqa: BLE001
                    pass
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def stop(self):
        self._stop = True
        if self._t:
            self._t.join()
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.link:
            gens, widths, mem = zip(*self.link)
            out["pcie_gen_min_max"] = [int(min(gens)), int(max(gens))]
            out["pcie_width_min_max"] = [int(min(widths)), int(max(widths))]
            out["mem_mhz_min_max"] = [int(min(mem)), int(max(mem))]
        return out


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(spec, row_begin, row_end, steps, warmup, label):
    """Time the reference's OpenMP CPU SpMV (best of its CSR/ELL row kernels) on the host cores over rows [row_begin, row_end) of
    the workload's matrix.  Returns (value GFLOP/s, ms_per_step, info dict)."""
    import oracle
    from spmv_openmp_cuda_b200 import synth

    t0 = time.time()
    mat = synth.host_csr(spec, row_begin, row_end)
    x = synth.host_vector(mat.N)
    cores = oracle.omp_max_threads()
    flops = 2.0 * mat.NZ
    results = {}
    if oracle.ref_available():
        kind = "reference"
        grid = int(min(65535, 4 * cores))  # gridRows caps the parallelism of *Blocks* (SpMV_CSR_OMP.c:70-76)
        rmat = oracle.ref_spmat(mat.M, mat.N, mat.NZ, mat.JA, mat.AS, irp=mat.IRP, rl=mat.RL)
        ell = synth.csr_to_ell_host(mat)  # rows of equal length (cfg4): the ELL arrays ARE the CSR arrays, no copy
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
    nprobe = 3 if mat.NZ < (1 << 29) else 2
    y_ref = None
    for name, m in cands:  # pick the fastest variant on a few probes each
        ts = []
        for _ in range(nprobe):
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
            "sample": "%s, rows [%d, %d) of the matrix (M=%d, nnz=%d), %d timed calls of %s after %d warm-ups; probes (s): %s; setup %.1f s"
                      % (label, row_begin, row_end, mat.M, mat.NZ, steps, best, warmup,
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


def log(msg):
    if os.environ.get("BENCH_VERBOSE"):
        sys.stderr.write("bench[%s] %s\n" % (os.environ.get("RANK", "0"), msg))


def side_line(name, steps, peak):
    """N=1 extra lines: cfg2 (the ELL headline of round 1) and cfg1 (L2-flushed isolated launches and CUDA-graph replay)."""
    import spmv_openmp_cuda_b200 as sp
    from spmv_openmp_cuda_b200 import synth
    if name == "cfg2":
        d_csr = synth.device_csr(synth.stencil27(128))
        dm, kind = d_csr.to_ell(sp.FMT_ELL_COLMAJOR), sp.ELL_ROWS
        d_csr.free()
    else:
        d_csr = synth.device_csr(synth.lap2d(1024))
        dm, kind = d_csr.to_ell(sp.FMT_ELL_COLMAJOR), sp.ELL_ROWS
        d_csr.free()
    x, y = sp.DeviceVector(dm.N), sp.DeviceVector(dm.M)
    synth.device_vector_fill(x.data_ptr(), dm.N)
    sp.time_kernel(kind, dm, x, y, reps=5)
    t = sp.time_kernel(kind, dm, x, y, reps=steps, flush_l2=(name == "cfg1"))
    ms = float(np.mean(t))
    B = dm.algorithmic_bytes
    out = {"workload": name, "kernel": "ell_colmajor_pair_kernel<idx16>" if dm.index_bits == 16 else "ell_colmajor_kernel",
           "rows": dm.M, "nnz": dm.NZ, "launches": steps, "ms_per_step": ms, "ms_min": float(np.min(t)),
           "l2": "flushed between launches (read of a 512 MB buffer)" if name == "cfg1" else "inputs exceed L2",
           "value": 2.0 * dm.NZ / (ms * 1e-3) / 1e9, "unit": UNIT, "algorithmic_bytes_per_launch": int(B),
           "roofline_frac": B / (ms * 1e-3) / 1e9 / peak}
    if name == "cfg1":  # back to back from a CUDA graph: what an iterative solver sees (L2-warm: 84 MB fit the 126 MB L2)
        b = sp.DeviceVector(dm.N)
        iters = 200
        g_ms = sp.iterate(kind, dm, x, b, iters, use_graph=True) / iters
        out["graph_replay_ms_per_step"] = g_ms
        out["graph_replay_roofline_frac"] = B / (g_ms * 1e-3) / 1e9 / peak
        b.free()
    x.free()
    y.free()
    dm.free()
    return out


class NcclFallbackShard:
    """Same interface as distributed.RowBlockShard for boxes where the GPUs cannot map each other's memory (CUDA IPC / peer access
    refused): the exchange is an NCCL all-gather of the y slices into the next x after the SpMV instead of peer stores from its
    epilogue.  Slower (the whole block travels, as a collective after the kernel) but the same arithmetic; the JSON line says which
    exchange ran."""

    def __init__(self, dm, splits, kind, torch, dist, lib, capi, why):
        self.dm, self.kind, self.torch, self.dist, self.lib, self.capi, self.why = dm, kind, torch, dist, lib, capi, why
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.r0, self.r1 = splits[self.rank], splits[self.rank + 1]
        assert len({b - a for a, b in zip(splits[:-1], splits[1:])}) == 1, "the all-gather fallback needs equal slices"
        self.x = [torch.zeros(dm.N, dtype=torch.float64, device="cuda") for _ in range(3)]
        self.y = torch.zeros(dm.M, dtype=torch.float64, device="cuda")
        self.halo_rows = (self.world - 1) * dm.M
        self._one = torch.zeros(1, device="cuda")

    def x_ptr(self, buf):
        return self.x[buf].data_ptr()

    def step(self, src, dst, stream=None):
        self.capi.check(self.lib.spmvb200_spmv_device(self.dm.handle, self.kind, self.x[src].data_ptr(), self.y.data_ptr(), stream), "spmv_device")
        self.dist.all_gather_into_tensor(self.x[dst], self.y)

    def barrier(self, stream=None):
        self.dist.all_reduce(self._one)

    def rows_of(self, buf, a, b):
        return self.x[buf][a:b].cpu().numpy()

    def spmv_host(self, x_slice, y_slice):
        torch = self.torch
        xs = x_slice if torch.is_tensor(x_slice) else torch.from_numpy(x_slice)
        ys = y_slice if torch.is_tensor(y_slice) else torch.from_numpy(y_slice)
        self.x[1][self.r0:self.r1].copy_(xs, non_blocking=True)
        self.dist.all_gather_into_tensor(self.x[2], self.x[1][self.r0:self.r1])
        self.capi.check(self.lib.spmvb200_spmv_device(self.dm.handle, self.kind, self.x[2].data_ptr(), self.y.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream), "spmv_device")
        ys.copy_(self.y, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def close(self):
        pass


def link_ceiling(nbytes_up, nbytes_down, world, dist, torch, reps=10, bufs=None):
    """Bare duplex copy of this rank's per-step host<->device bytes (pinned, two streams, nothing else): the floor any host-buffer
    step has on this box.  All ranks copy at once (they share the host's memory / PCIe root); max over ranks.  bufs = (x, y) numpy
    arrays that are page-locked already: the same measurement on THAT memory (4 KB pages registered in place are slower)."""
    if bufs is None:
        hu = torch.empty(max(nbytes_up, 8) // 8, dtype=torch.float64).pin_memory()
        hd = torch.empty(max(nbytes_down, 8) // 8, dtype=torch.float64).pin_memory()
    else:
        hu, hd = torch.from_numpy(bufs[0]), torch.from_numpy(bufs[1])
        assert hu.is_pinned() and hd.is_pinned()
    du, dd = torch.empty_like(hu, device="cuda"), torch.empty_like(hd, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    best = {"duplex": 1e30, "up": 1e30, "down": 1e30}
    for mode in ("duplex", "up", "down"):
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record(s1)
            s2.wait_event(e0)
            if mode != "down":
                with torch.cuda.stream(s1):
                    du.copy_(hu, non_blocking=True)
            if mode != "up":
                with torch.cuda.stream(s2):
                    hd.copy_(dd, non_blocking=True)
            e1.record(s1)
            e2.record(s2)
            torch.cuda.synchronize()
            best[mode] = min(best[mode], max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
    t = torch.tensor([best["duplex"], best["up"], best["down"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg2", "cfg4"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of one block of the end-to-end loop (default: min(steps, 50))")
    ap.add_argument("--e2e-blocks", type=int, default=5, help="blocks of the end-to-end loop; the median block is reported")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-side", action="store_true", help="N=1: skip the cfg2 / cfg1 side lines")
    ap.add_argument("--half-width", type=int, default=1 << 15, help="cfg4: band half width w")
    ap.add_argument("--ref-rows", type=int, default=0, help="reference arm: rows of the matrix to time (default: all)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N = args.gpus

    from spmv_openmp_cuda_b200 import synth

    # ---------------------------------------------------------------- reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        wl_name, scaling, spec, splits, fmt = workload(args.workload, N, args.half_width)
        steps, warm = min(args.steps, 30), min(args.warmup, 5)
        Mtot = splits[-1]
        rows = args.ref_rows or Mtot  # the FULL matrix the GPU arm runs (cfg4: 2^30 nnz, ~18 GB of host arrays)
        value, ms, info = cpu_reference(spec, 0, rows, steps, warm, args.workload)
        emit({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": steps, "warmup": warm,
              "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": make_config(wl_name, Mtot, info_nnz(info) if rows == Mtot else -1, N, fmt),
              "reference_note": "the reference's OpenMP CPU implementation on %d host threads, rows [0, %d) of the matrix per step%s"
                                % (info["cores"], rows, "" if rows == Mtot else " (a SAMPLE: --ref-rows)"),
              "cpu_baseline": info,
              "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return 0

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist

    import spmv_openmp_cuda_b200 as sp
    from spmv_openmp_cuda_b200 import capi
    from spmv_openmp_cuda_b200.distributed import RowBlockShard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    lib = capi.lib()
    capi.check(lib.spmvb200_set_device(local_rank), "set_device")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == N or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    nr = world  # ranks actually running
    peak, peak_src = measured_peak()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allmin_flag(ok):
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    # ---- this rank's row block, built on its GPU
    wl_name, scaling, spec, splits, fmt = workload(args.workload, nr, args.half_width)
    r0, r1 = splits[rank], splits[rank + 1]
    t_setup = time.time()
    d_csr = synth.device_csr(spec, r0, r1)
    col_range = d_csr.col_range
    if fmt == "ell":
        dm = d_csr.to_ell(sp.FMT_ELL_COLMAJOR)
        d_csr.free()
        kind = sp.ELL_ROWS
    else:
        dm, kind = d_csr, sp.CSR_ROWS
    Ncols, Mloc = dm.N, dm.M
    nnz_total = int(allsum(torch, dist, world, dm.NZ))
    rows_total = splits[-1]
    # algorithmic bytes of the GLOBAL SpMV, split evenly: x is counted once for the whole job
    rowmeta = 4 * rows_total if fmt == "ell" else 4 * (rows_total + 1)
    bytes_local = (12 * nnz_total + rowmeta + 8 * Ncols + 8 * rows_total) // nr

    stream = torch.cuda.current_stream().cuda_stream
    if world > 1:  # a stream of our own: the legacy default stream synchronises implicitly with every other blocking stream
        own = torch.cuda.Stream()
        torch.cuda.set_stream(own)
        stream = own.cuda_stream
    exchange = "peer stores from the SpMV epilogue (CUDA IPC) + flag barrier" if world > 1 else "none (one GPU)"
    try:
        if os.environ.get("BENCH_FORCE_NCCL_EXCHANGE") and world > 1:
            raise sp.SpmvB200Error("forced by BENCH_FORCE_NCCL_EXCHANGE")
        shard = RowBlockShard(dm, splits, kind, nbuf=3, col_range=col_range)
    except sp.SpmvB200Error as e:  # every rank raises together (RowBlockShard agrees on the outcome before returning)
        if world == 1:
            raise
        shard = NcclFallbackShard(dm, splits, kind, torch, dist, lib, capi, repr(e))
        exchange = "NCCL all-gather of the y slices after the SpMV (peer mapping unavailable: %s)" % repr(e)[:200]
        if rank == 0:
            sys.stderr.write("bench.py: %s\n" % exchange)
    X0 = shard.x_ptr(0)
    synth.device_vector_fill(X0, Ncols)  # replicated x0: every rank generates the whole vector (pure function of the index)
    capi.check(lib.spmvb200_tune(dm.handle, kind, X0, shard.x_ptr(1), stream), "tune")  # first-use pick, outside every timed region
    choice = dm.exact_choice
    kname = {"xwindow": "xwin_kernel", "sell": "sell_kernel", "stream": "csr_stream_kernel",
             "ell": "ell_colmajor_pair_kernel<idx16>" if dm.index_bits == 16 else "ell_colmajor_kernel"}.get(choice, choice)
    idx_bits = 16 if choice == "xwindow" else dm.index_bits
    log("setup %.1f s, pick %s, halo rows %d" % (time.time() - t_setup, choice, shard.halo_rows))

    seq = {"i": 0, "cur": 0}

    def step():
        """x <- A x; every RESET_EVERY steps the source is the pristine x0 again (a pointer choice, no copy)"""
        src = 0 if seq["i"] % RESET_EVERY == 0 else seq["cur"]
        dst = 2 if src == 1 else 1
        shard.step(src, dst, stream)
        seq["cur"], seq["i"] = dst, seq["i"] + 1

    def kernel_step():
        capi.check(lib.spmvb200_spmv_device(dm.handle, kind, X0, shard.x_ptr(2), stream), "spmv_device")

    def timed(fn, steps, warm):
        """W untimed warm-up steps, barrier + synchronize, then EXACTLY `steps` steps between two CUDA events on the launch stream,
        barrier + synchronize again; max over ranks.  Nothing but event records sits between the barrier and the first timed step
        (the NVML sampler thread and the events exist before it), and the ranks' STREAMS are aligned by one flag barrier across
        the GPUs right before the first event: with a step that exchanges data every step, host-side skew between the ranks'
        start would otherwise be paid by the early ranks as waiting inside their timed region (measured at 8 GPUs, 20 steps of
        0.32 ms: +50 % from NVML initialisation alone sitting after the barrier)."""
        sampler = ClockSampler(local_rank).start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(warm):
            fn()
        sync_all()
        l0 = lib.spmvb200_launch_count()
        if world > 1:
            shard.barrier(stream)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        sync_all()
        clocks = sampler.stop()
        return allmax(ev0.elapsed_time(ev1)) / steps, int(lib.spmvb200_launch_count() - l0) - (1 if world > 1 else 0), clocks

    ms_step, launches, clocks = timed(step, args.steps, args.warmup)
    ms_kernel = ms_step
    if world > 1:
        ms_kernel, _, _ = timed(kernel_step, args.steps, 3)
    gflops = 2.0 * nnz_total / (ms_step * 1e-3) / 1e9
    achieved = bytes_local / (ms_step * 1e-3) / 1e9  # per GPU: the dominant kernel's launch on one GPU

    # ---- parity of what was just timed: two iterations from x0, rows on BOTH sides of every slab boundary (the rows whose columns
    # were delivered by a peer) and a block in the middle, against the oracle's serial SpMV -- every rank checks its own rows
    import oracle
    xh0 = synth.host_vector(Ncols)
    shard.step(0, 1, stream)
    shard.step(1, 2, stream)
    sync_all()
    CH = 256
    blocks = sorted({(max(r0, a), min(r1, a + CH)) for a in (r0, r1 - CH, (r0 + r1) // 2) if r1 - r0 > 0})
    par_ok, rows_checked, x1_full = True, 0, np.full(Ncols, np.nan)
    for a, b in blocks:
        hb = synth.host_csr(spec, a, b)
        cols = hb.JA
        ca, cb = int(cols.min()), int(cols.max()) + 1
        h1 = synth.host_csr(spec, ca, cb)  # the rows of A that produce the x1 entries these rows read (square matrix)
        x1_full[ca:cb] = oracle.sgemv_serial(h1.IRP, h1.JA, h1.AS, xh0)
        want = oracle.sgemv_serial(hb.IRP, hb.JA, hb.AS, x1_full)
        got = shard.rows_of(2, a, b)
        par_ok &= bool(np.array_equal(want, got))
        rows_checked += b - a
    parity = {"bit_identical": allmin_flag(par_ok), "ranks_checked": nr, "rows_checked_per_rank": int(rows_checked),
              "what": "x2 = A(A x0) through the timed path; per rank: first %d rows, last %d rows (both read x rows delivered by the "
                      "neighbouring GPU's kernel) and %d rows in the middle, vs the oracle's sgemvSerial" % (CH, CH, CH)}

    # ---- end to end through the host-buffer entry point: caller-allocated pageable buffers, as the reference driver passes them
    e2e_steps = args.e2e_steps or min(args.steps, 50)
    hx = xh0  # numpy (malloc'ed, pageable); page-locked in place below by the one call a driver adds next to its malloc
    hy = np.empty(Mloc, dtype=np.float64)
    hx_slice = hx[r0:r1]

    def host_alloc(n):  # page-locked vector from the library (what a driver swaps its malloc for)
        import ctypes
        p = lib.spmvb200_host_alloc(max(n, 1) * 8)
        if not p:
            raise RuntimeError("spmvb200_host_alloc failed: " + lib.spmvb200_last_error().decode(errors="replace"))
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(n,)), p
    hpx, hpx_p = host_alloc(r1 - r0 if world > 1 else Ncols)
    hpy, hpy_p = host_alloc(Mloc)
    hpx[:] = hx_slice if world > 1 else hx

    def e2e_step_of(xbuf, ybuf):
        if world == 1:
            return lambda: capi.check(lib.spmvb200_spmv_host(dm.handle, kind, capi.ptr(xbuf), capi.ptr(ybuf), None), "spmv_host")
        return lambda: shard.spmv_host(xbuf, ybuf, timed=False) if hasattr(shard, "_h") else shard.spmv_host(xbuf, ybuf)

    def e2e_timed(fn, blocks_n):
        sync_all()
        t_first = time.perf_counter()
        fn()
        first_ms = (time.perf_counter() - t_first) * 1e3
        for _ in range(max(3, args.warmup)):
            fn()
        block_ms = []
        l0 = int(lib.spmvb200_launch_count())
        sampler = ClockSampler(local_rank).start()
        for _ in range(max(1, blocks_n)):
            sync_all()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                fn()
            torch.cuda.synchronize()
            block_ms.append(allmax((time.perf_counter() - t0) * 1e3) / e2e_steps)
        link = sampler.stop()
        return block_ms, int(lib.spmvb200_launch_count()) - l0, first_ms, link

    xbuf_pg = hx_slice if world > 1 else hx
    fn_pg = e2e_step_of(xbuf_pg, hy)
    sync_all()
    t_first = time.perf_counter()
    fn_pg()  # first call: builds the pipeline plan, plain pageable copies
    first_ms = (time.perf_counter() - t_first) * 1e3
    unreg = []
    for _ in range(3):  # the same call with the buffers left pageable (driver staging copies, no overlap)
        sync_all()
        t0 = time.perf_counter()
        fn_pg()
        unreg.append(allmax((time.perf_counter() - t0) * 1e3))
    # headline: page-locked vectors (the contract's "pinned host memory"), allocated through the library
    blocks_pin, e2e_launch, _, link = e2e_timed(e2e_step_of(hpx, hpy), args.e2e_blocks)
    e2e_ms = float(np.median(blocks_pin))
    # malloc'ed vectors page-locked in place (the other one-line change; 4 KB pages)
    capi.check(lib.spmvb200_host_register(capi.ptr(xbuf_pg), xbuf_pg.nbytes), "host_register")  # after main.cu:155,181
    capi.check(lib.spmvb200_host_register(capi.ptr(hy), hy.nbytes), "host_register")
    hy.fill(np.nan)
    blocks_pg, _, _, _ = e2e_timed(fn_pg, max(1, args.e2e_blocks // 2))
    y_check = hy.copy()
    # e2e parity: y = A x0 rows at both ends of the slice (they read halo rows uploaded by the NEIGHBOUR and pushed here)
    e2e_ok = True
    for a, b in blocks:
        hb = synth.host_csr(spec, a, b)
        e2e_ok &= bool(np.array_equal(oracle.sgemv_serial(hb.IRP, hb.JA, hb.AS, xh0), y_check[a - r0:b - r0]))
    e2e_ok &= bool(np.array_equal(y_check, hpy))
    parity["e2e_bit_identical"] = allmin_flag(e2e_ok)
    up_b, down_b = (r1 - r0 if world > 1 else Ncols) * 8, Mloc * 8
    ceil_ms = link_ceiling(up_b, down_b, world, dist, torch)
    ceil_reg_ms = link_ceiling(up_b, down_b, world, dist, torch, reps=5, bufs=(xbuf_pg, hy))
    capi.check(lib.spmvb200_host_unregister(None), "host_unregister")  # before the numpy buffers are freed

    out = {
        "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": make_config(wl_name, rows_total, nnz_total, nr, fmt),
        "step": "x <- A x, x replicated: the SpMV kernel stores the %d rows per GPU that other GPUs read into their next x "
                "(peer stores over NVLink) + flag barrier, inside the timed region; source reset to x0 every %d steps"
                % (shard.halo_rows, RESET_EVERY) if world > 1 else
                "x <- A x on one GPU (ping-pong vectors; source reset to x0 every %d steps)" % RESET_EVERY,
        "kernel_only": {"ms_per_step": ms_kernel, "value": 2.0 * nnz_total / (ms_kernel * 1e-3) / 1e9, "unit": UNIT,
                        "exchange_overhead_frac": ms_step / ms_kernel - 1.0},
        "exchange": exchange,
        "nvlink_bytes_per_step": int(allsum(torch, dist, world, shard.halo_rows * 8)),
        "hbm_gbs": achieved * nr, "clocks": clocks,
        "e2e": {"value": 2.0 * nnz_total / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "blocks_ms_per_step": [round(b, 4) for b in blocks_pin], "statistic": "median block of %d steps, host clock, max over ranks" % e2e_steps,
                "h2d_bytes_per_step": int(Ncols * 8), "d2h_bytes_per_step": int(rows_total * 8),
                "buffers": "page-locked x and y from spmvb200_host_alloc (cudaHostAlloc): the driver's malloc / free of its two vectors swapped "
                           "for the library's calls (INTEGRATION.md)",
                "first_call_ms": first_ms,
                "registered": {"ms_per_step": float(np.median(blocks_pg)), "blocks_ms_per_step": [round(b, 4) for b in blocks_pg],
                               "buffers": "caller-allocated pageable (numpy / malloc) x and y, page-locked IN PLACE by one "
                                          "spmvb200_host_register call each (4 KB pages)",
                               "link_ceiling_duplex_ms": ceil_reg_ms[0], "frac_achieved": ceil_reg_ms[0] / float(np.median(blocks_pg))},
                "unregistered": {"ms_per_step": float(np.median(unreg)),
                                 "buffers": "the same malloc'ed vectors left pageable: the CUDA driver's staging copies, no overlap"},
                "link_ceiling": {"duplex_ms": ceil_ms[0], "up_only_ms": ceil_ms[1], "down_only_ms": ceil_ms[2],
                                 "what": "bare pinned cudaMemcpyAsync of the same per-rank bytes, up and down at once on two streams, "
                                         "all ranks together, best of 10, max over ranks",
                                 "aggregate_gbs": (Ncols + rows_total) * 8 / (ceil_ms[0] * 1e-3) / 1e9,
                                 "frac_achieved": ceil_ms[0] / e2e_ms},
                "link_state": link,
                "path": "spmvb200_spmv_host: x host -> device in pieces, row chunks of the x-window kernel as their pieces land, y chunks "
                        "-> host while later chunks compute" if world == 1 else
                        "spmvb200_shard_spmv_host on every rank: its x slice up over its own PCIe link (halo rows first, delivered to the "
                        "peers by peer stores + flag barrier while the rest uploads), row chunks, its y slice down"},
        "gpu_launches": launches, "gpu_launches_e2e": e2e_launch,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(bytes_local),
                     "index_bits": idx_bits, "moved_bytes_per_launch": int(bytes_local - (2 * nnz_total // nr if idx_bits == 16 else 0)),
                     "frac_moved": (bytes_local - (2 * nnz_total // nr if idx_bits == 16 else 0)) / (ms_step * 1e-3) / 1e9 / peak,
                     "note": "per GPU = global algorithmic bytes / n_gpus; global = 12*nnz + 4*M(+1 for CSR) + 8*N + 8*M (DESIGN.md). "
                             "With index_bits 16 the kernel reads 2-byte column ids (10 B per non-zero): it moves fewer bytes than the "
                             "algorithmic figure it is scored against; frac_moved = the bytes it really moves / (t * peak)"},
        "parity": parity,
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file)).get(kname + "@" + args.workload)
            # {"bytes": dram read+write per launch, "capture": file under profiles/, "git": commit of the captured build}: a single-GPU
            # capture of the named workload -- per-GPU traffic of a partitioned run is not in the file
            if tr and nr == 1:
                out["roofline"]["traffic"] = tr["bytes"]
                out["roofline"]["traffic_source"] = {k: tr[k] for k in tr if k != "bytes"}
        except Exception:  # noqa: BLE001
            pass
    shard.close()
    dm.free()
    del hpx, hpy
    lib.spmvb200_host_free(hpx_p)
    lib.spmvb200_host_free(hpy_p)
    if rank == 0 and N == 1 and not args.no_side:
        for name in ("cfg2", "cfg1"):
            try:
                out[name] = side_line(name, 50, peak)
            except Exception as e:  # noqa: BLE001
                out[name] = {"error": repr(e)}
    if rank == 0 and N == 1 and not args.no_cpu:
        try:
            # bounded sample: the first 2^23 rows of the same matrix (2^28 nnz, ~4.5 GB of host arrays); --impl reference times all of it
            sample_rows = min(rows_total, 1 << 23)
            _, _, info = cpu_reference(spec, 0, sample_rows, args.cpu_steps, 2, args.workload)
            out["cpu_baseline"] = info
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def allsum(torch, dist, world, v):
    t = torch.tensor([int(v)], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    return int(t.item())


def info_nnz(info):
    """nnz out of the cpu_reference sample string (kept there for the human reader)"""
    import re
    m = re.search(r"nnz=(\d+)", info["sample"])
    return int(m.group(1)) if m else 0


if __name__ == "__main__":
    sys.exit(main())
