#!/usr/bin/env python
"""Benchmark of the B200-native fp64 CSR SpMV engine on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline workload (config C2 of BASELINE.json): 2D 5-point Laplacian on a 4096 x 4096 grid (16.8M rows, 83.9M nnz),
fp64, alpha = beta = 1 like the reference harness (benchmark/main.cpp:101-102). One step = one SpMV over the whole
matrix through the C ABI (spmv_b200_execute). With N > 1 ranks (torchrun, one rank per GPU) the grid grows to
(4096*N) x 4096, rows are cut into N nnz-balanced contiguous shards, x is replicated (one-shot SpMV needs no
exchange): weak scaling. The iterated configuration (C5, 27-point 384^3 power loop with the NCCL / halo exchange of x)
is measured by the same run and reported under "iterated".

Prints ONE JSON line on rank 0. Keys are described in DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

C2_GRID = 4096
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent
NOMINAL_HBM_GBS = 8000.0   # BASELINE.json north_star


def alg_bytes(m, n, nnz):
    return 12 * nnz + 4 * (m + 1) + 8 * n + 16 * m


def harness_bytes(m, nnz):
    # the reference harness' own model (benchmark/utils/statistics_logger.cpp:43): omits the read of x
    return 8 * (2 * m + nnz) + 4 * (m + 1 + nnz)


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _reasons(self):
        nv = self.nv
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            try:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            except Exception:
                return
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        for k, bit in names.items():
            if mask & bit:
                self.reasons.add(k)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self._reasons()
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU SpMV (cli/verification.cpp:56-66) on the box's host cores
# ----------------------------------------------------------------------------------------------------------------
def host_c2_rows(rows):
    """First `rows` rows of the C2 matrix on the host (numpy restatement of the device generator)."""
    from spmv_acc_b200 import synth
    return synth.stencil2d_numpy(C2_GRID, 0, rows)


def time_reference_spmv(csr, x, y, reps):
    import oracle
    import ctypes as C
    lib = oracle.oracle._ref_lib() if oracle.have_ref() else oracle.oracle._port_lib()
    fn = lib.ref_host_spmv_axpby if oracle.have_ref() else lib.port_host_spmv_axpby
    P = oracle.oracle._p
    args = (C.c_double(1.0), C.c_double(1.0), P(csr.val, C.c_double), P(csr.rowptr, C.c_int), P(csr.col, C.c_int),
            C.c_int(csr.rows), C.c_int(csr.cols), C.c_int(csr.nnz), P(x, C.c_double), P(y, C.c_double))
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(*args)
        times.append(time.perf_counter() - t0)
    return times, ("reference" if oracle.have_ref() else "port")


def run_reference(args, rank):
    if rank != 0:
        return
    from spmv_acc_b200 import synth
    total_steps = args.steps + args.warmup
    budget_s = 90.0
    # calibrate on 1M rows, then size the per-step sample so the whole run fits the budget
    probe = host_c2_rows(1 << 20)
    x = synth.vector_numpy(C2_GRID * C2_GRID, 2)
    y = synth.vector_numpy(probe.rows, 3)
    t, kind = time_reference_spmv(probe, x, y, 2)
    rows_per_s = probe.rows / min(t)
    rows = int(min(C2_GRID * C2_GRID, max(1 << 16, rows_per_s * budget_s / total_steps)))
    csr = probe if rows == probe.rows else host_c2_rows(rows)
    y = synth.vector_numpy(csr.rows, 3)
    time_reference_spmv(csr, x, y, args.warmup)
    times, kind = time_reference_spmv(csr, x, y, args.steps)
    sec = float(np.mean(times))
    gflops = 2.0 * csr.nnz / sec / 1e9
    sample = f"first {csr.rows} of {C2_GRID * C2_GRID} rows of the C2 matrix per step ({csr.nnz} nnz), serial host_spmv"
    line = {
        "impl": "reference", "metric": "fp64 CSR SpMV GFLOP/s (2*nnz/t)", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 2D 5-point Laplacian 4096x4096 grid, fp64, alpha=beta=1", "sample": sample},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(),
                         "note": "host_spmv (cli/verification.cpp:56-66) is serial: 1 thread is all it can use"},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "effective_gbs": alg_bytes(csr.rows, csr.cols, csr.nnz) / sec / 1e9,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def sync_all(torch, dist, world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(torch, dist, world, value):
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(torch, dist, world, value):
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def build_headline_shard(torch, rank, world):
    """Rows [lo, hi) of the (4096*world) x 4096 grid Laplacian owned by this rank (nnz-balanced)."""
    from spmv_acc_b200 import shard_bounds, synth
    NY = C2_GRID * world
    if world == 1:
        lo, hi = 0, NY * C2_GRID
        bounds = np.array([lo, hi], dtype=np.int64)
    else:
        counts = synth.stencil_row_counts_device("stencil2d", C2_GRID, NY)
        rowptr = synth._rowptr_from_counts_device(counts)
        del counts
        bounds = shard_bounds(rowptr, NY * C2_GRID, world).astype(np.int64)
        del rowptr
        torch.cuda.empty_cache()
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    csr = synth.stencil2d_device(C2_GRID, lo, hi, NY=NY)
    return csr, lo, hi, bounds


def time_steps(torch, dist, world, fn, steps, warmup, sampler=None):
    for _ in range(warmup):
        fn()
    sync_all(torch, dist, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.start()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.synchronize()
    if sampler:
        sampler.stop()
    sync_all(torch, dist, world)
    return max_over_ranks(torch, dist, world, e0.elapsed_time(e1))  # ms for `steps` steps


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from spmv_acc_b200 import CsrDesc, HostMatrix, SpmvPlan, make_options, synth, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    csr, lo, hi, bounds = build_headline_shard(torch, rank, world)
    n_global = C2_GRID * C2_GRID * world
    opt = make_options(args.tile, args.short_max, args.medium_max, args.vec_div, args.flags)
    t0 = time.perf_counter()
    plan = SpmvPlan(CsrDesc(csr.rows, csr.cols, csr.nnz, csr.rowptr, csr.col, csr.val), opt)
    torch.cuda.synchronize()
    plan_first_ms = (time.perf_counter() - t0) * 1e3  # includes loading the CUDA module (first use of the library)
    plan.destroy()
    t0 = time.perf_counter()
    plan = SpmvPlan(CsrDesc(csr.rows, csr.cols, csr.nnz, csr.rowptr, csr.col, csr.val), opt)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    info = plan.info()
    x = synth.vector_device(n_global, 2)
    y = synth.vector_device(csr.rows, 3 + rank)
    alpha = beta = 1.0

    def step():
        plan.execute(alpha, beta, x, y)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    total_ms = time_steps(torch, dist, world, step, args.steps, args.warmup, sampler)
    ms_per_step = total_ms / args.steps
    nnz_total = sum_over_ranks(torch, dist, world, float(csr.nnz))
    rows_total = sum_over_ranks(torch, dist, world, float(csr.rows))
    gflops = 2.0 * nnz_total / (ms_per_step * 1e-3) / 1e9
    # roofline of the dominant kernel (rank 0's launch): algorithmic bytes of this rank's shard per launch. With
    # several ranks x is replicated but a shard only reads the entries its columns reference (own rows + halo):
    # count those (4096-entry blocks from the analysis) instead of the whole replicated vector.
    if world == 1:
        n_ref = csr.cols
    else:
        from spmv_acc_b200 import col_block_bitmap
        n_ref = min(csr.cols, int(col_block_bitmap(csr.col, csr.nnz, csr.cols, 12).sum()) * 4096)
    b_alg = alg_bytes(csr.rows, n_ref, csr.nnz)
    peak, peak_src = measured_peak()
    achieved = b_alg / (ms_per_step * 1e-3) / 1e9
    traffic = None
    prof = ROOT / "profiles" / "ncu_c2_summary.json"
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    dominant = "k_spmv_warp (direct form)" if info.direct else {
        0: "k_spmv_rows<TMA, SHORT>", 1: "k_spmv_rows<TMA, MEDIUM>", 2: "k_spmv_mixed<TMA>"}[
        int(np.argmax(list(info.tiles_per_kind)))]

    line = {
        "metric": "fp64 CSR SpMV GFLOP/s (2*nnz/t)", "value": gflops, "unit": "GFLOP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"C2: 2D 5-point Laplacian, {C2_GRID * world}x{C2_GRID} grid "
                        f"({int(rows_total)} rows, {int(nnz_total)} nnz), fp64, int32 indices, alpha=beta=1",
            "sharding": "nnz-balanced contiguous row shards, x replicated, no exchange (one-shot SpMV)",
            "cache": "inputs larger than L2 (>= 1.0 GB streamed per step vs 126 MB L2); no flush between steps",
            "tile_nnz": info.tile_nnz, "uses_tma": bool(info.uses_tma),
            "tiles_per_kind": list(info.tiles_per_kind), "split_rows": info.nsplit_rows,
        },
        "effective_gbs": alg_bytes(int(rows_total), n_global, int(nnz_total)) / (ms_per_step * 1e-3) / 1e9 / 1.0,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": dominant, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": b_alg, "frac_of_nominal_8TBs": achieved / NOMINAL_HBM_GBS,
            "harness_model_gbs": harness_bytes(csr.rows, csr.nnz) / (ms_per_step * 1e-3) / 1e9,
        },
        "gpu_launches": args.steps * info.launches_per_execute,
        "plan_create_ms": plan_ms, "plan_create_first_call_ms": plan_first_ms,
        "clocks": sampler.summary() if sampler else None,
    }

    # ---- context baseline: cuSPARSE on the same device buffers (rank 0 only, not the product path) ----
    if rank == 0 and not args.no_context:
        line["context"] = run_cusparse(torch, csr, x, y, args)

    # ---- e2e: the reference-facing host-buffer call (H2D x, y0; SpMV; D2H y) every step ----
    if not args.no_e2e:
        line["e2e"] = run_e2e(torch, dist, world, csr, x, n_global, nnz_total, args)

    # ---- CPU baseline: the reference's own host_spmv on the box's host cores (rank 0, N = 1) ----
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = run_cpu_baseline(torch, csr, x, args)

    plan.destroy()
    del csr, x, y
    torch.cuda.empty_cache()

    # ---- the other one-shot configurations of BASELINE.json (C3 uniform-random, C4 R-MAT): same timing; with N > 1
    # ranks every rank multiplies its nnz-balanced row shard against the replicated x (strong scaling, no exchange) ----
    if not args.no_other_configs:
        other = run_other_configs(torch, dist, rank, world, args, peak)
        if rank == 0:
            line["other_configs"] = other

    # ---- iterated configuration C5 (power loop with exchange of x) ----
    if not args.no_iterated:
        try:
            from spmv_acc_b200.sharded import bench_power_loop
            line["iterated"] = bench_power_loop(args.iter_grid, args.iters, exchange=args.exchange,
                                                overlap=not args.no_overlap, graph=args.graph,
                                                fused=not args.no_fused)
        except Exception as e:  # the headline number must survive a failure of the secondary measurement
            line["iterated"] = {"error": f"{type(e).__name__}: {e}"}

    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cusparse(torch, csr, x, y, args):
    import ctypes as C
    from spmv_acc_b200 import _lib
    out = {}
    try:
        X = _lib.ctx()
        yy = y.clone()
        for alg, name in ((0, "cusparse_alg_default"), (2, "cusparse_csr_alg2")):
            h = C.c_void_p()
            rc = X.spmv_b200_ctx_cusparse_create(C.byref(h), csr.rows, csr.cols, csr.nnz, csr.rowptr.data_ptr(),
                                                 csr.col.data_ptr(), csr.val.data_ptr(), x.data_ptr(), yy.data_ptr(),
                                                 alg)
            if rc != 0:
                out[name] = {"error": rc}
                continue
            stream = torch.cuda.current_stream().cuda_stream
            for _ in range(10):
                X.spmv_b200_ctx_cusparse_spmv(h, 1.0, 1.0, stream)
            torch.cuda.synchronize()
            reps = max(20, min(args.steps, 200))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                X.spmv_b200_ctx_cusparse_spmv(h, 1.0, 1.0, stream)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[name] = {"ms": ms, "gflops": 2.0 * csr.nnz / (ms * 1e-3) / 1e9,
                         "effective_gbs": alg_bytes(csr.rows, csr.cols, csr.nnz) / (ms * 1e-3) / 1e9}
            X.spmv_b200_ctx_cusparse_destroy(h)
        # cub::DeviceSpmv::CsrMV computes y = A*x (no alpha / beta), like the reference's adapter benchmark_cub.hpp:25
        h = C.c_void_p()
        rc = X.spmv_b200_ctx_cub_create(C.byref(h), csr.rows, csr.cols, csr.nnz, csr.rowptr.data_ptr(),
                                        csr.col.data_ptr(), csr.val.data_ptr(), x.data_ptr(), yy.data_ptr())
        if rc != 0:
            out["cub_device_spmv"] = {"error": rc}
        else:
            stream = torch.cuda.current_stream().cuda_stream
            call = lambda: X.spmv_b200_ctx_cub_spmv(h, csr.rows, csr.cols, csr.nnz, csr.rowptr.data_ptr(),  # noqa: E731
                                                    csr.col.data_ptr(), csr.val.data_ptr(), x.data_ptr(),
                                                    yy.data_ptr(), stream)
            for _ in range(5):
                call()
            torch.cuda.synchronize()
            reps = max(20, min(args.steps, 200))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                call()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out["cub_device_spmv"] = {"ms": ms, "gflops": 2.0 * csr.nnz / (ms * 1e-3) / 1e9,
                                      "effective_gbs": alg_bytes(csr.rows, csr.cols, csr.nnz) / (ms * 1e-3) / 1e9,
                                      "note": "y = A*x only (beta = 0, alpha = 1)"}
            X.spmv_b200_ctx_cub_destroy(h)
    except Exception as e:
        out["error"] = f"{type(e).__name__}: {e}"
    return out


def run_other_configs(torch, dist, rank, world, args, peak):
    """C3 and C4 of BASELINE.json, device resident, default plan options: ms / GFLOP/s / fraction of the HBM roofline,
    next to cuSPARSE and CUB on the same buffers (N = 1). With N > 1 ranks every rank generates the same matrix, takes
    the nnz-balanced row shard `spmv_b200_shard_bounds` gives it as a window of the row pointers (no copy) and
    multiplies it against the replicated x: strong scaling of the one-shot SpMV, no exchange. Not the headline;
    reported so that one run covers every one-shot config."""
    from spmv_acc_b200 import CsrDesc, SpmvPlan, shard_bounds, synth
    out = {}
    makers = {
        "C3 uniform-random 1e7 x 1e7, 32 nnz/row": lambda: synth.uniform_device(10_000_000, 10_000_000, 32, seed=1),
        "C4 R-MAT 2^24 rows, 2^28 nnz": lambda: synth.rmat_device(24, 16, seed=1),
    }
    for name, make in makers.items():
        try:
            csr = make()
            lo, hi = 0, csr.rows
            if world > 1:
                b = shard_bounds(csr.rowptr, csr.rows, world)
                lo, hi = int(b[rank]), int(b[rank + 1])
            nnz_local = int(csr.rowptr[hi].item()) - int(csr.rowptr[lo].item())
            plan = SpmvPlan(CsrDesc(hi - lo, csr.cols, nnz_local, csr.rowptr[lo:hi + 1], csr.col, csr.val))
            info = plan.info()
            x = synth.vector_device(csr.cols, 2)
            y = synth.vector_device(hi - lo, 3 + rank)
            for _ in range(5):
                plan.execute(1.0, 1.0, x, y)
            sync_all(torch, dist, world)
            reps = 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                plan.execute(1.0, 1.0, x, y)
            e1.record()
            e1.synchronize()
            sync_all(torch, dist, world)
            ms_local = e0.elapsed_time(e1) / reps
            ms = max_over_ranks(torch, dist, world, ms_local)
            b_local = alg_bytes(hi - lo, csr.cols, nnz_local)  # per rank: its rows, its non-zeros, the whole x
            gbs = b_local / (ms_local * 1e-3) / 1e9
            rec = {"ms": ms, "gflops": 2.0 * csr.nnz / (ms * 1e-3) / 1e9, "n_gpus": world,
                   "scaling": "strong" if world > 1 else "single GPU",
                   "effective_gbs_rank0": gbs, "frac_of_measured_peak_rank0": gbs / peak,
                   "frac_of_nominal_8TBs_rank0": gbs / NOMINAL_HBM_GBS,
                   "form": "direct (warp per row block, no shared memory)" if info.direct else "tiled (TMA)",
                   "tile_nnz": info.tile_nnz, "tiles_per_kind_rank0": list(info.tiles_per_kind),
                   "split_rows_rank0": info.nsplit_rows, "launches_per_spmv": info.launches_per_execute,
                   "rows_rank0": hi - lo, "nnz_rank0": nnz_local,
                   "bound": "x gathers (L1 lines in flight), see DESIGN.md §3.5"}
            if world == 1:
                rec.update({"effective_gbs": gbs, "frac_of_measured_peak": gbs / peak,
                            "frac_of_nominal_8TBs": gbs / NOMINAL_HBM_GBS})
                if not args.no_context:
                    rec["context"] = run_cusparse(torch, csr, x, y, args)
            out[name] = rec
            plan.destroy()
            del csr, x, y
            torch.cuda.empty_cache()
        except Exception as e:
            # (a failure is the same on every rank — same code, same matrix — so the ranks stay in step)
            out[name] = {"error": f"{type(e).__name__}: {e}"}
    return out


def run_e2e(torch, dist, world, csr, x, n_global, nnz_total, args):
    from spmv_acc_b200 import HostMatrix, synth
    h = synth.to_host(csr)
    t0 = time.perf_counter()
    hm = HostMatrix(h.rows, h.cols, h.rowptr, h.col, h.val)
    cold_s = time.perf_counter() - t0
    hx = torch.empty(n_global, dtype=torch.float64, pin_memory=True)
    hy = torch.empty(h.rows, dtype=torch.float64, pin_memory=True)
    hx.copy_(x)
    hy.zero_()
    steps = max(3, min(args.steps, args.e2e_steps))
    for _ in range(3):
        hm.spmv(1.0, 1.0, hx, hy)
    sync_all(torch, dist, world)
    t0 = time.perf_counter()
    for _ in range(steps):
        hm.spmv(1.0, 1.0, hx, hy)  # synchronous: returns after the D2H copy of y completed
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    sec = max_over_ranks(torch, dist, world, sec)
    x_lo, x_hi = hm.x_range()  # only the referenced part of x is copied (a row shard reads its rows' columns + halo)
    hm.destroy()
    return {"value": 2.0 * nnz_total / sec / 1e9, "unit": "GFLOP/s", "ms_per_step": sec * 1e3,
            "h2d_bytes_per_step": int(8 * ((x_hi - x_lo) + h.rows)), "d2h_bytes_per_step": int(8 * h.rows),
            "steps": steps, "api": "spmv_b200_hostmat_spmv (matrix resident, x and y0 copied in, y copied out "
                                   "from/to pinned host memory every step; cli/main.cpp:99-118 pattern)",
            "cold_upload_and_analyse_ms": cold_s * 1e3}


def run_cpu_baseline(torch, csr, x, args):
    from spmv_acc_b200 import synth
    try:
        rows = 1 << 22  # bounded sample: first 4.19M rows (20.9M nnz), ~0.1 s per pass, 12 passes
        h = host_c2_rows(rows)
        hx = x.cpu().numpy()
        y = synth.vector_numpy(rows, 3)
        time_reference_spmv(h, hx, y, 2)
        times, kind = time_reference_spmv(h, hx, y, 10)
        sec = float(np.median(times))
        return {"value": 2.0 * h.nnz / sec / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
                "sample": f"first {rows} of {C2_GRID * C2_GRID} rows of the C2 matrix ({h.nnz} nnz), median of 10 passes",
                "effective_gbs": alg_bytes(h.rows, h.cols, h.nnz) / sec / 1e9, "host_cores_available": os.cpu_count()}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--short-max", type=int, default=0)
    ap.add_argument("--medium-max", type=int, default=0)
    ap.add_argument("--vec-div", type=int, default=0)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-context", action="store_true")
    ap.add_argument("--no-iterated", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C3 / C4 lines (N = 1 only)")
    ap.add_argument("--iter-grid", type=int, default=384)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--exchange", default="auto", choices=["auto", "allgather", "halo"])
    ap.add_argument("--no-fused", action="store_true", help="halo exchange with NCCL send/recv instead of the fused push")
    ap.add_argument("--graph", action="store_true", help="replay the power loop from a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="do not overlap the halo exchange with interior rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (one process per GPU)")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
