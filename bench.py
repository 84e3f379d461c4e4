#!/usr/bin/env python
"""Benchmark of the B200-native fp64 CSR SpMV engine on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline workload at every N (config C5 of BASELINE.json, the one its metric "1/2/4/8 B200" is defined on): the 3D
27-point stencil on a 384^3 grid (56.6M rows, 1.52G nnz), fp64, int32 indices, iterated x <- A*x (alpha = 1, beta = 0).
One step = one iteration = one SpMV over the whole matrix plus the exchange of x. With N > 1 ranks (torchrun, one rank
per GPU) the rows are cut into N nnz-balanced contiguous shards (strong scaling); the halo rows of x are pushed into the
neighbours' memory by the SpMV kernels themselves (spmv_b200_halo_loop_*), NCCL exchanges are timed beside it.
The one-shot configs C2 / C3 / C4 are measured by the same run and reported under "other_configs", each with its own
roofline object and an out-of-timed-region check of sampled rows against the reference's CPU SpMV ("verified").

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

C5_GRID = 384
C2_GRID = 4096
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent
NOMINAL_HBM_GBS = 8000.0   # BASELINE.json north_star
METRIC = "fp64 CSR SpMV GFLOP/s (2*nnz/t)"


def workload_name(N):
    n = N ** 3
    nnz = (3 * N - 2) ** 3
    return (f"C5: 3D 27-point stencil {N}^3 ({n} rows, {nnz} nnz), fp64, int32 indices, power loop x <- A*x, "
            f"alpha=1, beta=0")


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


def profiled_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this round (profiles/): ncu
    counters cannot be read inside an unprofiled run, so the number is labelled with its source."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        rec = json.loads(p.read_text())[key]
        note = f"; {rec['note']}" if rec.get("note") else ""
        return rec["dram_bytes_per_launch"], f"profiles/{rec['source']} (ncu --set full, same kernel and matrix{note})"
    except Exception:
        return None, "no ncu capture for this configuration"


def roofline(b_alg, ms, peak, peak_src, kernel, traffic_key=None, rows=0, nnz=0, streamed=None):
    achieved = b_alg / (ms * 1e-3) / 1e9
    traffic, src = profiled_traffic(traffic_key) if traffic_key else (None, "not captured at this N")
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": traffic, "traffic_source": src, "kernel": kernel, "peak_source": peak_src,
         "algorithmic_bytes_per_launch": b_alg, "frac_of_nominal_8TBs": achieved / NOMINAL_HBM_GBS,
         "harness_model_gbs": harness_bytes(rows, nnz) / (ms * 1e-3) / 1e9 if rows else None}
    if streamed is not None and streamed != b_alg:
        # `achieved` counts the CSR arrays as the caller holds them (12 B per non-zero); the staged-x
        # form streams a 16-bit column index from the plan instead of the 32-bit one, so the bytes that actually cross
        # the HBM interface are fewer and `frac` can exceed 1. The two lines below are the same launch in bytes moved.
        r["streamed_bytes_per_launch"] = streamed
        r["streamed_gbs"] = streamed / (ms * 1e-3) / 1e9
        r["streamed_frac_of_peak"] = r["streamed_gbs"] / peak
    if traffic:
        r["dram_gbs_from_traffic"] = traffic / (ms * 1e-3) / 1e9
    return r


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
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
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(index: int):
    """Pins this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers of the
    end-to-end path are allocated next to the GPU's PCIe root (every rank on its own node instead of all on node 0)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return {"numa_node": None, "note": "the platform reports no NUMA affinity for this GPU"}
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus_bound": len(allowed)}
    except Exception as e:
        return {"numa_node": None, "note": f"{type(e).__name__}: {e}"}


# ----------------------------------------------------------------------------------------------------------------
# the reference's own CPU SpMV (cli/verification.cpp:56-66) on the box's host cores
# ----------------------------------------------------------------------------------------------------------------
def host_c5_rows(N, rows):
    """First `rows` rows of the C5 matrix on the host (numpy restatement of the device generator)."""
    from spmv_acc_b200 import synth
    return synth.stencil3d_numpy(N, 0, rows)


def time_reference_spmv(csr, x, y, reps, alpha=1.0, beta=0.0):
    import ctypes as C
    import oracle
    lib = oracle.oracle._ref_lib() if oracle.have_ref() else oracle.oracle._port_lib()
    fn = lib.ref_host_spmv_axpby if oracle.have_ref() else lib.port_host_spmv_axpby
    P = oracle.oracle._p
    args = (C.c_double(alpha), C.c_double(beta), P(csr.val, C.c_double), P(csr.rowptr, C.c_int), P(csr.col, C.c_int),
            C.c_int(csr.rows), C.c_int(csr.cols), C.c_int(csr.nnz), P(x, C.c_double), P(y, C.c_double))
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(*args)
        times.append(time.perf_counter() - t0)
    return times, ("reference" if oracle.have_ref() else "port")


def cpu_sample(N, rows):
    from spmv_acc_b200 import synth
    csr = host_c5_rows(N, rows)
    x = synth.vector_numpy(min(N ** 3, rows + N * N + N + 2), 2)  # the columns these rows reference
    csr.cols = x.size
    return csr, x, np.zeros(csr.rows)


def run_reference(args, rank):
    if rank != 0:
        return
    N = args.grid
    total_steps = args.steps + args.warmup
    budget_s = 60.0
    probe, x, y = cpu_sample(N, min(N ** 3, 1 << 19))  # calibrate, then size the per-step sample to the budget
    t, kind = time_reference_spmv(probe, x, y, 2)
    rows_per_s = probe.rows / min(t)
    rows = int(min(N ** 3, max(1 << 14, rows_per_s * budget_s / total_steps)))
    if rows != probe.rows:
        probe, x, y = cpu_sample(N, rows)
    time_reference_spmv(probe, x, y, args.warmup)
    times, kind = time_reference_spmv(probe, x, y, args.steps)
    sec = float(np.mean(times))
    gflops = 2.0 * probe.nnz / sec / 1e9
    sample = (f"first {probe.rows} of {N ** 3} rows of the C5 matrix per step ({probe.nnz} nnz), serial host_spmv "
              f"(cli/verification.cpp:56-66), alpha=1, beta=0")
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(N), "sample": sample},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(),
                         "note": "host_spmv is serial: one thread is all the reference's CPU path can use"},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "effective_gbs": alg_bytes(probe.rows, probe.cols, probe.nnz) / sec / 1e9,
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


def min_over_ranks(torch, dist, world, value):
    return -max_over_ranks(torch, dist, world, -value)


def dominant_kernel(info, halo=False):
    if info.direct:
        return "k_spmv_warp (direct form, one warp per row block, no shared memory)"
    kind = int(np.argmax(list(info.tiles_per_kind)))
    name = ("SHORT", "MEDIUM", "MIXED")[kind]
    if info.xstage and info.ring_ctas > 0 and kind < 2:
        return (f"k_spmv_ring<{name}{', HALO' if halo else ''}> (staged x, persistent: {info.ring_ctas} CTAs/SM x "
                f"{info.ring_stages} stages, producer warps issue the TMA copies"
                f"{', halo flag protocol inside' if halo else ''})")
    if kind == 2:
        return "k_spmv_mixed<TMA>"
    return (f"k_spmv_rows{'_halo' if halo else ''}<TMA, {name}{', staged x' if info.xstage else ''}>"
            + (" with the halo flag protocol inside" if halo else ""))


def streamed_bytes(info, m, n, nnz):
    """Bytes the plan's form streams per SpMV when every array is read once: the staged-x form reads a 16-bit local
    column index (plan-time re-encoding, 2 B/nnz) instead of colindex (4 B/nnz)."""
    return alg_bytes(m, n, nnz) - (2 * nnz if info.xstage else 0)


def verify_sampled(torch, csr, plan, alpha, beta, seed_x=2, seed_y=3, count=20000, must_include=(), x=None):
    """Out of the timed region: one SpMV on fresh vectors, sampled rows against the reference's CPU SpMV."""
    from oracle import sampled
    from spmv_acc_b200 import synth
    try:
        x = synth.vector_device(csr.cols, seed_x) if x is None else x
        y0 = synth.vector_device(csr.rows, seed_y)
        y = y0.clone()
        plan.execute(alpha, beta, x, y)
        torch.cuda.synchronize()
        rows = sampled.sample_rows(csr.rows, count, must_include=must_include)
        res = sampled.check_sampled_rows(csr.rowptr, csr.col, csr.val, x, y0, y, alpha, beta, rows)
        y2 = y0.clone()
        plan.execute(alpha, beta, x, y2)
        torch.cuda.synchronize()
        res["bitwise_reproducible"] = bool(torch.equal(y.view(torch.int64), y2.view(torch.int64)))
        res["ok"] = bool(res["ok"] and res["bitwise_reproducible"])
        res["alpha_beta"] = [alpha, beta]
        return res
    except Exception as e:
        return {"ok": False, "error": f"{type(e).__name__}: {e}"}


def time_plan(torch, dist, world, plan, alpha, beta, x, y, reps, warm=5, stats=None):
    """Mean ms per SpMV over `reps` back-to-back launches (one event pair around all of them: what the lines report).
    With `stats` (a dict) a second pass times every launch with its own event pair and adds the median and the minimum
    (SURVEY.md §8d asks for both; an event record between launches costs a few microseconds, so the mean stays the
    back-to-back figure)."""
    for _ in range(warm):
        plan.execute(alpha, beta, x, y)
    sync_all(torch, dist, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.execute(alpha, beta, x, y)
    e1.record()
    e1.synchronize()
    mean = e0.elapsed_time(e1) / reps
    if stats is not None:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            plan.execute(alpha, beta, x, y)
            ev[i + 1].record()
        ev[-1].synchronize()
        per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        stats["ms_median_per_launch_events"] = per[len(per) // 2]
        stats["ms_min_per_launch_events"] = per[0]
        stats["reps"] = reps
    sync_all(torch, dist, world)
    return mean


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from spmv_acc_b200 import sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    numa = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peak, peak_src = measured_peak()
    N = args.grid

    # ---- headline: C5 power loop, strong scaling ----
    t0 = time.perf_counter()
    shard = sharded.build_shard("stencil3d", N=N)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    info = shard.plan.info()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    head = sharded.time_power_loop(shard, "fused", iters=args.steps, warmup=args.warmup, sampler=sampler)
    ms_per_step = head["ms_per_iter"]
    csr = shard.csr
    lo, hi = int(shard.bounds[rank]), int(shard.bounds[rank + 1])
    # roofline of the dominant kernel (rank 0's launch): algorithmic bytes of this rank's shard per launch. x is
    # replicated, but a shard only reads the entries its columns reference (own rows + halo, 4096-entry blocks)
    n_ref = shard.n if world == 1 else min(shard.n, int(shard.need.sum()) * (1 << sharded.BLOCK_SHIFT))
    b_alg = alg_bytes(csr.rows, n_ref, csr.nnz)
    kernel = dominant_kernel(info, halo=bool(head.get("single_launch_kernel")))
    line = {
        "metric": METRIC, "value": head["value"], "unit": "GFLOP/s", "n_gpus": world, "steps": head["iters"],
        "warmup": head["warmup_iters"], "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_name(N),
            "sharding": f"{world} nnz-balanced contiguous row shard(s), x replicated where referenced",
            "exchange": head["exchange"],
            "cache": f"inputs larger than L2 ({alg_bytes(csr.rows, n_ref, csr.nnz) / 1e9:.2f} GB streamed per step "
                     f"and GPU vs 126 MB L2); no flush between steps",
            "tile_nnz": info.tile_nnz, "uses_tma": bool(info.uses_tma), "tiles_per_kind": list(info.tiles_per_kind),
            "split_rows": info.nsplit_rows, "launches_per_step": head.get("launches_per_iteration"),
            "iterations_per_graph_launch": head.get("iterations_per_graph_launch"),
        },
        "effective_gbs": head["effective_gbs"],
        "roofline": roofline(b_alg, ms_per_step, peak, peak_src, kernel, "c5" if world == 1 and N == C5_GRID else None,
                             csr.rows, csr.nnz, streamed=streamed_bytes(info, csr.rows, n_ref, csr.nnz)),
        "gpu_launches": head["iters"] * int(head.get("launches_per_iteration") or 1),
        "x_checksum_first_16th": head["x_checksum_first_16th"],
        "build_matrix_and_plan_s": build_s,
        "clocks": sampler.summary() if sampler else None,
        "host_numa": numa,
    }
    # correctness gate of the headline configuration, outside the timed region: sampled rows (across the z-plane
    # boundaries of the shard) of one SpMV against the reference's CPU SpMV, and bitwise reproducibility
    plane = N * N
    edge = [r for z in (0, 1, (hi - lo) // plane // 2, (hi - lo) // plane - 1) for r in range(z * plane - 8, z * plane + 8)]
    ver = verify_sampled(torch, csr, shard.plan, 1.0, 0.0, count=20000, must_include=edge)
    ok_all = min_over_ranks(torch, dist, world, 1.0 if ver.get("ok") else 0.0) > 0.5
    line["verified"] = bool(ok_all)
    line["verification"] = ver

    # ---- the exchange BASELINE.json names (NCCL all-gather of x) and the launch forms, beside the headline; the same
    # number of iterations from the same x_0, so that every mode must end with the same x (x_checksum_first_16th) ----
    iterated = {"fused": head}
    if not args.no_iterated and world > 1:
        for mode in ("fused_allgather_multicast", "fused_allgather", "nccl_allgather", "nccl_halo", "fused_multi_launch"):
            try:
                iterated[mode] = sharded.time_power_loop(shard, mode, iters=args.steps, warmup=args.warmup)
            except Exception as e:
                iterated[mode] = {"error": f"{type(e).__name__}: {e}"}
    elif not args.no_iterated and not args.quick:
        try:
            iterated["fused_multi_launch"] = sharded.time_power_loop(shard, "fused_multi_launch",
                                                                    iters=args.steps, warmup=args.warmup)
        except Exception as e:
            iterated["fused_multi_launch"] = {"error": f"{type(e).__name__}: {e}"}
    line["iterated"] = iterated
    sums = {r.get("x_checksum_first_16th") for r in iterated.values() if "error" not in r}
    line["iterated_checksums_agree"] = len(sums) == 1

    # ---- e2e: the reference-facing host-buffer call on the same workload (H2D x; SpMV; D2H y every step) ----
    if not args.no_e2e:
        line["e2e"] = run_e2e(torch, dist, world, shard, args)

    # ---- CPU baseline: the reference's own host_spmv on the box's host cores (rank 0, N = 1) ----
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = run_cpu_baseline(N)

    shard.destroy()
    del shard, csr
    torch.cuda.empty_cache()

    # ---- an iterated matrix whose halo is not sparse (C3-shaped): the all-gather is what `auto` picks there ----
    if not args.no_iterated and world > 1 and not args.quick:
        try:
            ush = sharded.build_shard("uniform", m=10_000_000, k=32)
            rec = sharded.time_power_loop(ush, "fused", iters=20, warmup=4)
            rec["workload"] = f"{ush.name}, values / 32, x <- A*x"
            try:
                nc = sharded.time_power_loop(ush, "nccl_allgather", iters=20, warmup=4)
                rec["nccl_allgather_ms_per_iter"] = nc["ms_per_iter"]
                rec["checksums_agree"] = nc["x_checksum_first_16th"] == rec["x_checksum_first_16th"]
            except Exception as e:
                rec["nccl_allgather_error"] = f"{type(e).__name__}: {e}"
            line["iterated_uniform"] = rec
            ush.destroy()
            del ush
            torch.cuda.empty_cache()
        except Exception as e:
            line["iterated_uniform"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- the one-shot configurations of BASELINE.json (C2, C3, C4) ----
    if not args.no_other_configs:
        other = run_other_configs(torch, dist, rank, world, args, peak, peak_src)
        if rank == 0:
            line["other_configs"] = other

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
            for _ in range(5):
                X.spmv_b200_ctx_cusparse_spmv(h, 1.0, 1.0, stream)
            torch.cuda.synchronize()
            reps = 20
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
            reps = 20
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


def run_other_configs(torch, dist, rank, world, args, peak, peak_src):
    """C2, C3 and C4 of BASELINE.json, device resident, default plan options, one-shot SpMV with alpha = beta = 1 like
    the reference harness (benchmark/main.cpp:101-102): ms / GFLOP/s / a full roofline object / sampled rows against
    the reference's CPU SpMV, next to cuSPARSE and CUB on the same buffers (N = 1), plus the timings of
    (alpha, beta) = (0.75, -0.5) and (1, 0). With N > 1 ranks every rank generates the same matrix, takes the
    nnz-balanced row shard `spmv_b200_shard_bounds` gives it as a window of the row pointers (no copy) and multiplies
    it against the replicated x: strong scaling of the one-shot SpMV, no exchange."""
    from spmv_acc_b200 import CsrDesc, SpmvPlan, shard_bounds, synth
    out = {}
    makers = {
        "C2 2D 5-point Laplacian 4096x4096 grid": ("c2", lambda: synth.stencil2d_device(C2_GRID)),
        "C3 uniform-random 1e7 x 1e7, 32 nnz/row": ("c3", lambda: synth.uniform_device(10_000_000, 10_000_000, 32, seed=1)),
        "C4 R-MAT 2^24 rows, 2^28 nnz": ("c4", lambda: synth.rmat_device(24, 16, seed=1)),
    }
    for name, (key, make) in makers.items():
        try:
            csr = make()
            lo, hi = 0, csr.rows
            if world > 1:
                b = shard_bounds(csr.rowptr, csr.rows, world)
                lo, hi = int(b[rank]), int(b[rank + 1])
            nnz_local = int(csr.rowptr[hi].item()) - int(csr.rowptr[lo].item())
            window = synth.Csr(hi - lo, csr.cols, csr.rowptr[lo:hi + 1], csr.col, csr.val)
            plan = SpmvPlan(CsrDesc(hi - lo, csr.cols, nnz_local, window.rowptr, csr.col, csr.val))
            info = plan.info()
            x = synth.vector_device(csr.cols, 2)
            y = synth.vector_device(hi - lo, 3 + rank)
            tstats = {}
            ms_local = time_plan(torch, dist, world, plan, 1.0, 1.0, x, y, 30, stats=tstats)
            ms = max_over_ranks(torch, dist, world, ms_local)
            b_local = alg_bytes(hi - lo, csr.cols, nnz_local)  # per rank: its rows, its non-zeros, the whole x
            rec = {"ms": ms, "gflops": 2.0 * csr.nnz / (ms * 1e-3) / 1e9, "n_gpus": world,
                   "scaling": "strong" if world > 1 else "single GPU", "alpha_beta": [1.0, 1.0],
                   "timing_rank0": tstats,
                   "roofline": roofline(b_local, ms_local, peak, peak_src, dominant_kernel(info),
                                        key if world == 1 else None, hi - lo, nnz_local,
                                        streamed=streamed_bytes(info, hi - lo, csr.cols, nnz_local)),
                   "form": "direct (warp per row block, no shared memory)" if info.direct else "tiled (TMA)",
                   "tile_nnz": info.tile_nnz, "tiles_per_kind_rank0": list(info.tiles_per_kind),
                   "split_rows_rank0": info.nsplit_rows, "launches_per_spmv": info.launches_per_execute,
                   "rows_rank0": hi - lo, "nnz_rank0": nnz_local}
            # the other (alpha, beta) pairs of SURVEY.md §8d
            y2 = y.clone()
            rec["ms_alpha_0.75_beta_-0.5"] = max_over_ranks(torch, dist, world,
                                                            time_plan(torch, dist, world, plan, 0.75, -0.5, x, y2, 10, 2))
            y2.copy_(y)
            rec["ms_alpha_1_beta_0"] = max_over_ranks(torch, dist, world,
                                                      time_plan(torch, dist, world, plan, 1.0, 0.0, x, y2, 10, 2))
            del y2
            # correctness gate outside the timed region: the longest rows (split across row blocks), the rows around
            # them and a random sample against the reference's CPU SpMV
            lens = (window.rowptr[1:] - window.rowptr[:-1])
            longest = torch.topk(lens, min(128, lens.numel())).indices.cpu().numpy().tolist()
            ver = verify_sampled(torch, window, plan, 0.75, -0.5, count=20000, must_include=longest, x=x)
            rec["verified"] = bool(min_over_ranks(torch, dist, world, 1.0 if ver.get("ok") else 0.0) > 0.5)
            rec["verification"] = ver
            del lens
            if world == 1 and not args.no_context:
                rec["context"] = run_cusparse(torch, csr, x, y, args)
            out[name] = rec
            plan.destroy()
            del csr, window, x, y
            torch.cuda.empty_cache()
        except Exception as e:
            # (a failure is the same on every rank — same code, same matrix — so the ranks stay in step)
            out[name] = {"error": f"{type(e).__name__}: {e}"}
    return out


def run_e2e(torch, dist, world, shard, args):
    """The same step through the reference-facing host-buffer call: x of the current iterate comes from pinned host
    memory, y goes back to pinned host memory (spmv_b200_hostmat_spmv on the device-resident matrix, the CLI's
    pattern cli/main.cpp:99-118). beta = 0 with SPMV_B200_FLAG_BETA0_SKIP_Y: y0 does not travel."""
    from spmv_acc_b200 import HostMatrix, make_options, synth, FLAG_BETA0_SKIP_Y
    try:
        csr = shard.csr
        t0 = time.perf_counter()
        hm = HostMatrix(csr.rows, csr.cols, csr.rowptr, csr.col, csr.val, make_options(flags=FLAG_BETA0_SKIP_Y))
        cold_s = time.perf_counter() - t0
        hx = torch.empty(shard.n, dtype=torch.float64, pin_memory=True)
        hy = torch.empty(csr.rows, dtype=torch.float64, pin_memory=True)
        x_lo, x_hi = hm.x_range()  # only the referenced part of x is copied (a row shard reads its rows' columns + halo)
        hx[x_lo:x_hi].copy_(synth.vector_device(shard.n, 2)[x_lo:x_hi])
        hy.zero_()
        steps = max(3, min(args.steps, args.e2e_steps))
        for _ in range(3):
            hm.spmv(1.0, 0.0, hx, hy)
        sync_all(torch, dist, world)
        t0 = time.perf_counter()
        for _ in range(steps):
            hm.spmv(1.0, 0.0, hx, hy)  # synchronous: returns after the D2H copy of y completed
        torch.cuda.synchronize()
        sec_local = (time.perf_counter() - t0) / steps
        sec = max_over_ranks(torch, dist, world, sec_local)
        h2d, d2h = int(8 * (x_hi - x_lo)), int(8 * csr.rows)
        hm.destroy()
        del hx, hy
        return {"value": 2.0 * shard.nnz_total / sec / 1e9, "unit": "GFLOP/s", "ms_per_step": sec * 1e3,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
                "pcie_gbs_rank0": {"h2d_plus_d2h_over_step": (h2d + d2h) / sec_local / 1e9},
                "api": "spmv_b200_hostmat_spmv (matrix resident on the device, x copied in from and y copied out to "
                       "pinned host memory every step, row chunks pipelined; cli/main.cpp:99-118 pattern)",
                "create_and_analyse_ms": cold_s * 1e3}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def run_cpu_baseline(N):
    try:
        rows = min(N ** 3, 1 << 21)  # bounded sample: first 2.1M rows (55M nnz), ~0.06 s per pass
        h, x, y = cpu_sample(N, rows)
        time_reference_spmv(h, x, y, 2)
        times, kind = time_reference_spmv(h, x, y, 40)
        sec = float(np.median(times))
        return {"value": 2.0 * h.nnz / sec / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
                "sample": f"first {rows} of {N ** 3} rows of the C5 matrix ({h.nnz} nnz), alpha=1, beta=0, median of 40 "
                          f"passes ({sum(times):.1f} s of CPU work)",
                "effective_gbs": alg_bytes(h.rows, h.cols, h.nnz) / sec / 1e9, "host_cores_available": os.cpu_count(),
                "note": "host_spmv (cli/verification.cpp:56-66) is serial: one thread is all it can use"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="timed iterations of the power loop")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=C5_GRID, help="edge of the 3D stencil grid (384 = config C5)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-context", action="store_true", help="skip cuSPARSE / CUB on the one-shot configs")
    ap.add_argument("--no-iterated", action="store_true", help="skip the NCCL exchanges timed beside the headline")
    ap.add_argument("--no-other-configs", action="store_true", help="skip C2 / C3 / C4")
    ap.add_argument("--quick", action="store_true", help="headline + e2e + cpu baseline only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 2)
    if args.quick:
        args.no_other_configs = True
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
