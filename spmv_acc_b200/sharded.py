"""Multi-GPU row sharding (one process per GPU, torch.distributed for the plumbing).

The one natural sharding of CSR SpMV: contiguous row shards balanced by nnz. Shard g owns rows
[bounds[g], bounds[g+1]) with bounds[g] = lower_bound(rowptr, g*nnz/G) (``spmv_b200_shard_bounds``, bit-exact
against oracle/analysis_port.c:port_shard_bounds). x is replicated; a one-shot SpMV needs no communication.

In an iterated loop (x <- A*x) every rank writes its y shard straight into its slice of the next x and the slices are
exchanged:
  * ``allgather`` — every slice goes to every rank (NCCL broadcasts of the unequal slices = all-gather-v). This is the
    exchange BASELINE.json names; it moves 8*n*(G-1)/G bytes into every GPU per iteration.
  * ``halo``      — only the 4096-entry blocks of x that a rank's columns actually reference are sent to it
    (grouped NCCL send/recv). The needed blocks come from the analysis (``spmv_b200_col_block_bitmap``) and are
    exchanged once at set-up. For a z-slab sharded 27-point stencil that is one xy-plane per neighbour instead of the
    whole vector; for a uniform-random matrix every block is needed and the schedule degenerates to the all-gather.
The reference has no multi-GPU path: the contract is that each shard's y equals the single-GPU y
for those rows, bit for bit (the kernels are deterministic and a shard's tiles do not depend on the other shards'
values, only on its own rowptr).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np

BLOCK_SHIFT = 12  # exchange granularity: 4096 entries of x = 32 KB


def _dist():
    import torch.distributed as dist
    return dist


def world_info() -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def merge_runs(blocks: np.ndarray, lo: int, hi: int, shift: int = BLOCK_SHIFT) -> List[Tuple[int, int]]:
    """Element ranges [a, b) inside [lo, hi) covered by the given ascending block ids, adjacent blocks merged."""
    runs: List[Tuple[int, int]] = []
    for b in blocks.tolist():
        a, e = max(b << shift, lo), min((b + 1) << shift, hi)
        if a >= e:
            continue
        if runs and runs[-1][1] == a:
            runs[-1] = (runs[-1][0], e)
        else:
            runs.append((a, e))
    return runs


def exchange_schedule(need: np.ndarray, bounds: np.ndarray, rank: int, shift: int = BLOCK_SHIFT):
    """need[q, b] = 1 if rank q references block b of x. Returns (sends, recvs): lists of (peer, start, end) element
    ranges of x, ordered by (peer, start) on both sides so that matching send/recv pairs line up."""
    G = need.shape[0]
    sends, recvs = [], []
    for p in range(G):
        if p == rank:
            continue
        # what p needs from my rows
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        if hi > lo:
            blk = np.arange(lo >> shift, ((hi - 1) >> shift) + 1)
            for a, e in merge_runs(blk[need[p, blk] != 0], lo, hi, shift):
                sends.append((p, a, e))
        # what I need from p's rows
        lo, hi = int(bounds[p]), int(bounds[p + 1])
        if hi > lo:
            blk = np.arange(lo >> shift, ((hi - 1) >> shift) + 1)
            for a, e in merge_runs(blk[need[rank, blk] != 0], lo, hi, shift):
                recvs.append((p, a, e))
    return sends, recvs


@dataclass
class PowerLoop:
    """x <- A*x on row shards. ``spmv(x_full, y_slice)`` must compute y_slice = A_shard * x_full (alpha=1, beta=0)."""
    n: int
    bounds: np.ndarray
    spmv: Callable
    x: "object"            # current x (full length, replicated where referenced)
    x_next: "object"       # next x
    need_local: np.ndarray  # uint8 [nblocks]: blocks of x this rank references
    exchange: str = "auto"
    block_shift: int = BLOCK_SHIFT
    sends: list = field(default_factory=list)
    recvs: list = field(default_factory=list)
    mode: str = ""
    bytes_in_per_iter: int = 0

    def __post_init__(self):
        import torch
        dist = _dist()
        self.rank, self.world = world_info()
        if self.world == 1:
            self.mode = "none"
            return
        mine = torch.from_numpy(np.ascontiguousarray(self.need_local)).to(self.x.device)
        allneed = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=self.x.device)
        dist.all_gather_into_tensor(allneed, mine) if self.x.is_cuda else dist.all_gather(
            list(allneed.view(self.world, -1).unbind(0)), mine)
        need = allneed.view(self.world, -1).cpu().numpy()
        self.sends, self.recvs = exchange_schedule(need, self.bounds, self.rank, self.block_shift)
        halo_in = sum(e - a for _, a, e in self.recvs) * 8
        full_in = (self.n - int(self.bounds[self.rank + 1] - self.bounds[self.rank])) * 8
        if self.exchange == "auto":
            # the sparse exchange pays off when it moves well under the full vector
            self.mode = "halo" if halo_in <= 0.5 * full_in else "allgather"
        else:
            self.mode = self.exchange
        self.bytes_in_per_iter = halo_in if self.mode == "halo" else full_in

    def _exchange(self, v):
        dist = _dist()
        if self.mode == "allgather":
            for src in range(self.world):
                lo, hi = int(self.bounds[src]), int(self.bounds[src + 1])
                if hi > lo:
                    dist.broadcast(v[lo:hi], src)
        elif self.mode == "halo":
            ops = [dist.P2POp(dist.isend, v[a:e], p) for p, a, e in self.sends]
            ops += [dist.P2POp(dist.irecv, v[a:e], p) for p, a, e in self.recvs]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()

    def step(self):
        lo, hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.spmv(self.x, self.x_next[lo:hi])
        if self.world > 1:
            self._exchange(self.x_next)
        self.x, self.x_next = self.x_next, self.x

    def run(self, iters: int):
        for _ in range(iters):
            self.step()
        return self.x


def bits_checksum(t) -> int:
    """Order-independent checksum of the exact bit patterns (wrapping int64 sum) — equal iff multisets of bits match."""
    import torch
    return int(t.view(torch.int64).sum().item())


def build_stencil3d_power_loop(N: int, exchange: str = "auto", options=None):
    """Rank-local pieces of the C5 configuration: 27-point averaging stencil on an N^3 grid, rows sharded by nnz."""
    import torch
    from . import CsrDesc, SpmvPlan, col_block_bitmap, make_options, shard_bounds, synth, FLAG_BETA0_SKIP_Y
    rank, world = world_info()
    n = N ** 3
    if world == 1:
        bounds = np.array([0, n], dtype=np.int64)
    else:
        counts = synth.stencil_row_counts_device("stencil3d", N)
        rowptr = synth._rowptr_from_counts_device(counts)
        del counts
        bounds = shard_bounds(rowptr, n, world).astype(np.int64)
        del rowptr
        torch.cuda.empty_cache()
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    csr = synth.stencil3d_device(N, lo, hi)
    opt = options if options is not None else make_options(flags=FLAG_BETA0_SKIP_Y)
    plan = SpmvPlan(CsrDesc(csr.rows, csr.cols, csr.nnz, csr.rowptr, csr.col, csr.val), opt)
    need = col_block_bitmap(csr.col, csr.nnz, n, BLOCK_SHIFT) if world > 1 else np.zeros(0, np.uint8)
    x = synth.vector_device(n, 2)
    x_next = torch.zeros_like(x)

    def spmv(xf, ys):
        plan.execute(1.0, 0.0, xf, ys)

    loop = PowerLoop(n=n, bounds=bounds, spmv=spmv, x=x, x_next=x_next, need_local=need, exchange=exchange)
    return loop, plan, csr


def bench_power_loop(N: int = 384, iters: int = 100, exchange: str = "auto", warmup: int = 3) -> dict:
    """Times `iters` iterations of x <- A*x for the 27-point N^3 stencil on all ranks (device events, max over ranks)."""
    import torch
    dist = _dist()
    rank, world = world_info()
    loop, plan, csr = build_stencil3d_power_loop(N, exchange)
    nnz_local = csr.nnz

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loop.run(warmup)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(iters)
    e1.record()
    e1.synchronize()
    sync()
    ms = e0.elapsed_time(e1)
    stats = torch.tensor([ms, float(nnz_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        ms, nnz_total = float(mx[0]), float(stats[1])
    else:
        nnz_total = float(nnz_local)
    n = N ** 3
    sec_iter = ms * 1e-3 / iters
    lo, hi = int(loop.bounds[rank]), int(loop.bounds[rank + 1])
    out = {
        "workload": f"C5: 27-point stencil {N}^3 ({n} rows, {int(nnz_total)} nnz), x <- A*x, {iters} iterations, "
                    f"alpha=1, beta=0",
        "scaling": "strong", "n_gpus": world, "iters": iters, "ms_per_iter": sec_iter * 1e3,
        "value": 2.0 * nnz_total / sec_iter / 1e9, "unit": "GFLOP/s",
        "effective_gbs": (12 * nnz_total + 4 * (n + 1) + 8 * n + 16 * n) / sec_iter / 1e9,
        "exchange": loop.mode, "exchange_bytes_in_per_iter_rank0": loop.bytes_in_per_iter,
        # bit pattern checksum of x[0 : n/16] after the last iteration: that range belongs to rank 0 for every
        # world size <= 8, so equal values across runs with 1/2/4/8 GPUs prove bitwise-identical results
        "x_checksum_first_16th": bits_checksum(loop.x[0:n // 16]),
        "total_timed_ms": ms,
    }
    plan.destroy()
    return out
