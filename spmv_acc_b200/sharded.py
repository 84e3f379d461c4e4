"""Multi-GPU row sharding (one process per GPU, torch.distributed for the plumbing).

The one natural sharding of CSR SpMV: contiguous row shards balanced by nnz. Shard g owns rows
[bounds[g], bounds[g+1]) with bounds[g] = lower_bound(rowptr, g*nnz/G) (``spmv_b200_shard_bounds``, bit-exact
against oracle/analysis_port.c:port_shard_bounds). x is replicated; a one-shot SpMV needs no communication.

In an iterated loop (x <- A*x) every rank writes its y shard straight into its slice of the next x and the slices are
exchanged:
  * ``allgather`` — every slice goes to every rank: ONE grouped NCCL collective per iteration (ncclAllGather in place
    when the slices have equal length, otherwise one ncclGroup of broadcasts of the unequal slices = all-gather-v) on a
    communication stream, overlapped with the row blocks that only read this rank's own columns. This is the exchange
    BASELINE.json names; it moves 8*n*(G-1)/G bytes into every GPU per iteration.
  * ``halo``      — only the 4096-entry blocks of x that a rank's columns actually reference are sent to it
    (grouped NCCL send/recv). The needed blocks come from the analysis (``spmv_b200_col_block_bitmap``) and are
    exchanged once at set-up. For a z-slab sharded 27-point stencil that is one xy-plane per neighbour instead of the
    whole vector; for a uniform-random matrix every block is needed and the schedule degenerates to the all-gather.
The reference has no multi-GPU path: the contract is that each shard's y equals the single-GPU y
for those rows, bit for bit (the kernels are deterministic and a shard's tiles do not depend on the other shards'
values, only on its own rowptr).
"""
from __future__ import annotations

import os
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


def _ranges(marks: np.ndarray) -> List[Tuple[int, int]]:
    """Maximal runs [a, b) of True in a boolean array."""
    edges = np.flatnonzero(np.diff(np.concatenate(([False], np.asarray(marks, dtype=bool), [False])).astype(np.int8)))
    return [(int(edges[i]), int(edges[i + 1])) for i in range(0, edges.size, 2)]


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


def align_bounds(bounds: np.ndarray, rows: int = 16) -> np.ndarray:
    """Interior shard boundaries rounded to a multiple of `rows` (16 rows = one 128-byte line of x): a slice of x that
    starts at an odd element is not 16-byte aligned, and NCCL then moves it 2.5x slower (8 GPUs, 453 MB: 1.78 ms
    against 0.70 ms, profiles/r2_allgather_loop_probe_n8.jsonl); the nnz balance moves by at most 8 rows."""
    b = np.asarray(bounds, dtype=np.int64).copy()
    if b.size > 2:
        b[1:-1] = np.clip((b[1:-1] + rows // 2) // rows * rows, b[0], b[-1])
        b = np.maximum.accumulate(b)
    return b


def allgather_schedule(bounds: np.ndarray, rank: int):
    """(sends, recvs) of the exchange in which every rank's whole slice goes to every other rank."""
    G = len(bounds) - 1
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    sends = [(p, lo, hi) for p in range(G) if p != rank and hi > lo]
    recvs = [(p, int(bounds[p]), int(bounds[p + 1])) for p in range(G) if p != rank and bounds[p + 1] > bounds[p]]
    return sends, recvs


def split_boundary_interior(sends, tile_row, tile_reads_halo, lo: int):
    """(boundary, interior) tile ranges of a shard for the halo exchange, or None if most of the shard is boundary.
    Boundary = row blocks that produce a row another rank receives, plus (when known) every row block that reads an
    entry of x owned by another rank: once the boundary blocks of an iteration are done, this rank neither owes anybody
    a row of that iteration nor reads a halo entry of it, which is what lets a neighbour overwrite the halo early
    (fused push) while the interior is still being multiplied."""
    tr = np.asarray(tile_row, dtype=np.int64)
    nt = tr.size - 1
    if nt <= 0:
        return None
    marks = np.zeros(nt, dtype=bool)
    for _, a, e in sends:
        t0 = int(np.searchsorted(tr, a - lo, side="right")) - 1
        t1 = int(np.searchsorted(tr, e - lo, side="left"))
        marks[max(t0, 0):min(t1, nt)] = True
    if tile_reads_halo is not None:
        marks |= np.asarray(tile_reads_halo, dtype=bool)[:nt]
    if marks.sum() * 2 > nt:
        return None  # nothing to hide the exchange behind
    return _ranges(marks), _ranges(~marks)


def split_by_remote_reads(tile_row, tile_reads_halo):
    """(readers of remote columns, the rest) as tile ranges, or None if most row blocks read remote columns."""
    if tile_reads_halo is None:
        return None
    nt = np.asarray(tile_row).size - 1
    marks = np.asarray(tile_reads_halo, dtype=bool)[:nt]
    if nt <= 0 or marks.sum() * 2 > nt:
        return None
    return _ranges(marks), _ranges(~marks)


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
    # optional overlap of the halo exchange with the rows nobody waits for:
    spmv_tiles: Optional[Callable] = None   # spmv_tiles(x_full, y_slice, tile_lo, tile_hi): row blocks of the shard
    tile_row: Optional[np.ndarray] = None   # [ntiles+1] first local row of each row block
    tile_reads_halo: Optional[np.ndarray] = None  # bool [ntiles]: the row block references x entries of other ranks
    overlap: bool = True
    boundary: list = field(default_factory=list)   # tile ranges whose rows are sent to other ranks
    interior: list = field(default_factory=list)   # the remaining tile ranges
    sends: list = field(default_factory=list)
    recvs: list = field(default_factory=list)
    mode: str = ""
    bytes_in_per_iter: int = 0
    _ag_native: Optional[bool] = None  # None: not tried yet; False: the backend refused, use broadcasts

    def __post_init__(self):
        import torch
        dist = _dist()
        self.rank, self.world = world_info()
        if self.world == 1:
            self.mode = "none"
            return
        # set-up only: the bitmaps travel as host objects, so any backend (nccl, gloo) works
        gathered = [None] * self.world
        dist.all_gather_object(gathered, np.ascontiguousarray(self.need_local, dtype=np.uint8))
        need = np.stack(gathered)
        self.sends, self.recvs = exchange_schedule(need, self.bounds, self.rank, self.block_shift)
        halo_in = sum(e - a for _, a, e in self.recvs) * 8
        full_in = (self.n - int(self.bounds[self.rank + 1] - self.bounds[self.rank])) * 8
        if self.exchange == "auto":
            # the sparse exchange pays off when it moves well under the full vector
            self.mode = "halo" if halo_in <= 0.5 * full_in else "allgather"
        else:
            self.mode = self.exchange
        self.bytes_in_per_iter = halo_in if self.mode == "halo" else full_in
        self._plan_overlap()

    def _plan_overlap(self):
        """halo: row blocks that produce rows another rank needs are computed first; their exchange then runs on a
        second stream while the rest of the shard is multiplied. allgather: the collective of x runs while the row
        blocks that read only this rank's own columns are multiplied; the others follow."""
        self.overlapped = False
        if not (self.overlap and self.spmv_tiles is not None and self.tile_row is not None):
            return
        if self.mode == "allgather":
            split = split_by_remote_reads(self.tile_row, self.tile_reads_halo)
        elif self.mode == "halo":
            self.boundary_reads_all_halo = self.tile_reads_halo is not None
            split = split_boundary_interior(self.sends, self.tile_row, self.tile_reads_halo,
                                            int(self.bounds[self.rank]))
        else:
            split = None
        if split is None:
            return
        self.boundary, self.interior = split
        self.overlapped = True
        if self.x.is_cuda:
            import torch
            # high priority: when the collective's CTAs and the SpMV's are both waiting for an SM, the collective's go
            # first (its duration, not the SpMV's, bounds the iteration)
            self.comm_stream = torch.cuda.Stream(priority=-1)

    def _allgather(self, v):
        """Every rank's slice of v to every rank, one collective."""
        dist = _dist()
        sizes = np.diff(self.bounds)
        lo, hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        if self._ag_native is not False and v.is_cuda:
            try:
                if np.all(sizes == sizes[0]):
                    dist.all_gather_into_tensor(v, v[lo:hi])  # in place: ncclAllGather, no staging copy
                else:
                    # Unequal slices (nnz-balanced shards), in place: one NCCL group of point-to-point sends / receives.
                    # Measured at 8 GPUs on 453 MB of x (profiles/r2_allgather_probe_n8.jsonl): 0.754 ms against 0.868 ms
                    # for dist.all_gather with a list (a group of broadcasts), 0.910 ms for ncclAllGather of padded
                    # slots + copies into place, and 0.698 ms for ncclAllGather of equal slices.
                    ops = []
                    for d in range(1, self.world):
                        dst, src = (self.rank + d) % self.world, (self.rank - d) % self.world
                        a, e = int(self.bounds[src]), int(self.bounds[src + 1])
                        if hi > lo:
                            ops.append(dist.P2POp(dist.isend, v[lo:hi], dst))
                        if e > a:
                            ops.append(dist.P2POp(dist.irecv, v[a:e], src))
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()
                self._ag_native = True
                return
            except Exception:
                if self._ag_native is True:
                    raise
                self._ag_native = False  # backend without grouped point-to-point: broadcasts below
        for src in range(self.world):
            a, e = int(self.bounds[src]), int(self.bounds[src + 1])
            if e > a:
                dist.broadcast(v[a:e], src)

    def _exchange(self, v):
        dist = _dist()
        if self.mode == "allgather":
            self._allgather(v)
        elif self.mode == "halo":
            ops = [dist.P2POp(dist.isend, v[a:e], p) for p, a, e in self.sends]
            ops += [dist.P2POp(dist.irecv, v[a:e], p) for p, a, e in self.recvs]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()

    def step(self):
        lo, hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        ys = self.x_next[lo:hi]
        if self.world > 1 and self.mode == "allgather":
            # Invariant between steps: this rank's own slice of x is final, the other slices are one step old. The
            # collective that refreshes them runs next to the row blocks that do not read them.
            if getattr(self, "overlapped", False) and self.x.is_cuda:
                import torch
                cur = torch.cuda.current_stream()
                self.comm_stream.wait_event(cur.record_event())
                with torch.cuda.stream(self.comm_stream):
                    self._allgather(self.x)
                    done = self.comm_stream.record_event()
                for t0, t1 in self.interior:
                    self.spmv_tiles(self.x, ys, t0, t1)
                cur.wait_event(done)
                for t0, t1 in self.boundary:
                    self.spmv_tiles(self.x, ys, t0, t1)
            else:
                self._allgather(self.x)
                self.spmv(self.x, ys)
        elif self.world > 1 and getattr(self, "overlapped", False):
            for t0, t1 in self.boundary:
                self.spmv_tiles(self.x, ys, t0, t1)
            if self.x.is_cuda:
                import torch
                cur = torch.cuda.current_stream()
                self.comm_stream.wait_event(cur.record_event())
                with torch.cuda.stream(self.comm_stream):
                    self._exchange(self.x_next)
                    done = self.comm_stream.record_event()
                for t0, t1 in self.interior:
                    self.spmv_tiles(self.x, ys, t0, t1)
                cur.wait_event(done)
            else:
                self._exchange(self.x_next)
                for t0, t1 in self.interior:
                    self.spmv_tiles(self.x, ys, t0, t1)
        else:
            self.spmv(self.x, ys)
            if self.world > 1:
                self._exchange(self.x_next)
        self.x, self.x_next = self.x_next, self.x

    def finish(self):
        """Makes every slice of x current (allgather mode leaves the other ranks' slices one step behind)."""
        if self.world > 1 and self.mode == "allgather":
            self._allgather(self.x)
        return self.x

    def run(self, iters: int, finish: bool = True):
        for _ in range(iters):
            self.step()
        return self.finish() if finish else self.x

    def capture(self):
        """Captures two iterations (one ping-pong period of the x buffers) into a CUDA graph: the kernels, the grouped
        NCCL send/recv and the stream fork/join are then replayed without any host work in the loop."""
        import torch
        assert self.x.is_cuda
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside the capture (NCCL connections, lazy module loading)
            self.step()
            self.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.step()
            self.step()
        return self.graph

    def run_graph(self, iters: int):
        """Replays the captured pair of iterations iters/2 times (iters must be even)."""
        assert iters % 2 == 0 and getattr(self, "graph", None) is not None
        for _ in range(iters // 2):
            self.graph.replay()
        return self.x


def make_halo_desc(plan, buf_ptrs, lo: int, hi: int, wait_flags: list, signal_flags: list, push: list,
                   boundary: list, interior: list, flags: int = 0):
    """Fills a ``spmv_b200_halo_loop_desc``. push[b] = [(row_lo, row_hi, dst_address)] for destination buffer parity b
    (dst already offset so that it is indexed by the shard-local row)."""
    from . import _lib
    if len(wait_flags) > _lib.MAX_PUSH or max(len(push[0]), len(push[1])) > _lib.MAX_PUSH:
        raise RuntimeError("too many neighbours / push ranges for the fused halo exchange")
    if max(len(boundary), len(interior)) > _lib.MAX_RANGES:
        raise RuntimeError("too many row-block ranges for the fused halo exchange")
    d = _lib.HaloLoopDesc()
    d.plan = plan._h
    d.buf[0], d.buf[1] = int(buf_ptrs[0]), int(buf_ptrs[1])
    d.row_lo, d.row_hi = int(lo), int(hi)
    d.n_neigh = len(wait_flags)
    for j, (w, g) in enumerate(zip(wait_flags, signal_flags)):
        d.wait_flags[j], d.signal_flags[j] = int(w), int(g)
    for b in (0, 1):
        d.push[b].count = len(push[b])
        d.push[b].multicast_mask = 0
        for j, entry in enumerate(push[b]):  # (row_lo, row_hi, dst_address[, is_multicast_address])
            rl, rh, dst = entry[:3]
            d.push[b].row_lo[j], d.push[b].row_hi[j], d.push[b].dst[j] = int(rl), int(rh), int(dst)
            if len(entry) > 3 and entry[3]:
                d.push[b].multicast_mask |= 1 << j
    d.n_boundary, d.n_interior = len(boundary), len(interior)
    for j, (t0, t1) in enumerate(boundary):
        d.boundary[2 * j], d.boundary[2 * j + 1] = int(t0), int(t1)
    for j, (t0, t1) in enumerate(interior):
        d.interior[2 * j], d.interior[2 * j + 1] = int(t0), int(t1)
    d.flags = int(flags)
    return d


class SymmBuffer:
    """Exchange memory from torch's symmetric memory (cuMemCreate + cuMulticast*, set up by the rendezvous): the same
    allocation on every rank of the group, each rank's copy mapped into every other rank (``peer_address``) and, on
    NVSwitch systems, one multicast address whose stores the switch replicates into all copies."""

    def __init__(self, nbytes: int):
        import torch
        import torch.distributed._symmetric_memory as symm
        dist = _dist()
        self.nbytes = int(nbytes)
        self._t = symm.empty(self.nbytes, dtype=torch.uint8, device="cuda")
        self._h = symm.rendezvous(self._t, dist.group.WORLD)
        self.address = int(self._t.data_ptr())
        self.peer_address = [int(p) for p in self._h.buffer_ptrs]
        self.multicast_address = int(self._h.multicast_ptr)

    def tensor(self, dtype: str, count: int, offset: int):
        import torch
        dt = getattr(torch, dtype)
        size = torch.empty(0, dtype=dt).element_size()
        return self._t[offset:offset + count * size].view(dt)

    def release(self):
        self._h = None
        self._t = None


class FusedHaloLoop:
    """x <- A*x with the halo exchange fused into the SpMV kernels (``spmv_b200_halo_loop_*``): the epilogue of the
    kernels stores the rows a neighbour references straight into that neighbour's copy of the next x (peer memory
    mapped through CUDA IPC, NVLink stores), and iterations of neighbouring GPUs are ordered with flags in each other's
    memory that the boundary row blocks of the SpMV kernel itself wait for and raise: one kernel launch per iteration,
    replayed from a CUDA graph, no NCCL kernel, no host synchronisation. Double buffering: iteration k reads
    buf[k % 2] and writes buf[(k+1) % 2] here and in the neighbours; the boundary row blocks of iteration k start only
    after every neighbour has raised its flag to k, which means the neighbour's rows for x_k have arrived and the
    neighbour no longer reads the buffer about to be overwritten."""

    def __init__(self, base: PowerLoop, plan, flags: int = 0, multicast: bool = False):
        import torch
        from . import HaloLoop, PeerBuffer
        dist = _dist()
        assert base.world > 1 and base.x.is_cuda
        self.base, self.plan = base, plan
        self.rank, self.world = base.rank, base.world
        self.lo, self.hi = int(base.bounds[self.rank]), int(base.bounds[self.rank + 1])
        self.neigh = sorted({p for p, _, _ in base.sends} | {p for p, _, _ in base.recvs})
        n = base.n
        self.xbytes = (8 * n + 255) // 256 * 256
        self.multicast = bool(multicast)
        if self.multicast:
            # all-gather through the switch: the exchange memory comes from torch's symmetric memory, and ONE store to
            # its multicast address puts a row into every GPU's copy of the next x (7 peer stores become one)
            self.own = SymmBuffer(2 * self.xbytes + 4 * self.world)
            if not self.own.multicast_address:
                raise RuntimeError("this system gives no NVLink multicast address (no NVSwitch multicast support)")
            peer_address = {p: self.own.peer_address[p] for p in self.neigh}
            self.peer = {}
        else:
            # one exported allocation per rank: [x buffer 0 | x buffer 1 | flags], opened by the neighbours with THEIR
            # device current (that is what maps it for their kernels; a torch-IPC tensor is mapped for the owner's device)
            self.own = PeerBuffer.alloc(2 * self.xbytes + 4 * self.world)
        self.bufs = [self.own.tensor("float64", n, 0), self.own.tensor("float64", n, self.xbytes)]
        self.flags = self.own.tensor("int32", self.world, 2 * self.xbytes)
        self.flags.zero_()
        self.bufs[0].copy_(base.x)
        torch.cuda.synchronize()
        if self.multicast:
            dist.barrier()
            # every row of the shard, both buffer parities: destination = the multicast image of the buffer
            push = [[(0, self.hi - self.lo, self.own.multicast_address + b * self.xbytes + 8 * self.lo, True)]
                    for b in (0, 1)]
        else:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, (self.own.handle, self.own.nbytes))
            self.peer = {p: PeerBuffer.open(*everyone[p]) for p in self.neigh}
            peer_address = {p: self.peer[p].address for p in self.neigh}
            # push descriptors for both buffer parities: destination = neighbour's buffer, indexed by my local row
            push = [[(a - self.lo, e - self.lo, peer_address[p] + b * self.xbytes + 8 * self.lo)
                     for p, a, e in base.sends] for b in (0, 1)]
        # Full aligned lines per peer store (SPMV_B200_HALO_ALIGN_PUSH) where the link bounds the iteration: what
        # arrives in a GPU per iteration at the ~400 GB/s small peer stores are absorbed with, against the SpMV's own
        # time at the HBM rate. All-gather pushes on 4 or more GPUs qualify; halo pushes and 2 GPUs do not.
        from . import _lib as _l
        info = plan.info()
        bytes_in = 8 * sum(e - a for _, a, e in base.recvs)
        t_push, t_spmv = bytes_in / 4.0e11, (10.0 * info.nnz + 20.0 * info.m) / 6.5e12
        self.align_push = t_push > t_spmv
        if self.align_push:
            flags |= _l.HALO_ALIGN_PUSH
        # boundary-first schedule only if the boundary blocks are known to contain every reader of halo entries
        self.split = bool(getattr(base, "overlapped", False) and getattr(base, "boundary_reads_all_halo", False))
        desc = make_halo_desc(
            plan, [self.bufs[0].data_ptr(), self.bufs[1].data_ptr()], self.lo, self.hi,
            [self.flags.data_ptr() + 4 * p for p in self.neigh],
            [peer_address[p] + 2 * self.xbytes + 4 * self.rank for p in self.neigh], push,
            base.boundary if self.split else [], base.interior if self.split else [], flags)
        self.loop = HaloLoop(desc)
        self.k = 0
        dist.barrier()

    def close(self):
        """Waits for the enqueued iterations (raises if a neighbour's flag timed out) and releases the mappings."""
        err = None
        try:
            self.loop.sync()
        except Exception as e:  # keep going: the peers must still be released collectively
            err = e
        _dist().barrier()
        self.loop.destroy()
        for pb in self.peer.values():
            pb.release()
        self.peer = {}
        _dist().barrier()
        self.bufs, self.flags = [], None
        self.own.release()
        if err is not None:
            raise err

    def run(self, iters: int):
        self.loop.run(int(iters))
        self.k += int(iters)
        return self.x

    def sync(self):
        self.loop.sync()

    @property
    def x(self):
        return self.bufs[self.k % 2]


def bits_checksum(t) -> int:
    """Order-independent checksum of the exact bit patterns (wrapping int64 sum) — equal iff multisets of bits match."""
    import torch
    return int(t.view(torch.int64).sum().item())


@dataclass
class Shard:
    """This rank's rows of an iterated configuration: matrix, plan and what the exchange schedules need."""
    name: str
    n: int
    bounds: np.ndarray
    csr: "object"
    plan: "object"
    need: np.ndarray
    tile_row: Optional[np.ndarray]
    reads_halo: Optional[np.ndarray]
    nnz_total: int = 0

    def destroy(self):
        self.plan.destroy()


def build_shard(kind: str = "stencil3d", N: int = 384, m: int = 0, k: int = 32, options=None) -> Shard:
    """kind = "stencil3d": 27-point averaging stencil on an N^3 grid (config C5); kind = "uniform": m x m matrix with k
    uniformly random columns per row, values scaled by 1/k (C3-shaped; every rank needs all of x). Rows are sharded
    by nnz (``spmv_b200_shard_bounds``); each rank generates only its own rows."""
    import torch
    from . import CsrDesc, SpmvPlan, col_block_bitmap, make_options, shard_bounds, synth, FLAG_BETA0_SKIP_Y
    rank, world = world_info()
    if kind == "stencil3d":
        n = N ** 3
        if world == 1:
            bounds = np.array([0, n], dtype=np.int64)
        else:
            counts = synth.stencil_row_counts_device("stencil3d", N)
            rowptr = synth._rowptr_from_counts_device(counts)
            del counts
            bounds = align_bounds(shard_bounds(rowptr, n, world).astype(np.int64))
            del rowptr
            torch.cuda.empty_cache()
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        csr = synth.stencil3d_device(N, lo, hi)
        name = f"27-point stencil {N}^3"
    elif kind == "uniform":
        n = int(m)
        bounds = np.array([(n * g) // world for g in range(world + 1)], dtype=np.int64)  # k per row: rows = nnz balance
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        csr = synth.uniform_device(n, n, k, seed=1, r_lo=lo, r_hi=hi)
        csr.val.mul_(1.0 / k)  # keeps 100 iterations of x <- A*x inside the fp64 range
        name = f"uniform-random {n} x {n}, {k} nnz/row"
    else:
        raise ValueError(kind)
    opt = options if options is not None else make_options(flags=FLAG_BETA0_SKIP_Y)
    plan = SpmvPlan(CsrDesc(csr.rows, csr.cols, csr.nnz, csr.rowptr, csr.col, csr.val), opt)
    need = col_block_bitmap(csr.col, csr.nnz, n, BLOCK_SHIFT) if world > 1 else np.zeros(0, np.uint8)
    info = plan.info()
    tile_row = reads_halo = None
    if info.nsplit_rows == 0 and world > 1:
        cmin, cmax = plan.tile_col_range()
        reads_halo = (cmin < lo) | (cmax >= hi)
        tile_row = plan.export("tile_row")
    nnz_total = csr.nnz
    if world > 1:
        t = torch.tensor([float(csr.nnz)], dtype=torch.float64, device="cuda")
        _dist().all_reduce(t)
        nnz_total = int(t.item())
    return Shard(name, n, bounds, csr, plan, need, tile_row, reads_halo, nnz_total)


def make_loop(shard: Shard, exchange: str = "auto", overlap: bool = True) -> PowerLoop:
    """A fresh loop (x = the seeded start vector) over an existing shard."""
    import torch
    from . import synth
    plan = shard.plan

    def spmv(xf, ys):
        plan.execute(1.0, 0.0, xf, ys)

    def spmv_tiles(xf, ys, t0, t1):
        plan.execute_tiles(1.0, 0.0, xf, ys, t0, t1)

    x = synth.vector_device(shard.n, 2)
    loop = PowerLoop(n=shard.n, bounds=shard.bounds, spmv=spmv, x=x, x_next=torch.zeros_like(x),
                     need_local=shard.need, exchange=exchange,
                     spmv_tiles=spmv_tiles if shard.tile_row is not None else None, tile_row=shard.tile_row,
                     tile_reads_halo=shard.reads_halo, overlap=overlap)
    loop.plan = plan
    return loop


def build_stencil3d_power_loop(N: int, exchange: str = "auto", options=None, overlap: bool = True):
    """(loop, plan, csr) of the C5 configuration for this rank."""
    shard = build_shard("stencil3d", N=N, options=options)
    return make_loop(shard, exchange, overlap), shard.plan, shard.csr


class LocalLoop:
    """The natively enqueued loop of a single rank (no neighbours): same runner as FusedHaloLoop, nothing to order."""

    def __init__(self, base: PowerLoop, plan, flags: int = 0):
        from . import HaloLoop
        self.bufs = [base.x, base.x_next]
        desc = make_halo_desc(plan, [self.bufs[0].data_ptr(), self.bufs[1].data_ptr()], 0, base.n, [], [], [[], []],
                              [], [], flags)
        self.loop = HaloLoop(desc)
        self.k = 0
        self.split = False

    def run(self, iters: int):
        self.loop.run(int(iters))
        self.k += int(iters)
        return self.x

    def close(self):
        self.loop.sync()
        self.loop.destroy()

    @property
    def x(self):
        return self.bufs[self.k % 2]


MODES = ("fused", "fused_multi_launch", "nccl_halo", "nccl_allgather", "fused_allgather", "fused_allgather_multicast")


def time_power_loop(shard: Shard, mode: str = "fused", iters: int = 100, warmup: int = 4, sampler=None) -> dict:
    """Times `iters` iterations of x <- A*x on all ranks (CUDA events on the launching stream, max over ranks).

    mode: "fused" (halo rows pushed by the SpMV kernels, one launch per iteration, CUDA graph; the single-rank loop
    runs through the same native runner), "fused_multi_launch" (the same protocol as separate wait / boundary / flag /
    interior launches without a graph), "nccl_halo" (grouped NCCL send/recv of the halo blocks overlapped with the
    interior row blocks, CUDA graph), "nccl_allgather" (one grouped NCCL all-gather of x per iteration overlapped with
    the row blocks that read only local columns, CUDA graph), "fused_allgather" (the all-gather done by the SpMV kernel
    itself: every row is stored into every other GPU's copy of the next x as it is produced, the kernel waits for all
    peers' flags before its first row block and raises its own after its last; one launch per iteration, no NCCL),
    "fused_allgather_multicast" (the same with ONE store per row to an NVLink multicast address: the NVSwitch replicates
    it into every GPU's copy of the next x, so a GPU sends its slice once instead of once per peer)."""
    import torch
    from . import _lib
    dist = _dist()
    rank, world = world_info()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(4, int(warmup))
    iters = int(iters)
    fused_ag = mode in ("fused_allgather", "fused_allgather_multicast")
    exchange = "allgather" if fused_ag else {"nccl_allgather": "allgather", "nccl_halo": "halo"}.get(mode, "auto")
    loop = make_loop(shard, exchange, overlap=not fused_ag)
    native = mode in ("fused", "fused_multi_launch") and (world == 1 or loop.mode == "halo")
    dense_halo = mode in ("fused", "fused_multi_launch") and world > 1 and loop.mode == "allgather"
    if (fused_ag or dense_halo) and world > 1:
        # every rank needs (nearly) all of x: the kernels push whole slices to every peer (whole-shard schedule)
        if world - 1 > _lib.MAX_PUSH:
            if fused_ag:
                raise RuntimeError(f"fused all-gather supports at most {_lib.MAX_PUSH + 1} GPUs")
        else:
            loop.sends, loop.recvs = allgather_schedule(loop.bounds, rank)
            loop.overlapped = False
            native = True
    note = ""
    runner = None
    if native:
        flags = (_lib.HALO_NO_GRAPH | _lib.HALO_MULTI_LAUNCH) if mode == "fused_multi_launch" else 0
        try:
            if world > 1:
                runner = FusedHaloLoop(loop, shard.plan, flags, multicast=mode == "fused_allgather_multicast")
            else:
                runner = LocalLoop(loop, shard.plan, flags)
        except Exception as e:  # no peer path / IPC unavailable: keep the NCCL exchange
            if mode == "fused_allgather_multicast":
                raise  # never time a fallback under this name
            native, note = False, f"{type(e).__name__}: {e}"
    use_graph = False
    comm_sms = 0
    if native:
        run = runner.run
    else:
        if world > 1 and getattr(loop, "overlapped", False):
            # SMs left to the collective that runs beside the persistent SpMV CTAs (spmv_b200_plan_set_comm_sms). Off by
            # default: measured at 2 and 8 GPUs it does not help -- the SpMV keeps ~20 MB of loads in flight and a
            # co-running NCCL copy kernel gets bandwidth in proportion to its own few hundred KB, SMs or not
            # (profiles/r2_overlap_probe_n2.jsonl, r2_iter_modes_vs_comm_sms_n8.jsonl)
            comm_sms = int(os.environ.get("SPMV_B200_COMM_SMS", "0"))
            if comm_sms:
                shard.plan.set_comm_sms(comm_sms)
        if world > 1 and mode.startswith("fused"):
            note = note or f"the halo is not sparse for this matrix (exchange = {loop.mode}): NCCL exchange used"
        try:
            iters += iters % 2  # the captured unit is a pair of iterations (one ping-pong period of the buffers)
            warmup += warmup % 2
            loop.capture()  # two eager iterations, then the capture of two more (not executed)
            use_graph = True
            run = lambda k: loop.run_graph(k)  # noqa: E731
            warmup = max(2, warmup - 2)
        except Exception as e:
            note = (note + "; " if note else "") + f"graph capture failed ({type(e).__name__}: {e}); eager launches"
            run = lambda k: loop.run(k, finish=False)  # noqa: E731
    try:
        run(warmup)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.start()
        e0.record()
        run(iters)
        e1.record()
        e1.synchronize()
        if sampler:
            sampler.stop()
        sync()
    finally:
        if comm_sms:
            shard.plan.set_comm_sms(0)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    n, nnz_total = shard.n, float(shard.nnz_total)
    sec_iter = ms * 1e-3 / iters
    lo, hi = int(shard.bounds[rank]), int(shard.bounds[rank + 1])
    x = runner.x if native else loop.x
    out = {
        "mode": mode, "scaling": "strong", "n_gpus": world, "iters": iters, "warmup_iters": warmup + (2 if use_graph else 0),
        "ms_per_iter": sec_iter * 1e3, "value": 2.0 * nnz_total / sec_iter / 1e9, "unit": "GFLOP/s",
        "effective_gbs": (12 * nnz_total + 4 * (n + 1) + 8 * n + 16 * n) / sec_iter / 1e9,
        "exchange": ("none (one rank)" if world == 1 else
                     "all-gather by the SpMV kernel: every row stored into every peer's next x (NVLink stores), all "
                     "peers' flags waited for and raised inside the kernel"
                     + (" -- one multimem.st per row to the NVLink multicast address, replicated by the NVSwitch"
                        if mode == "fused_allgather_multicast" else "")
                     if native and (fused_ag or dense_halo) else
                     "halo rows pushed by the SpMV kernels into peer memory (NVLink stores), flags waited for and "
                     "raised inside the kernel" if native else f"NCCL {loop.mode}"),
        "exchange_bytes_in_per_iter_rank0": loop.bytes_in_per_iter,
        "exchange_overlapped_with_local_rows": bool(getattr(loop, "overlapped", False)),
        "sms_left_to_the_collective": comm_sms,
        "boundary_row_blocks_rank0": int(sum(b - a for a, b in loop.boundary)),
        # bit pattern checksum of this rank-0-owned range after the last iteration: equal values across runs with
        # 1/2/4/8 GPUs and across modes prove bitwise-identical results (n/16 rows belong to rank 0 for <= 8 ranks)
        "x_checksum_first_16th": bits_checksum(x[0:n // 16]),
        "total_timed_ms": ms,
    }
    if native:
        info = runner.loop.info()
        out.update({"launches_per_iteration": int(info.launches_per_iteration),
                    "iterations_per_graph_launch": int(info.uses_graph),
                    "single_launch_kernel": bool(info.single_launch),
                    "boundary_first": bool(getattr(runner, "split", False))})
        runner.close()
    else:
        out["cuda_graph"] = use_graph
    if note:
        out["note"] = note
    return out


def bench_power_loop(N: int = 384, iters: int = 100, mode: str = "fused", warmup: int = 4) -> dict:
    """One-call form: builds the C5 shard of this rank, times one mode, frees everything."""
    shard = build_shard("stencil3d", N=N)
    try:
        out = time_power_loop(shard, mode, iters, warmup)
        out["workload"] = f"C5: {shard.name} ({shard.n} rows, {shard.nnz_total} nnz), x <- A*x, alpha=1, beta=0"
        return out
    finally:
        shard.destroy()
