"""Synthetic CSR matrices for the BASELINE.json configurations.

Every generator exists twice: ``*_numpy`` (CPU, for oracles / fixtures / small tests) and ``*_device`` (CUDA kernels
of ``libspmv_b200_gen.so`` + torch scans/sorts, for sizes that do not fit comfortably on the host). Both follow the
same counter-based specification and produce bit-identical arrays (tests/test_synth.py).

Shapes:
  C1 ``circuit``     rajat03-shaped 7602x7602 stand-in (the real file is a git-LFS stub in the reference)
  C2 ``stencil2d``   5-point Laplacian on an N x N grid, diag 4, off-diagonals -1
  C3 ``uniform``     m x n, exactly k distinct uniformly drawn columns per row
  C4 ``rmat``        R-MAT scale s, edge factor f, (a,b,c,d) = (0.57,0.19,0.19,0.05), duplicates kept
  C5 ``stencil3d``   27-point averaging stencil on an N^3 grid, value 1/27
Row ranges [r_lo, r_hi) generate a contiguous row shard with a rowptr rebased to 0 (global column indices).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Tuple

import numpy as np

from . import _lib

_U = np.uint64
_MASK = (1 << 64) - 1


# ----------------------------------------------------------------------------------------------------------------
# counter-based RNG (restates mix64 / hash3 / sym_unit of csrc/gen.cu)
# ----------------------------------------------------------------------------------------------------------------
def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + _U(0x9E3779B97F4A7C15)).astype(_U)
        z = ((z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)).astype(_U)
        z = ((z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)).astype(_U)
        return (z ^ (z >> _U(31))).astype(_U)


def hash3(seed: int, a, b) -> np.ndarray:
    a = np.asarray(a, dtype=_U)
    b = np.asarray(b, dtype=_U)
    with np.errstate(over="ignore"):
        s = _U((int(seed) * 0xD1342543DE82EF95) & _MASK)
        return _mix64(_mix64((s + a).astype(_U)) ^ (b * _U(0x2545F4914F6CDD1D)).astype(_U))


def sym_unit(h: np.ndarray) -> np.ndarray:
    return (h >> _U(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def vector_numpy(n: int, seed: int) -> np.ndarray:
    return sym_unit(hash3(seed, np.arange(n, dtype=_U), _U(0x5EED)))


@dataclass
class Csr:
    rows: int
    cols: int
    rowptr: Any   # int32 [rows+1]
    col: Any      # int32 [nnz]
    val: Any      # float64 [nnz]

    @property
    def nnz(self) -> int:
        return int(self.val.shape[0])


def _rowptr_from_counts(counts: np.ndarray) -> np.ndarray:
    rp = np.zeros(counts.size + 1, dtype=np.int64)
    np.cumsum(counts, out=rp[1:])
    assert rp[-1] < 2 ** 31
    return rp.astype(np.int32)


# ----------------------------------------------------------------------------------------------------------------
# numpy generators
# ----------------------------------------------------------------------------------------------------------------
def stencil2d_numpy(N: int, r_lo: int = 0, r_hi: Optional[int] = None, NY: Optional[int] = None) -> Csr:
    """5-point Laplacian on a grid of NY rows x N points (NY defaults to N)."""
    NY = N if NY is None else NY
    r_hi = NY * N if r_hi is None else r_hi
    r = np.arange(r_lo, r_hi, dtype=np.int64)
    i, j = r // N, r % N
    offs = [(-N, i > 0, -1.0), (-1, j > 0, -1.0), (0, np.ones_like(i, bool), 4.0), (1, j < N - 1, -1.0),
            (N, i < NY - 1, -1.0)]
    counts = sum(m.astype(np.int64) for _, m, _ in offs)
    rowptr = _rowptr_from_counts(counts)
    col = np.zeros(int(rowptr[-1]), np.int32)
    val = np.zeros(int(rowptr[-1]), np.float64)
    pos = rowptr[:-1].astype(np.int64).copy()
    for off, mask, v in offs:
        idx = pos[mask]
        col[idx] = (r[mask] + off).astype(np.int32)
        val[idx] = v
        pos[mask] += 1
    return Csr(r_hi - r_lo, NY * N, rowptr, col, val)


def stencil3d_numpy(N: int, r_lo: int = 0, r_hi: Optional[int] = None) -> Csr:
    r_hi = N ** 3 if r_hi is None else r_hi
    r = np.arange(r_lo, r_hi, dtype=np.int64)
    x, y, z = r % N, (r // N) % N, r // (N * N)
    masks = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = ((z + dz >= 0) & (z + dz < N) & (y + dy >= 0) & (y + dy < N) & (x + dx >= 0) & (x + dx < N))
                masks.append(((dz * N + dy) * N + dx, ok))
    counts = sum(m.astype(np.int64) for _, m in masks)
    rowptr = _rowptr_from_counts(counts)
    col = np.zeros(int(rowptr[-1]), np.int32)
    val = np.full(int(rowptr[-1]), 1.0 / 27.0, np.float64)
    pos = rowptr[:-1].astype(np.int64).copy()
    for off, mask in masks:
        col[pos[mask]] = (r[mask] + off).astype(np.int32)
        pos[mask] += 1
    return Csr(r_hi - r_lo, N ** 3, rowptr, col, val)


def uniform_numpy(m: int, n: int, k: int, seed: int = 1, r_lo: int = 0, r_hi: Optional[int] = None) -> Csr:
    """Exactly k distinct columns per row; candidates hash3(seed,row,t) % n for t = 0,1,.. (duplicates skipped)."""
    r_hi = m if r_hi is None else r_hi
    assert 1 <= k <= 64 and k <= n
    rows = r_hi - r_lo
    col = np.zeros(rows * k, np.int32)
    for li, r in enumerate(range(r_lo, r_hi)):
        have: list = []
        t = 0
        while len(have) < k:
            batch = (hash3(seed, _U(r), np.arange(t, t + 2 * k, dtype=_U)) >> _U(11)) % _U(n)
            for cand in batch.tolist():
                t += 1
                if cand not in have:
                    have.append(cand)
                    if len(have) == k:
                        break
        col[li * k:(li + 1) * k] = np.sort(np.array(have, dtype=np.int64)).astype(np.int32)
    rr = np.repeat(np.arange(r_lo, r_hi, dtype=_U), k)
    qq = np.tile(np.arange(k, dtype=_U), rows)
    val = sym_unit(hash3(seed ^ 0xABCDEF, rr, qq))
    rowptr = (np.arange(rows + 1, dtype=np.int64) * k).astype(np.int32)
    return Csr(rows, n, rowptr, col, val)


RMAT_ABC = (0.57, 0.19, 0.19)


def rmat_numpy(scale: int, edge_factor: int = 16, seed: int = 1, abc: Tuple[float, float, float] = RMAT_ABC) -> Csr:
    m = 1 << scale
    ne = edge_factor * m
    a, b, c = abc
    e = np.arange(ne, dtype=_U)
    row = np.zeros(ne, _U)
    colv = np.zeros(ne, _U)
    for lvl in range(scale):
        u = (hash3(seed, e, _U(lvl)) >> _U(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        quad = np.where(u < a, 0, np.where(u < a + b, 1, np.where(u < a + b + c, 2, 3))).astype(_U)
        row = (row << _U(1)) | (quad >> _U(1))
        colv = (colv << _U(1)) | (quad & _U(1))
    keys = np.sort(((row << _U(32)) | colv).astype(np.int64))
    col = (keys & 0xFFFFFFFF).astype(np.int32)
    val = sym_unit(hash3(seed ^ 0x1234567, np.arange(ne, dtype=_U), _U(1)))
    rowptr = np.searchsorted(keys, np.arange(m + 1, dtype=np.int64) << 32, side="left").astype(np.int32)
    return Csr(m, m, rowptr, col, val)


def circuit_numpy(n: int = 7602, target_nnz: int = 32653, seed: int = 20230616) -> Csr:
    """rajat03-shaped stand-in: diagonal + a few random off-diagonals per row + a handful of dense-ish rows,
    trimmed/padded to exactly ``target_nnz`` entries; values U(-1,1); columns ascending within a row."""
    rows = np.arange(n, dtype=_U)
    extra = (hash3(seed, rows, _U(7)) % _U(5)).astype(np.int64) + 1          # 1..5 off-diagonals
    dense_rows = (hash3(seed, np.arange(12, dtype=_U), _U(99)) % _U(n)).astype(np.int64)
    extra[dense_rows] = 60 + (hash3(seed, dense_rows.astype(_U), _U(5)) % _U(40)).astype(np.int64)
    cols_per_row = []
    for r in range(n):
        cand = (hash3(seed, _U(r), np.arange(int(extra[r]) * 2 + 4, dtype=_U)) % _U(n)).astype(np.int64)
        s = {r}
        for cnd in cand.tolist():
            if len(s) >= extra[r] + 1:
                break
            s.add(cnd)
        cols_per_row.append(sorted(s))
    # adjust to the exact nnz of the real matrix: drop / add off-diagonals deterministically from the end
    total = sum(len(c) for c in cols_per_row)
    r = n - 1
    while total > target_nnz:
        if len(cols_per_row[r]) > 1:
            drop = [cc for cc in cols_per_row[r] if cc != r][-1]
            cols_per_row[r].remove(drop)
            total -= 1
        r = r - 1 if r > 0 else n - 1
    r = 0
    while total < target_nnz:
        cand = (r * 7919 + 13) % n
        if cand not in cols_per_row[r]:
            cols_per_row[r] = sorted(cols_per_row[r] + [cand])
            total += 1
        r = (r + 1) % n
    counts = np.array([len(c) for c in cols_per_row], dtype=np.int64)
    rowptr = _rowptr_from_counts(counts)
    col = np.concatenate([np.array(c, dtype=np.int32) for c in cols_per_row])
    val = sym_unit(hash3(seed ^ 0x51, np.arange(col.size, dtype=_U), _U(3)))
    return Csr(n, n, rowptr, col, val)


# ----------------------------------------------------------------------------------------------------------------
# device generators (torch tensors on the current CUDA device)
# ----------------------------------------------------------------------------------------------------------------
def _stream() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)


def _ck(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed with CUDA status {rc}")


def vector_device(n: int, seed: int):
    import torch
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    _ck(_lib.gen().spmv_b200_gen_vector(n, seed, out.data_ptr(), _stream()), "gen_vector")
    return out


def _rowptr_from_counts_device(counts):
    import torch
    rp = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts, 0, out=rp[1:])
    if int(rp[-1]) >= 2 ** 31:
        raise ValueError("matrix too large for int32 indices")
    return rp.to(torch.int32)


def _stencil_device(kind: str, dims: tuple, total_rows: int, r_lo: int, r_hi: Optional[int]) -> Csr:
    import torch
    r_hi = total_rows if r_hi is None else r_hi
    rows = r_hi - r_lo
    G = _lib.gen()
    counts = torch.empty(rows, dtype=torch.int32, device="cuda")
    _ck(getattr(G, f"spmv_b200_gen_{kind}_counts")(*dims, r_lo, r_hi, counts.data_ptr(), _stream()), "gen counts")
    rowptr = _rowptr_from_counts_device(counts)
    del counts
    nnz = int(rowptr[-1])
    col = torch.empty(nnz, dtype=torch.int32, device="cuda")
    val = torch.empty(nnz, dtype=torch.float64, device="cuda")
    _ck(getattr(G, f"spmv_b200_gen_{kind}_fill")(*dims, r_lo, r_hi, rowptr.data_ptr(), col.data_ptr(),
                                                  val.data_ptr(), _stream()), "gen fill")
    return Csr(rows, total_rows, rowptr, col, val)


def stencil2d_device(N: int, r_lo: int = 0, r_hi: Optional[int] = None, NY: Optional[int] = None) -> Csr:
    NY = N if NY is None else NY
    return _stencil_device("stencil2d", (N, NY), NY * N, r_lo, r_hi)


def stencil3d_device(N: int, r_lo: int = 0, r_hi: Optional[int] = None) -> Csr:
    return _stencil_device("stencil3d", (N,), N ** 3, r_lo, r_hi)


def stencil_row_counts_device(kind: str, N: int, NY: Optional[int] = None):
    """int32 nnz-per-row of the whole stencil matrix (used to place nnz-balanced shard boundaries)."""
    import torch
    dims = (N, N if NY is None else NY) if kind == "stencil2d" else (N,)
    total = dims[0] * dims[1] if kind == "stencil2d" else N ** 3
    counts = torch.empty(total, dtype=torch.int32, device="cuda")
    _ck(getattr(_lib.gen(), f"spmv_b200_gen_{kind}_counts")(*dims, 0, total, counts.data_ptr(), _stream()), "counts")
    return counts


def uniform_device(m: int, n: int, k: int, seed: int = 1, r_lo: int = 0, r_hi: Optional[int] = None) -> Csr:
    import torch
    r_hi = m if r_hi is None else r_hi
    rows = r_hi - r_lo
    if rows * k >= 2 ** 31:
        raise ValueError("matrix too large for int32 indices")
    rowptr = (torch.arange(rows + 1, dtype=torch.int64, device="cuda") * k).to(torch.int32)
    col = torch.empty(rows * k, dtype=torch.int32, device="cuda")
    val = torch.empty(rows * k, dtype=torch.float64, device="cuda")
    _ck(_lib.gen().spmv_b200_gen_uniform_fill(r_lo, r_hi, n, k, seed, col.data_ptr(), val.data_ptr(), _stream()),
        "gen_uniform_fill")
    return Csr(rows, n, rowptr, col, val)


def rmat_device(scale: int, edge_factor: int = 16, seed: int = 1,
                abc: Tuple[float, float, float] = RMAT_ABC) -> Csr:
    import torch
    m = 1 << scale
    ne = edge_factor * m
    if ne >= 2 ** 31:
        raise ValueError("matrix too large for int32 indices")
    keys = torch.empty(ne, dtype=torch.int64, device="cuda")
    G = _lib.gen()
    _ck(G.spmv_b200_gen_rmat_edges(scale, ne, abc[0], abc[1], abc[2], seed, keys.data_ptr(), _stream()), "rmat_edges")
    keys = torch.sort(keys)[0]
    rowptr = torch.empty(m + 1, dtype=torch.int32, device="cuda")
    col = torch.empty(ne, dtype=torch.int32, device="cuda")
    val = torch.empty(ne, dtype=torch.float64, device="cuda")
    _ck(G.spmv_b200_gen_rmat_finish(m, ne, keys.data_ptr(), seed, rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                    _stream()), "rmat_finish")
    torch.cuda.current_stream().synchronize()
    return Csr(m, m, rowptr, col, val)


def to_device(csr: Csr) -> Csr:
    import torch
    return Csr(csr.rows, csr.cols, torch.from_numpy(np.ascontiguousarray(csr.rowptr)).cuda(),
               torch.from_numpy(np.ascontiguousarray(csr.col)).cuda(),
               torch.from_numpy(np.ascontiguousarray(csr.val)).cuda())


def to_host(csr: Csr) -> Csr:
    return Csr(csr.rows, csr.cols, csr.rowptr.cpu().numpy(), csr.col.cpu().numpy(), csr.val.cpu().numpy())


def algorithmic_bytes(m: int, n: int, nnz: int) -> int:
    """Compulsory HBM bytes of one SpMV (BASELINE.json): 12*nnz + 4*(m+1) + 8*n + 16*m."""
    return 12 * nnz + 4 * (m + 1) + 8 * n + 16 * m


# ----------------------------------------------------------------------------------------------------------------
# SuiteSparse-shaped stand-ins (the matrices of the reference's evaluation, examples/large-data-set-batch.sh:23-52, are
# not shipped and there is no network): same rows / columns / average row length, a structure of the same family.
# Device generators only (torch RNG with a fixed seed); used by the selector study, not by the headline benchmark.
# ----------------------------------------------------------------------------------------------------------------
SUITESPARSE_SHAPES = {
    # name: (rows, cols, nnz per row, family)
    "boneS10": (914_898, 914_898, 30.81, "fem3"),           # 3 dof per node, banded FEM
    "Bump_2911": (2_911_419, 2_911_419, 22.44, "fem3"),
    "Cube_Coup_dt6": (2_164_760, 2_164_760, 29.88, "fem3"),
    "dielFilterV3real": (1_102_824, 1_102_824, 40.99, "fem1"),
    "Ga41As41H72": (268_096, 268_096, 34.98, "skewed"),     # quantum chemistry: a few hundred long rows
    "Hardesty3": (8_217_820, 7_591_564, 4.92, "short_rect"),
    "largebasis": (440_020, 440_020, 12.64, "fem1"),
    "RM07R": (381_689, 381_689, 98.16, "blocks"),           # CFD, dense blocks of ~100
    "TSOPF_RS_b2383": (38_120, 38_120, 424.22, "blocks"),   # power network, dense blocks of ~400
    "vas_stokes_2M": (2_146_677, 2_146_677, 30.34, "skewed"),
}


def suitesparse_like_device(name: str, seed: int = 1, device: str = "cuda", shrink: int = 1) -> Csr:
    """`shrink` divides rows and columns (host-side tests of the generator itself)."""
    import torch
    m, n, avg, family = SUITESPARSE_SHAPES[name]
    m, n = max(64, m // shrink), max(64, n // shrink)
    g = torch.Generator(device=device)
    g.manual_seed(seed * 7919 + sum(map(ord, name)))
    dev = device
    if family in ("fem1", "fem3", "blocks"):
        # rows of a node share a block of consecutive columns around the diagonal plus blocks further away (a banded
        # graph with `reach` neighbours per node, `dof` unknowns per node)
        dof = {"fem1": 1, "fem3": 3, "blocks": 1}[family]
        per = max(1, int(round(avg / dof)))                  # neighbour nodes per row
        width = {"fem1": 6, "fem3": 4, "blocks": int(avg)}[family]
        nodes = (m + dof - 1) // dof
        r = torch.arange(m, device=dev, dtype=torch.int64)
        node = r // dof
        if family == "blocks":
            # one dense run of `avg` columns per row, starting at a block boundary
            start = (node // width) * width - width // 2 + torch.randint(0, width, (m,), generator=g, device=dev)
            start = start.clamp(0, max(n - width, 0))
            lens = torch.full((m,), min(width, n), dtype=torch.int64, device=dev)
            rowptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
            torch.cumsum(lens, 0, out=rowptr[1:])
            owner = torch.repeat_interleave(r, lens)
            col = start[owner] + (torch.arange(int(rowptr[-1]), device=dev) - rowptr[owner])
        else:
            # neighbour nodes: the node itself, +-1.. in the same "line", and +-stride lines (a 3-D mesh numbering)
            stride1 = max(2, int(round(nodes ** (1.0 / 3.0))))
            stride2 = stride1 * stride1
            offs = [0]
            k = 1
            while len(offs) < per:
                for o in (k, -k, k * stride1, -k * stride1, k * stride2, -k * stride2):
                    if len(offs) < per:
                        offs.append(o)
                k += 1
            offs = torch.tensor(sorted(offs), device=dev, dtype=torch.int64)
            nb = (node[:, None] + offs[None, :])                      # [m, per] neighbour node ids
            ok = (nb >= 0) & (nb < nodes)
            cols = (nb[:, :, None] * dof + torch.arange(dof, device=dev)[None, None, :]).reshape(m, per * dof)
            okc = ok[:, :, None].expand(m, per, dof).reshape(m, per * dof) & (cols < n)
            lens = okc.sum(1)
            rowptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
            torch.cumsum(lens, 0, out=rowptr[1:])
            col = cols[okc]
    elif family == "short_rect":
        lens = torch.randint(3, 8, (m,), generator=g, device=dev)          # 3..7, mean 5
        rowptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens, 0, out=rowptr[1:])
        nnz = int(rowptr[-1])
        owner = torch.repeat_interleave(torch.arange(m, device=dev), lens)
        centre = (owner.double() * (n / m)).long()
        col = (centre + torch.randint(-2000, 2001, (nnz,), generator=g, device=dev)).clamp(0, n - 1)
    else:  # skewed: log-normal row lengths with a heavy tail, columns clustered around the diagonal + a random part
        z = torch.randn(m, generator=g, device=dev, dtype=torch.float64)
        lens = torch.exp(0.9 * z)
        lens = (lens * (avg / float(lens.mean()))).round().long().clamp(1, n)
        heavy = torch.randint(0, m, (max(1, m // 2000),), generator=g, device=dev)
        lens[heavy] = (lens[heavy] * 40).clamp(max=min(n, 20000))
        rowptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens, 0, out=rowptr[1:])
        nnz = int(rowptr[-1])
        owner = torch.repeat_interleave(torch.arange(m, device=dev), lens)
        near = owner + torch.randint(-300, 301, (nnz,), generator=g, device=dev)
        far = torch.randint(0, n, (nnz,), generator=g, device=dev)
        pick = torch.rand(nnz, generator=g, device=dev) < 0.8
        col = torch.where(pick, near, far).clamp(0, n - 1)
    nnz = int(rowptr[-1])
    if nnz >= 2 ** 31:
        raise ValueError("matrix too large for int32 indices")
    # sort the columns inside every row (the reference's readers produce sorted rows, cli/sparse_format.h:105-112)
    owner = torch.repeat_interleave(torch.arange(m, device=dev), (rowptr[1:] - rowptr[:-1]))
    key = owner * n + col
    key = torch.sort(key)[0]
    col = (key % n).to(torch.int32)
    val = torch.rand(nnz, generator=g, device=dev, dtype=torch.float64) * 2.0 - 1.0
    return Csr(m, n, rowptr.to(torch.int32), col, val)
