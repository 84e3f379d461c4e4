"""On-disk matrix formats of the reference CLI: readers and writers.

  * ``.csr`` text  — 5 lines: free-form header | nnz values | nnz column indices | rows+1 row offsets | cols entries
    of x, numbers separated by single spaces (cli/csr_mtx_reader.hpp:49-111).
  * ``bin2``       — little-endian int32 magic 0x20211015 | version 2 | value type (1 pattern, 2 int, 3 real,
    4 complex) | rows | cols | nnz | row_ptr | col_index | values (cli/csr_binary_reader.hpp:37-101; writer spec
    tools/suitesparse-dl/conv/conv.go:121-193).
  * MatrixMarket   — ``%%MatrixMarket matrix coordinate {real|integer|pattern|complex} {general|symmetric|Hermitian}``,
    1-based, symmetric off-diagonals mirrored, entries sorted by (row, col), duplicates kept
    (cli/matrix_market_reader.hpp:50-303, cli/sparse_format.h:100-128).

tests/test_formats.py checks every reader against the reference's own reader compiled in place (oracle/_ref).
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

from .synth import Csr

BIN2_MAGIC = 0x20211015
BIN2_VERSION = 2
TP_BOOL, TP_INT, TP_FLOAT, TP_COMPLEX = 1, 2, 3, 4


def _fmt_floats(a: np.ndarray) -> str:
    return " ".join(repr(float(v)) for v in a)


def write_csr_text(path, csr: Csr, x: np.ndarray, header: str = "csr matrix") -> None:
    rp, col, val = np.asarray(csr.rowptr), np.asarray(csr.col), np.asarray(csr.val)
    with open(path, "w") as f:
        f.write(header.replace("\n", " ") + "\n")
        f.write(_fmt_floats(val) + "\n")
        f.write(" ".join(str(int(c)) for c in col) + "\n")
        f.write(" ".join(str(int(r)) for r in rp) + "\n")
        f.write(_fmt_floats(np.asarray(x)) + "\n")


def read_csr_text(path) -> Tuple[Csr, np.ndarray]:
    """rows = len(line 3) - 1, cols = len(line 4), nnz = len(line 1) (cli/csr_mtx_reader.hpp:107-111)."""
    with open(path, "r") as f:
        lines = f.read().split("\n")
    if len(lines) < 5:
        raise ValueError(f"{path}: a .csr file needs 5 lines")

    def nums(line, dtype):
        toks = [t for t in line.split(" ") if t != ""]
        # the reference parses every token with atof (csr_mtx_reader.hpp:153-158), then narrows
        return np.array(toks, dtype=np.float64).astype(dtype)  # one vectorised parse per line

    val = nums(lines[1], np.float64)
    col = nums(lines[2], np.int32)
    rowptr = nums(lines[3], np.int32)
    x = nums(lines[4], np.float64)
    return Csr(rowptr.size - 1, x.size, rowptr, col, val), x


def write_bin2(path, csr: Csr, val_type: int = TP_FLOAT) -> None:
    rp = np.ascontiguousarray(csr.rowptr, dtype="<i4")
    col = np.ascontiguousarray(csr.col, dtype="<i4")
    with open(path, "wb") as f:
        f.write(struct.pack("<6i", BIN2_MAGIC, BIN2_VERSION, val_type, int(csr.rows), int(csr.cols), int(col.size)))
        f.write(rp.tobytes())
        f.write(col.tobytes())
        if val_type == TP_BOOL:
            pass
        elif val_type == TP_INT:
            f.write(np.ascontiguousarray(csr.val, dtype="<i4").tobytes())
        else:
            f.write(np.ascontiguousarray(csr.val, dtype="<f8").tobytes())


def read_bin2(path) -> Csr:
    data = Path(path).read_bytes()
    if len(data) < 24:
        raise ValueError(f"{path}: truncated bin2 header")
    magic, version, val_type, rows, cols, nnz = struct.unpack_from("<6i", data, 0)
    if magic != BIN2_MAGIC:
        raise ValueError(f"{path}: mismatch magic number")
    if version != BIN2_VERSION:
        raise ValueError(f"{path}: only bin file version 2 is supported")
    if val_type not in (TP_BOOL, TP_INT, TP_FLOAT, TP_COMPLEX):
        raise ValueError(f"{path}: matrix value type not supported")
    off = 24
    rowptr = np.frombuffer(data, dtype="<i4", count=rows + 1, offset=off).astype(np.int32)
    off += 4 * (rows + 1)
    col = np.frombuffer(data, dtype="<i4", count=nnz, offset=off).astype(np.int32)
    off += 4 * nnz
    if val_type == TP_BOOL:
        val = np.ones(nnz, dtype=np.float64)
    elif val_type == TP_INT:
        val = np.frombuffer(data, dtype="<i4", count=nnz, offset=off).astype(np.float64)
    else:
        val = np.frombuffer(data, dtype="<f8", count=nnz, offset=off).astype(np.float64)
    return Csr(rows, cols, rowptr, col, val)


def coo_to_csr(rows: int, cols: int, r: np.ndarray, c: np.ndarray, v: np.ndarray) -> Csr:
    """Sort by (row, col), count rows, prefix-sum (cli/sparse_format.h:100-128). Duplicates are kept."""
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    rowptr = np.zeros(rows + 1, dtype=np.int64)
    np.add.at(rowptr, r.astype(np.int64) + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return Csr(rows, cols, rowptr, c.astype(np.int32), v.astype(np.float64))


def _read_mtx_coo(path):
    """Header checks, 1-based -> 0-based, symmetric / Hermitian mirroring: the parsing half of
    cli/matrix_market_reader.hpp:50-303. Returns (rows, cols, row, col, val) in file order."""
    with open(path, "r") as f:
        first = f.readline()
        if not first.startswith("%%MatrixMarket matrix coordinate"):
            raise ValueError("Can only read MatrixMarket format that is in coordinate form")
        toks = first.split()
        field, sym = toks[3], toks[4]
        pattern = field == "pattern"
        if field not in ("real", "integer", "pattern", "complex", "double"):
            raise ValueError("MatrixMarket data type does not match matrix format")
        if sym not in ("general", "symmetric", "Hermitian"):
            raise ValueError("Can only read MatrixMarket format that is either symmetric, general or hermitian")
        mirrored = sym != "general"
        line = f.readline()
        while line and line.startswith("%"):
            line = f.readline()
        rows, cols, nz = (int(t) for t in line.split()[:3])
        body = f.read()
    # fast path: every entry line has the same number of tokens (2 pattern, 3 real / integer, 4 complex) and there are
    # no comment lines after the size line -> one vectorised parse instead of a Python loop per entry
    ntok = 2 if pattern else (4 if field == "complex" else 3)
    flat = None
    if "%" not in body:
        try:
            flat = np.array(body.split(), dtype=np.float64)
        except ValueError:
            flat = None
        if flat is not None and (flat.size % ntok != 0):
            flat = None
    if flat is not None:
        ent = flat.reshape(-1, ntok)
        r1, c1 = ent[:, 0].astype(np.int64), ent[:, 1].astype(np.int64)
        v1 = np.ones(r1.size) if pattern else ent[:, 2].copy()  # complex keeps only the first value token
        if r1.size and (r1.max() > rows or c1.max() > cols):
            raise ValueError("index out of bounds in matrix market file")
        if mirrored:
            # the mirrored copy of an off-diagonal entry follows the entry itself, like the reference reader emits it
            off = r1 != c1
            reps = np.where(off, 2, 1)
            pos = np.cumsum(reps) - reps           # position of each original entry in the output
            total = int(reps.sum())
            rr = np.empty(total, np.int64)
            cc = np.empty(total, np.int64)
            vv = np.empty(total, np.float64)
            rr[pos], cc[pos], vv[pos] = r1 - 1, c1 - 1, v1
            rr[pos[off] + 1], cc[pos[off] + 1], vv[pos[off] + 1] = c1[off] - 1, r1[off] - 1, v1[off]
        else:
            rr, cc, vv = r1 - 1, c1 - 1, v1
        return rows, cols, rr, cc, vv
    rr, cc, vv = [], [], []
    for line in body.split("\n"):
        if not line.strip() or line.startswith("%"):
            continue
        p = line.split()
        r, c = int(p[0]), int(p[1])
        if r > rows or c > cols:
            raise ValueError("index out of bounds in matrix market file")
        v = 1.0 if pattern else float(p[2])  # complex keeps only the first value token
        rr.append(r - 1); cc.append(c - 1); vv.append(v)
        if mirrored and r != c:
            rr.append(c - 1); cc.append(r - 1); vv.append(v)
    return rows, cols, np.array(rr, dtype=np.int64), np.array(cc, dtype=np.int64), np.array(vv, dtype=np.float64)


def read_mtx(path) -> Csr:
    return coo_to_csr(*_read_mtx_coo(path))


def read_mtx_device(path):
    """MatrixMarket file -> CsrDesc with device arrays: the text is parsed on the host, the COO -> CSR conversion (sort by
    (row, col), row offsets) runs on the GPU (spmv_b200_coo_to_csr)."""
    import torch
    from . import api
    rows, cols, r, c, v = _read_mtx_coo(path)
    return api.coo_to_csr(rows, cols, torch.from_numpy(r.astype(np.int32)).cuda(),
                          torch.from_numpy(c.astype(np.int32)).cuda(), torch.from_numpy(v).cuda())


def write_mtx(path, csr: Csr, symmetric_lower_only: bool = False) -> None:
    rp, col, val = np.asarray(csr.rowptr), np.asarray(csr.col), np.asarray(csr.val)
    rows_of = np.repeat(np.arange(csr.rows), np.diff(rp))
    keep = col <= rows_of if symmetric_lower_only else np.ones(col.size, bool)
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real {'symmetric' if symmetric_lower_only else 'general'}\n")
        f.write("% written by spmv_acc_b200.formats\n")
        f.write(f"{csr.rows} {csr.cols} {int(keep.sum())}\n")
        for r, c, v in zip(rows_of[keep], col[keep], val[keep]):
            f.write(f"{int(r) + 1} {int(c) + 1} {float(v)!r}\n")


def load(path, fmt: str = "csr") -> Tuple[Csr, Optional[np.ndarray]]:
    if fmt == "csr":
        return read_csr_text(path)
    if fmt == "bin2":
        return read_bin2(path), None
    if fmt == "mtx":
        return read_mtx(path), None
    raise ValueError("unsupported matrix format.")
