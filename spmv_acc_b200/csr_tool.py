"""Matrix diagnostics with the output format of the reference's `csr-tool` (tools/main.cpp:117-182): `nnz` prints the
non-zeros and the average row length of every block of rows, `dist` the histogram of row lengths. They explain roofline
misses (row-length skew, empty rows) next to the bins / tile kinds of the plan.

    python -m spmv_acc_b200.csr_tool {nnz|dist} <matrix> [-f csr|bin2|mtx] [-p PARTS]
"""
from __future__ import annotations

import argparse
from typing import List

import numpy as np

from . import formats


def part_nnz_lines(rowptr: np.ndarray, parts: int = 0) -> List[str]:
    """tools/main.cpp:125-157. parts == 0 means one part per row; a part holds ceil(m / parts) rows."""
    rp = np.asarray(rowptr, dtype=np.int64)
    m = rp.size - 1
    out = ["[part ID] [part nnz] [avg-nnz/row]"]
    if m <= 0:
        return out
    parts = m if parts == 0 else parts
    rows_per_part = m // parts + (0 if m % parts == 0 else 1)
    full = m // rows_per_part
    for i in range(full):
        a, b = rp[i * rows_per_part], rp[(i + 1) * rows_per_part]
        out.append(f"{i} {int(b - a)} {_fmt((b - a) / rows_per_part)}")
    if full != parts and m - full * rows_per_part > 0:  # the last, shorter part (tools/main.cpp:152-156)
        a, rows = rp[full * rows_per_part], m - full * rows_per_part
        out.append(f"{full} {int(rp[m] - a)} {_fmt((rp[m] - a) / rows)}")
    return out


def dist_lines(rowptr: np.ndarray) -> List[str]:
    """tools/main.cpp:159-181: "<row length> = <number of rows>", ascending row length."""
    lens = np.diff(np.asarray(rowptr, dtype=np.int64))
    k, c = np.unique(lens, return_counts=True)
    return [f"{int(a)} = {int(b)}" for a, b in zip(k, c)]


def _fmt(v: float) -> str:
    # operator<< of a double: 6 significant digits, no trailing zeros
    return f"{v:.6g}"


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="csr-tool", description="csr analyzing tool.")
    ap.add_argument("mode", choices=["nnz", "dist"])
    ap.add_argument("matrix")
    ap.add_argument("-f", "--format", default="csr", choices=["csr", "bin2", "mtx"])
    ap.add_argument("-p", "--parts", type=int, default=0)
    a = ap.parse_args(argv)
    csr, _ = formats.load(a.matrix, a.format)
    lines = part_nnz_lines(csr.rowptr, a.parts) if a.mode == "nnz" else dist_lines(csr.rowptr)
    print("\n".join(lines))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
