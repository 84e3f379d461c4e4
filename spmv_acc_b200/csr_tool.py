"""Matrix diagnostics with the output format of the reference's `csr-tool` (tools/main.cpp:117-182): `nnz` prints the
non-zeros and the average row length of every block of rows, `dist` the histogram of row lengths. They explain roofline
misses (row-length skew, empty rows) next to the bins / tile kinds of the plan.

    python -m spmv_acc_b200.csr_tool {nnz|dist|plan} <matrix> [-f csr|bin2|mtx] [-p PARTS] [--device]

`plan` prints what the device analysis pass of the engine (csrc/analysis.cu) found for the matrix — rows and non-zeros
per length bin, row blocks per kernel kind, split rows, the gather-coalescing statistic and the form the plan took —
and `--device` computes the `nnz` / `dist` tables on the GPU from the uploaded row pointers (matrices of BASELINE size
take seconds on the host). Both need a GPU; without the flag `nnz` / `dist` are pure host code.
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


def device_rowptr(rowptr):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("--device / plan need a CUDA device")
    return torch.as_tensor(np.asarray(rowptr, dtype=np.int64)).cuda() if not hasattr(rowptr, "is_cuda") else rowptr.long()


def part_nnz_lines_device(rowptr, parts: int = 0) -> List[str]:
    """part_nnz_lines with the row pointers on the device (same output, line for line)."""
    import torch
    rp = device_rowptr(rowptr)
    m = rp.numel() - 1
    out = ["[part ID] [part nnz] [avg-nnz/row]"]
    if m <= 0:
        return out
    parts = m if parts == 0 else parts
    rows_per_part = m // parts + (0 if m % parts == 0 else 1)
    cuts = torch.arange(0, m + 1, rows_per_part, device=rp.device)
    if int(cuts[-1]) != m and m // rows_per_part != parts:
        cuts = torch.cat([cuts, torch.tensor([m], device=rp.device)])
    nnz = (rp[cuts[1:]] - rp[cuts[:-1]]).cpu().numpy()
    rows = (cuts[1:] - cuts[:-1]).cpu().numpy()
    out += [f"{i} {int(a)} {_fmt(a / r)}" for i, (a, r) in enumerate(zip(nnz, rows))]
    return out


def dist_lines_device(rowptr) -> List[str]:
    import torch
    rp = device_rowptr(rowptr)
    k, c = torch.unique(rp[1:] - rp[:-1], return_counts=True)
    return [f"{int(a)} = {int(b)}" for a, b in zip(k.cpu().numpy(), c.cpu().numpy())]


def plan_lines(csr) -> List[str]:
    """The device analysis of the engine for this matrix (spmv_b200_plan_get_info), as text."""
    import torch
    from . import CsrDesc, SpmvPlan
    dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).cuda()  # noqa: E731
    rowptr, col, val = dev(csr.rowptr, np.int32), dev(csr.col, np.int32), dev(csr.val, np.float64)
    plan = SpmvPlan(CsrDesc(csr.rows, csr.cols, int(csr.nnz), rowptr, col, val))
    i = plan.info()
    plan.destroy()
    bins = ("short (<= %d)" % i.short_max, "medium (<= %d)" % i.medium_max, "long (<= %d)" % i.tile_nnz,
            "very long (> %d)" % i.tile_nnz)
    out = [f"rows {i.m} cols {i.n} nnz {i.nnz} avg-nnz/row {_fmt(i.nnz / max(i.m, 1))}",
           "[row-length bin] [rows] [nnz]"]
    out += [f"{b}: {int(r)} {int(z)}" for b, r, z in zip(bins, i.bin_rows, i.bin_nnz)]
    out.append(f"row blocks of {i.tile_nnz} items: {i.ntiles} (SHORT {i.tiles_per_kind[0]}, MEDIUM {i.tiles_per_kind[1]}, "
               f"MIXED {i.tiles_per_kind[2]}), rows split across blocks: {i.nsplit_rows}")
    out.append(f"lines of x per gather of 32 consecutive rows (sampled): {_fmt(i.gather_lines / max(i.gather_active, 1))}")
    form = ("direct (one warp per row block, no shared memory)" if i.direct else
            f"staged x, persistent ring {i.ring_ctas} CTAs/SM x {i.ring_stages} stages" if i.xstage and i.ring_ctas else
            "staged x, one row block per CTA" if i.xstage else "tiled, x gathered per element")
    out.append(f"form: {form}; kernel launches per SpMV: {i.launches_per_execute}; plan memory: {i.workspace_bytes} bytes")
    return out


def _fmt(v: float) -> str:
    # operator<< of a double: 6 significant digits, no trailing zeros
    return f"{v:.6g}"


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="csr-tool", description="csr analyzing tool.")
    ap.add_argument("mode", choices=["nnz", "dist", "plan"])
    ap.add_argument("matrix")
    ap.add_argument("-f", "--format", default="csr", choices=["csr", "bin2", "mtx"])
    ap.add_argument("-p", "--parts", type=int, default=0)
    ap.add_argument("--device", action="store_true", help="compute the nnz / dist tables on the GPU")
    a = ap.parse_args(argv)
    csr, _ = formats.load(a.matrix, a.format)
    if a.mode == "plan":
        lines = plan_lines(csr)
    elif a.device:
        lines = part_nnz_lines_device(csr.rowptr, a.parts) if a.mode == "nnz" else dist_lines_device(csr.rowptr)
    else:
        lines = part_nnz_lines(csr.rowptr, a.parts) if a.mode == "nnz" else dist_lines(csr.rowptr)
    print("\n".join(lines))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
