"""ORACLE — test infrastructure only: checks a sample of rows of a device-resident SpMV result against the reference's
CPU SpMV (cli/verification.cpp:56-66, compiled in place when /root/reference was available at build time, otherwise
the C restatement) at sizes where the whole matrix cannot be multiplied on the host in reasonable time.

The sampled rows are cut out of the device CSR (row pointers, column indices, values) together with exactly the
entries of x they reference, copied to the host, and multiplied there as a small CSR matrix with the same per-row
order of operations as the full matrix: host_spmv accumulates a row left to right, so a row's result does not depend
on which other rows are present. Used by tests/ (full-size parity) and by bench.py's out-of-timed-region `verified`
flag. torch is only used to gather the sample on the device.
"""
from __future__ import annotations

import numpy as np

from . import oracle as _o

TOL = 1e-12  # north star: |y - y_ref| <= 1e-12 * (|beta*y0_i| + |alpha| * sum_j |a_ij * x_j|)


def sample_rows(rows: int, count: int, seed: int = 7, must_include=()):
    """Sorted unique row ids: `count` pseudo-random rows + the first / last 64 rows + the rows in `must_include`."""
    rng = np.random.default_rng(seed)
    parts = [rng.integers(0, rows, size=min(count, rows), dtype=np.int64), np.arange(min(64, rows)),
             np.arange(max(rows - 64, 0), rows)]
    extra = np.asarray(list(must_include), dtype=np.int64)
    if extra.size:
        parts.append(extra[(extra >= 0) & (extra < rows)])
    return np.unique(np.concatenate(parts))


def cut_rows(rowptr, col, val, x, rows_idx):
    """(sub_rowptr int32, sub_col int32 into x_sub, sub_val, x_sub) on the host for the given local rows of a device
    CSR; the device rowptr may be a window (rowptr[0] != 0)."""
    import torch
    idx = torch.as_tensor(np.asarray(rows_idx, dtype=np.int64), device=rowptr.device)
    s = rowptr[idx].to(torch.int64)
    e = rowptr[idx + 1].to(torch.int64)
    lens = e - s
    sub_rp = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=rowptr.device)
    torch.cumsum(lens, 0, out=sub_rp[1:])
    total = int(sub_rp[-1])
    if total >= 2 ** 31:
        raise ValueError("sample too large")
    # element k of the sample is element s[row(k)] + (k - sub_rp[row(k)]) of the matrix
    owner = torch.repeat_interleave(torch.arange(idx.numel(), device=rowptr.device), lens)
    src = s[owner] + (torch.arange(total, device=rowptr.device) - sub_rp[owner])
    sub_col_global = col[src].to(torch.int64)
    uniq, inv = torch.unique(sub_col_global, return_inverse=True)
    return (sub_rp.to(torch.int32).cpu().numpy(), inv.to(torch.int32).cpu().numpy(), val[src].cpu().numpy(),
            x[uniq].cpu().numpy())


def check_sampled_rows(rowptr, col, val, x, y0, y, alpha, beta, rows_idx):
    """Compares y[rows_idx] (device) with host_spmv on the cut-out rows. y0 may be None (treated as zeros: beta == 0
    runs that never read y). Returns dict(ok, worst (error / bound), rows, nnz, verify_y_failed)."""
    import torch
    rows_idx = np.asarray(rows_idx, dtype=np.int64)
    sub_rp, sub_col, sub_val, x_sub = cut_rows(rowptr, col, val, x, rows_idx)
    idx = torch.as_tensor(rows_idx, device=y.device)
    y_got = y[idx].cpu().numpy()
    y0_sub = y0[idx].cpu().numpy() if y0 is not None else np.zeros(rows_idx.size)
    y_ref = _o.best_host_spmv(alpha, beta, sub_rp, sub_col, sub_val, x_sub, y0_sub)
    bound = _o.port_row_bound(alpha, beta, sub_rp, sub_col, sub_val, x_sub, y0_sub)
    ok, worst, row = _o.check_rows(y_got, y_ref, bound, TOL)
    vr = _o.port_verify_y(y_got, y_ref)  # the reference's own acceptance test (cli/verification.cpp:15-38)
    return {"ok": bool(ok and vr["failed_count"] == 0), "worst_error_over_bound": float(worst),
            "worst_row": int(rows_idx[row]) if rows_idx.size else -1, "rows_checked": int(rows_idx.size),
            "nnz_checked": int(sub_rp[-1]), "longest_row_checked": int(np.diff(sub_rp).max()) if rows_idx.size else 0,
            "verify_y_failed_count": int(vr["failed_count"]),
            "against": "reference host_spmv (oracle/_ref)" if _o.have_ref() else "C restatement of host_spmv"}
