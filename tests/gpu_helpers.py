"""Shared helpers of the GPU parity tests: run the CUDA path through the C ABI and compare with the oracle."""
import numpy as np

import oracle
from spmv_acc_b200 import CsrDesc, SpmvPlan, make_options, synth

TOL = 1e-12  # north star: |y - y_ref| <= 1e-12 * (|beta*y0_i| + |alpha| * sum_j |a_ij * x_j|)


def desc_of(dcsr):
    return CsrDesc(dcsr.rows, dcsr.cols, dcsr.nnz, dcsr.rowptr, dcsr.col, dcsr.val)


def gpu_spmv(hcsr, x, y0, alpha, beta, options=None, repeat=1):
    """Returns (y, plan_info_dict). y0 is not modified."""
    import torch
    d = synth.to_device(hcsr)
    plan = SpmvPlan(desc_of(d), options)
    dx = torch.from_numpy(np.ascontiguousarray(x)).cuda() if hcsr.cols else torch.zeros(1, dtype=torch.float64, device="cuda")
    ys = []
    for _ in range(repeat):
        dy = torch.from_numpy(np.ascontiguousarray(y0)).cuda() if hcsr.rows else torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.execute(alpha, beta, dx, dy)
        torch.cuda.synchronize()
        ys.append(dy.cpu().numpy()[:hcsr.rows].copy())
    info = plan.info()
    plan.destroy()
    for y in ys[1:]:
        assert np.array_equal(y, ys[0], equal_nan=True), "run-to-run results differ bitwise"
    return ys[0], info


def assert_parity(hcsr, x, y0, alpha, beta, y, y_ref=None, what=""):
    if y_ref is None:
        y_ref = oracle.best_host_spmv(alpha, beta, hcsr.rowptr, hcsr.col, hcsr.val, x, y0)
    bound = oracle.port_row_bound(alpha, beta, hcsr.rowptr, hcsr.col, hcsr.val, x, y0)
    ok, worst, row = oracle.check_rows(y, y_ref, bound, TOL)
    assert ok, f"{what}: row {row} exceeds the fp64 bound by {worst:.3g}x (y={y[row]!r}, ref={y_ref[row]!r})"
    # the reference's own acceptance test (cli/verification.cpp:15-38) must pass as well
    r = oracle.port_verify_y(y, y_ref)
    assert r["failed_count"] == 0, f"{what}: reference verify_y failed at {r['first_failed_at']}"
