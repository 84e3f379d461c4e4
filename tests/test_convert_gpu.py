"""COO -> CSR on the device (the step in front of the hot path for MatrixMarket input) against the host conversion and
against the reference's own MatrixMarket reader compiled in place (cli/matrix_market_reader.hpp, cli/sparse_format.h)."""
import numpy as np
import pytest

import oracle
from spmv_acc_b200 import SpmvB200Error, coo_to_csr, formats, synth

pytestmark = pytest.mark.gpu


def _to_host(d):
    return (d.row_ptr.cpu().numpy(), d.col_index.cpu().numpy(), d.values.cpu().numpy())


@pytest.mark.parametrize("rows,cols,nnz,seed", [(1, 1, 1, 0), (50, 7, 0, 1), (300, 200, 5000, 2), (70000, 65000, 400000, 3),
                                                (5, 100000, 30000, 4)])
def test_device_conversion_equals_host_conversion(rows, cols, nnz, seed):
    import torch
    rng = np.random.default_rng(seed)
    r = rng.integers(0, rows, nnz).astype(np.int32)
    c = rng.integers(0, cols, nnz).astype(np.int32)
    if nnz > 10:  # duplicates of (row, col) with different values: input order must be kept (stable)
        r[5:10], c[5:10] = r[0], c[0]
    v = rng.standard_normal(nnz)
    ref = formats.coo_to_csr(rows, cols, r.astype(np.int64), c.astype(np.int64), v)
    d = coo_to_csr(rows, cols, torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(v).cuda())
    rp, col, val = _to_host(d)
    assert np.array_equal(rp, ref.rowptr) and np.array_equal(col, ref.col) and np.array_equal(val, ref.val)


def test_mtx_file_through_device_conversion_equals_reference_reader(tmp_path):
    h = synth.rmat_numpy(11, 8, seed=3)
    # R-MAT keeps duplicate (row, col) pairs, whose order after the reference's std::sort is unspecified: merge them
    keys = np.repeat(np.arange(h.rows, dtype=np.int64), np.diff(h.rowptr)) * h.cols + h.col
    uniq, first = np.unique(keys, return_index=True)
    h = formats.coo_to_csr(h.rows, h.cols, uniq // h.cols, uniq % h.cols, h.val[first])
    for sym in (False, True):
        path = tmp_path / f"m_{int(sym)}.mtx"
        formats.write_mtx(path, h, symmetric_lower_only=sym)
        d = formats.read_mtx_device(path)
        rp, col, val = _to_host(d)
        host = formats.read_mtx(path)
        assert np.array_equal(rp, host.rowptr) and np.array_equal(col, host.col) and np.array_equal(val, host.val)
        if oracle.have_ref():
            r_rp, r_col, r_val, _, r_cols = oracle.ref_read(str(path), "mtx")
            assert (d.rows, d.cols, d.nnz) == (r_rp.size - 1, r_cols, r_col.size)
            assert np.array_equal(rp, r_rp) and np.array_equal(col, r_col) and np.array_equal(val, r_val)


def test_out_of_range_index_is_rejected():
    import torch
    r = torch.tensor([0, 3], dtype=torch.int32, device="cuda")
    c = torch.tensor([0, 1], dtype=torch.int32, device="cuda")
    v = torch.ones(2, dtype=torch.float64, device="cuda")
    with pytest.raises(SpmvB200Error):
        coo_to_csr(3, 2, r, c, v)
