"""Synthetic generators: numpy restatements (CPU) and, on the GPU, bit-equality of the CUDA generators with them."""
import numpy as np
import pytest

from spmv_acc_b200 import synth


def test_stencil2d_shape_and_values():
    c = synth.stencil2d_numpy(16)
    assert c.rows == 256 and c.nnz == 5 * 256 - 4 * 16
    assert np.all(np.diff(c.rowptr) >= 3) and np.all(np.diff(c.rowptr) <= 5)
    # the Laplacian of the constant vector vanishes in the interior
    y = np.add.reduceat(c.val, c.rowptr[:-1])
    assert y.reshape(16, 16)[1:-1, 1:-1].max() == 0.0


def test_stencil3d_nnz_formula_and_row_sums():
    N = 7
    c = synth.stencil3d_numpy(N)
    assert c.nnz == (3 * N - 2) ** 3
    y = np.add.reduceat(c.val, c.rowptr[:-1]).reshape(N, N, N)
    assert np.allclose(y[1:-1, 1:-1, 1:-1], 1.0, atol=1e-15)


def test_uniform_has_exactly_k_distinct_sorted_columns():
    c = synth.uniform_numpy(50, 97, 32, seed=1)
    cols = c.col.reshape(50, 32)
    assert np.all(np.diff(cols, axis=1) > 0) and cols.min() >= 0 and cols.max() < 97
    assert np.all(np.abs(c.val) <= 1.0)


def test_rmat_is_power_law_with_empty_rows():
    c = synth.rmat_numpy(12, 16, seed=1)
    lens = np.diff(c.rowptr)
    assert c.nnz == 16 << 12 and lens.max() > 40 * lens.mean() and (lens == 0).sum() > c.rows // 10
    assert np.all(c.col >= 0) and np.all(c.col < c.cols)


def test_row_range_generation_is_a_slice_of_the_full_matrix():
    full = synth.stencil3d_numpy(6)
    part = synth.stencil3d_numpy(6, 50, 140)
    lo, hi = full.rowptr[50], full.rowptr[140]
    assert np.array_equal(part.rowptr, full.rowptr[50:141] - lo)
    assert np.array_equal(part.col, full.col[lo:hi]) and np.array_equal(part.val, full.val[lo:hi])


def test_circuit_standin_has_the_documented_shape():
    c = synth.circuit_numpy()
    assert (c.rows, c.cols, c.nnz) == (7602, 7602, 32653)  # examples/batch.sh:51-52
    assert np.diff(c.rowptr).max() >= 60


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["stencil2d", "stencil3d", "uniform", "rmat", "vector", "shard"])
def test_device_generators_match_numpy(name):
    import torch
    if name == "stencil2d":
        d, h = synth.stencil2d_device(37), synth.stencil2d_numpy(37)
    elif name == "stencil3d":
        d, h = synth.stencil3d_device(11), synth.stencil3d_numpy(11)
    elif name == "uniform":
        d, h = synth.uniform_device(300, 1000, 32, seed=1), synth.uniform_numpy(300, 1000, 32, seed=1)
    elif name == "rmat":
        d, h = synth.rmat_device(11, 16, seed=1), synth.rmat_numpy(11, 16, seed=1)
    elif name == "shard":
        d, h = synth.stencil3d_device(9, 100, 400), synth.stencil3d_numpy(9, 100, 400)
    else:
        assert np.array_equal(synth.vector_device(1000, 2).cpu().numpy(), synth.vector_numpy(1000, 2))
        return
    torch.cuda.synchronize()
    assert (d.rows, d.cols, d.nnz) == (h.rows, h.cols, h.nnz)
    assert np.array_equal(d.rowptr.cpu().numpy(), h.rowptr)
    assert np.array_equal(d.col.cpu().numpy(), h.col)
    assert np.array_equal(d.val.cpu().numpy(), h.val)
