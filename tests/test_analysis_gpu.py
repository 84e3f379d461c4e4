"""GPU parity tests of the row analysis: every exported integer array must equal the CPU restatement bit for bit,
and the tile partition must equal the output of the reference's own merge-path `partition` kernel."""
import ctypes
from pathlib import Path

import numpy as np
import pytest

import oracle
from gpu_helpers import desc_of
from spmv_acc_b200 import SpmvPlan, make_options, shard_bounds, synth

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
ARRAYS = ["tile_row", "tile_elem", "tile_split", "tile_kind", "tile_part", "tile_maxlen", "row_bin", "split_rows"]


def _ragged(seed, m, choices):
    rng = np.random.default_rng(seed)
    lens = rng.choice(choices, size=m)
    rp = np.zeros(m + 1, np.int32)
    rp[1:] = np.cumsum(lens)
    nnz = int(rp[-1])
    return synth.Csr(m, 1000, rp, rng.integers(0, 1000, nnz).astype(np.int32), rng.standard_normal(nnz))


def _matrices():
    yield "stencil2d", synth.stencil2d_numpy(100)
    yield "stencil3d", synth.stencil3d_numpy(20)
    yield "uniform", synth.uniform_numpy(700, 900, 32, seed=1)
    yield "rmat", synth.rmat_numpy(13, 16, seed=1)
    yield "circuit", synth.circuit_numpy()
    yield "ragged", _ragged(1, 3000, [0, 0, 1, 2, 3, 5, 9, 17, 40, 300, 700, 5000, 20000])
    yield "empty_rows", synth.Csr(50, 7, np.zeros(51, np.int32), np.zeros(0, np.int32), np.zeros(0))
    yield "one_row", _ragged(2, 1, [100000])


@pytest.mark.parametrize("opts", [(0, 0, 0), (256, 4, 16), (512, 8, 64), (4096, 16, 256)])
def test_analysis_arrays_bit_exact(opts):
    T, S, L = opts
    for name, h in _matrices():
        d = synth.to_device(h)
        plan = SpmvPlan(desc_of(d), make_options(T, S, L))
        info = plan.info()
        ref = oracle.port_analysis(h.rowptr, info.tile_nnz, info.short_max, info.medium_max)
        assert info.ntiles == ref["ntiles"], name
        for a in ARRAYS:
            got = plan.export(a)
            assert np.array_equal(got, ref[a]), f"{name}: {a} differs (T={info.tile_nnz})"
        assert list(info.bin_rows) == ref["bin_rows"].tolist(), name
        assert list(info.bin_nnz) == ref["bin_nnz"].tolist(), name
        assert info.nsplit_rows == ref["nsplit"]
        if h.nnz > 0:
            assert (info.gather_active, info.gather_lines) == oracle.port_gather_stat(h.rowptr, h.col, info.medium_max)[:2], name
        kinds = np.bincount(ref["tile_kind"], minlength=3).tolist()
        assert list(info.tiles_per_kind) == kinds
        plan.destroy()


@pytest.mark.parametrize("opts", [(0, 0, 0), (256, 4, 16), (1024, 8, 64)])
def test_direct_form_arrays_bit_exact(opts):
    """Row-start bit flags, non-empty row list and per-tile base ordinal of the direct (warp-per-row-block) form."""
    from spmv_acc_b200 import FLAG_DIRECT
    T, S, L = opts
    for name, h in _matrices():
        d = synth.to_device(h)
        plan = SpmvPlan(desc_of(d), make_options(T, S, L, flags=FLAG_DIRECT))
        info = plan.info()
        if h.nnz == 0:
            assert info.direct == 0, name
            plan.destroy()
            continue
        assert info.direct == 1 and (T or info.tile_nnz == 2048), name
        ref = oracle.port_analysis(h.rowptr, info.tile_nnz, info.short_max, info.medium_max)
        assert np.array_equal(plan.export("tile_row"), ref["tile_row"]), name
        extra = oracle.port_direct_arrays(h.rowptr, ref["tile_row"])
        for a in ("row_start_bits", "nz_rows", "tile_nzbase"):
            assert np.array_equal(plan.export(a), extra[a]), f"{name}: {a} differs"
        plan.destroy()


def _banded(seed, m, half_width, per_row):
    """rows with `per_row` columns inside [r - half_width, r + half_width]: one run of lines per row block"""
    rng = np.random.default_rng(seed)
    cols = []
    for r in range(m):
        lo, hi = max(0, r - half_width), min(m, r + half_width + 1)
        cols.append(np.sort(rng.choice(np.arange(lo, hi), size=min(per_row, hi - lo), replace=False)))
    rp = np.zeros(m + 1, np.int32)
    rp[1:] = np.cumsum([c.size for c in cols])
    col = np.concatenate(cols).astype(np.int32)
    return synth.Csr(m, m, rp, col, rng.standard_normal(col.size))


def test_staged_x_arrays_bit_exact_and_only_for_regular_matrices():
    """lcol (16-bit local column indices) and the per-row-block segment tables of the staged-x form equal the CPU
    restatement (oracle/analysis_port.c:port_xstage); stencils and banded matrices take the form, matrices with
    scattered columns, split rows or mixed row blocks do not; SPMV_B200_FLAG_NO_XSTAGE switches it off."""
    from spmv_acc_b200 import FLAG_NO_XSTAGE
    takes = {"stencil2d": synth.stencil2d_numpy(100), "stencil3d": synth.stencil3d_numpy(20),
             "stencil3d_odd_n": synth.stencil3d_numpy(15), "banded": _banded(5, 4000, 40, 9),
             "stencil2d_rows_1000_3000": synth.stencil2d_numpy(100, 1000, 3000)}
    for name, h in takes.items():
        for T in (0, 512, 1792):
            d = synth.to_device(h)
            plan = SpmvPlan(desc_of(d), make_options(T))
            info = plan.info()
            assert info.xstage == 1 and info.xstage_lines > 0, (name, T, list(info.tiles_per_kind))
            ref = oracle.port_xstage(h.col, plan.export("tile_elem"))
            assert ref["failed"] == 0 and ref["max_lines"] == info.xstage_lines, name
            assert np.array_equal(plan.export("lcol"), ref["lcol"]), f"{name}: lcol differs (T={info.tile_nnz})"
            assert np.array_equal(plan.export("xdesc"), ref["xdesc"]), f"{name}: xdesc differs (T={info.tile_nnz})"
            plan.destroy()
            plan = SpmvPlan(desc_of(d), make_options(T, flags=FLAG_NO_XSTAGE))
            assert plan.info().xstage == 0 and plan.export("lcol").size == 0
            plan.destroy()
    for name, h in _matrices():
        if name in ("stencil2d", "stencil3d") or h.nnz == 0:
            continue
        d = synth.to_device(h)
        plan = SpmvPlan(desc_of(d))
        info = plan.info()
        if info.xstage:  # small matrices with few columns may qualify; then the arrays must still be exact
            ref = oracle.port_xstage(h.col, plan.export("tile_elem"))
            assert ref["failed"] == 0 and np.array_equal(plan.export("lcol"), ref["lcol"]), name
        else:
            assert name in ("uniform", "rmat", "ragged", "one_row", "circuit"), name
        plan.destroy()


def _mesh_like(m, strides, reach, seed):
    """Every row references itself +-1..reach along each of the given strides: many short runs of columns a few lines
    of x apart (the shape an unstructured-mesh matrix has after a bandwidth-reducing ordering)."""
    offs = sorted({0} | {sg * k * st for st in strides for k in range(1, reach + 1) for sg in (1, -1)})
    r = np.arange(m, dtype=np.int64)[:, None] + np.asarray(offs, dtype=np.int64)[None, :]
    ok = (r >= 0) & (r < m)
    rp = np.zeros(m + 1, dtype=np.int32)
    rp[1:] = np.cumsum(ok.sum(1))
    col = r[ok].astype(np.int32)
    rng = np.random.default_rng(seed)
    return synth.Csr(m, m, rp, col, rng.standard_normal(col.size))


def test_staged_x_merges_runs_of_lines_across_small_gaps():
    """Row blocks with more than 16 runs of x lines qualify when merging runs at most g lines apart (g = 1, 2, 4 ...
    32) leaves at most 16 segments of at most 256 lines in total: segment tables and 16-bit indices equal the CPU
    restatement bit for bit, the staged copy then holds lines no element references, and y is still right."""
    import torch
    from gpu_helpers import assert_parity
    cases = {"strides_1_200_reach_10_T2048": _mesh_like(50000, (1, 200), 10, 1),
             "strides_1_120_14400_reach_7": _mesh_like(120000, (1, 120, 14400), 7, 2),
             "strides_1_97_9409_reach_6": _mesh_like(30000, (1, 97, 9409), 6, 3),
             "strides_1_200_reach_10_T1024": _mesh_like(50000, (1, 200), 10, 4)}
    merged_somewhere = False
    for name, h in cases.items():
        d = synth.to_device(h)
        plan = SpmvPlan(desc_of(d), make_options(2048 if name.endswith("T2048") else 1024 if name.endswith("T1024") else 0))
        info = plan.info()
        ref = oracle.port_xstage(h.col, plan.export("tile_elem"))
        assert (ref["failed"] == 0) == bool(info.xstage), (name, ref["failed"], info.xstage)
        if info.xstage:
            assert ref["max_lines"] == info.xstage_lines, name
            assert np.array_equal(plan.export("xdesc"), ref["xdesc"]), f"{name}: xdesc differs"
            assert np.array_equal(plan.export("lcol"), ref["lcol"]), f"{name}: lcol differs"
            xd = ref["xdesc"].reshape(-1, 32)
            # a merged tile stages more lines than its elements reference
            te = plan.export("tile_elem")
            for t in range(0, xd.shape[0], max(1, xd.shape[0] // 50)):
                referenced = np.unique(h.col[te[t]:te[t + 1]] >> 4).size
                merged_somewhere |= xd[t, 1] > referenced
                assert xd[t, 1] >= referenced and 1 <= xd[t, 0] <= 16
            x, y0 = synth.vector_numpy(h.cols, 2), synth.vector_numpy(h.rows, 3)
            dy = torch.from_numpy(y0).cuda()
            plan.execute(0.75, -0.5, torch.from_numpy(x).cuda(), dy)
            torch.cuda.synchronize()
            assert_parity(h, x, y0, 0.75, -0.5, dy.cpu().numpy(), what=f"merged staged x {name}")
        plan.destroy()
    assert merged_somewhere, "no case exercised the merge path"


def test_tile_partition_equals_reference_merge_path_partition_kernel():
    """TILE_PART (T = 2048) against the reference's `partition` kernel, compiled in place into oracle/_ref/libref_gpu.so
    (benchmark/merge-path/merge_path_partition.h:7-17; launch shape of merge_path_spmv.cu:44)."""
    import torch
    so = ROOT / "oracle" / "_ref" / "libref_gpu.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libref_gpu.so not built")
    ref = ctypes.CDLL(str(so))
    ref.ref_gpu_merge_path_partition_2048.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    for name, h in _matrices():
        if h.nnz == 0:
            continue
        d = synth.to_device(h)
        plan = SpmvPlan(desc_of(d), make_options(2048))
        nt = plan.info().ntiles
        S = torch.zeros(nt + 1, dtype=torch.int32, device="cuda")
        assert ref.ref_gpu_merge_path_partition_2048(d.rowptr.data_ptr(), h.rows, nt + 1, S.data_ptr()) == 0
        ours = plan.export("tile_part")
        theirs = S.cpu().numpy()
        # the reference searches for idx*2048 without clamping to nnz; only the last entry can differ (target > nnz)
        assert np.array_equal(ours[:-1], theirs[:-1]), name
        assert np.array_equal(theirs, oracle.port_merge_path_partition(h.rowptr, nt + 1, 2048)), name
        plan.destroy()


@pytest.mark.parametrize("nshards", [1, 2, 3, 8])
def test_shard_bounds_bit_exact(nshards):
    for name, h in _matrices():
        d = synth.to_device(h)
        got = shard_bounds(d.rowptr, h.rows, nshards)
        assert np.array_equal(got, oracle.port_shard_bounds(h.rowptr, nshards)), name


def test_full_size_partition_properties_c2():
    """BASELINE config C2 (4096^2 grid): size-independent properties of the analysis at full size."""
    d = synth.stencil2d_device(4096)
    plan = SpmvPlan(desc_of(d))
    info = plan.info()
    assert info.m == 16777216 and info.nnz == 83869696
    assert info.ntiles == -(-(info.nnz + info.m) // info.tile_nnz)
    assert list(info.tiles_per_kind) == [info.ntiles, 0, 0] and info.nsplit_rows == 0
    assert list(info.bin_rows) == [info.m, 0, 0, 0]
    te, tr = plan.export("tile_elem").astype(np.int64), plan.export("tile_row").astype(np.int64)
    assert te[0] == 0 and te[-1] == info.nnz and np.all(np.diff(te) > 0) and np.all(np.diff(te) <= info.tile_nnz + 8)
    assert tr[0] == 0 and tr[-1] == info.m and np.all(np.diff(tr) > 0)
    rp = d.rowptr.cpu().numpy()
    assert np.array_equal(rp[tr], te)            # every tile starts exactly on a row boundary
    plan.destroy()
