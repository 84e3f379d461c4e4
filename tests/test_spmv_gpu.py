"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle (the reference's own
host_spmv where it was compiled here, else its plain-C restatement) on identical inputs.

Bar: per row |y - y_ref| <= 1e-12 * (|beta*y0_i| + |alpha| * sum_j |a_ij * x_j|)  (north star), plus the reference's
own verify_y (rel 1e-7, cli/verification.cpp:15-38), plus bitwise run-to-run reproducibility."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN_CASES, load_golden
from gpu_helpers import assert_parity, desc_of, gpu_spmv
from spmv_acc_b200 import (FLAG_BETA0_SKIP_Y, FLAG_DIRECT, FLAG_NO_DIRECT, FLAG_NO_TMA, FLAG_NO_XSTAGE, CsrDesc, HostMatrix, SpmvB200Error, SpmvPlan,
                           cache_invalidate, cache_revalidations, cache_size, host_spmv, make_options, sparse_csr_spmv,
                           sparse_spmv, synth)

pytestmark = pytest.mark.gpu

AB = [(1.0, 1.0), (0.75, -0.5), (1.0, 0.0), (0.0, 2.0), (-1.25, 1e-3)]
OPTS = {
    "default": None,
    "no_tma": make_options(flags=FLAG_NO_TMA),
    "small_tiles": make_options(256, 4, 16, 2, flags=FLAG_NO_DIRECT),
    "big_tiles": make_options(8192, 16, 256, 16, flags=FLAG_NO_DIRECT),
    "no_staged_x": make_options(flags=FLAG_NO_XSTAGE),
    "staged_x_no_ring": make_options(flags=1 << 25),
    "staged_x_no_ring_40_registers": make_options(flags=(1 << 25) | (1 << 24)),
    "staged_x_small_tiles": make_options(512, 4, 16, 4),
    "staged_x_T1792_4_lanes": make_options(1792, 0, 0, 8),
    "tiled_only": make_options(flags=FLAG_NO_DIRECT),
    "direct": make_options(flags=FLAG_DIRECT),
    "direct_scalar_loads": make_options(flags=FLAG_DIRECT | FLAG_NO_TMA),
    "direct_40_registers": make_options(flags=FLAG_DIRECT | (1 << 23)),
    "direct_T512_L64": make_options(512, 8, 64, flags=FLAG_DIRECT),
    "short_w8": make_options(flags=1 << 8),
    "short_w4_medium_w4": make_options(flags=(2 << 8) | (1 << 12)),
}


def _golden_csr(g):
    return synth.Csr(int(g["rows"]), int(g["cols"]), g["rowptr"], g["col"], g["val"])


@pytest.mark.parametrize("opt", list(OPTS))
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors(name, opt):
    """Committed fixtures: inputs and expected y produced by the reference's host_spmv (tests/golden/make_golden.py)."""
    g = load_golden(name)
    h = _golden_csr(g)
    for (a, b), y_ref in zip(g["ab"], g["y"]):
        y, _ = gpu_spmv(h, g["x"], g["y0"], float(a), float(b), OPTS[opt], repeat=2)
        assert_parity(h, g["x"], g["y0"], float(a), float(b), y, y_ref, what=f"{name}/{opt}/a={a},b={b}")


def _ragged(seed, m, n, choices):
    rng = np.random.default_rng(seed)
    lens = rng.choice(choices, size=m)
    rp = np.zeros(m + 1, np.int32)
    rp[1:] = np.cumsum(lens)
    nnz = int(rp[-1])
    return synth.Csr(m, n, rp, rng.integers(0, n, nnz).astype(np.int32), rng.standard_normal(nnz))


def _families():
    yield "stencil2d_200", synth.stencil2d_numpy(200)
    yield "stencil3d_24", synth.stencil3d_numpy(24)
    yield "stencil3d_15_odd_n", synth.stencil3d_numpy(15)          # n odd: the last entry of x cannot travel by TMA
    yield "stencil2d_rows_shard", synth.stencil2d_numpy(120, 3000, 9000)  # a row shard: columns outside its rows
    yield "uniform_3000x5000_32", synth.uniform_numpy(3000, 5000, 32, seed=1)
    yield "rmat_s14", synth.rmat_numpy(14, 16, seed=1)
    yield "ragged_mixed", _ragged(3, 5000, 3000, [0, 0, 1, 2, 3, 5, 9, 17, 40, 130, 300, 700, 5000])
    yield "giant_rows", _ragged(4, 40, 4096, [0, 1, 3, 60000, 250000])
    yield "all_len_1", _ragged(5, 9000, 100, [1])
    yield "many_empty", _ragged(6, 20000, 50, [0, 0, 0, 0, 0, 0, 0, 2])
    yield "medium_129", _ragged(7, 2000, 7000, [127, 128, 129, 130])


@pytest.mark.parametrize("opt", list(OPTS))
def test_seeded_families_against_oracle(opt):
    for name, h in _families():
        x = synth.vector_numpy(h.cols, 2)
        y0 = synth.vector_numpy(h.rows, 3)
        for a, b in AB[:3]:
            y, info = gpu_spmv(h, x, y0, a, b, OPTS[opt], repeat=2)
            assert_parity(h, x, y0, a, b, y, what=f"{name}/{opt}/a={a},b={b}")


@pytest.mark.parametrize("ctas,stages", [("1", "0"), ("1", "2"), ("2", "2"), ("2", "3"), ("3", "0")])
def test_staged_x_ring_geometries(ctas, stages, monkeypatch):
    """The persistent ring form of the staged-x kernels with other CTAs-per-SM / stages-per-CTA than the default, on
    whole matrices and on row-block ranges (each range is its own launch of the ring)."""
    import torch
    monkeypatch.setenv("SPMV_B200_RING_CTAS", ctas)
    monkeypatch.setenv("SPMV_B200_RING_STAGES", stages)
    for name, h in (("stencil3d_30", synth.stencil3d_numpy(30)), ("stencil2d_150", synth.stencil2d_numpy(150)),
                    ("stencil3d_15_odd_n", synth.stencil3d_numpy(15))):
        x, y0 = synth.vector_numpy(h.cols, 2), synth.vector_numpy(h.rows, 3)
        for a, b in AB[:2]:
            y, info = gpu_spmv(h, x, y0, a, b, None, repeat=2)
            assert info.xstage == 1 and info.ring_stages >= 2 and info.ring_ctas >= 1, (info.ring_ctas, info.ring_stages)
            if stages != "0":
                assert info.ring_stages <= int(stages)
            assert_parity(h, x, y0, a, b, y, what=f"ring {ctas}x{stages} {name} a={a} b={b}")
        d = synth.to_device(h)
        p = SpmvPlan(desc_of(d))
        nt = p.info().ntiles
        dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y0).cuda()
        cuts = [0, nt // 3, nt // 3 + 1, nt]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            p.execute_tiles(0.75, -0.5, dx, dy, lo, hi)
        torch.cuda.synchronize()
        assert_parity(h, x, y0, 0.75, -0.5, dy.cpu().numpy(), what=f"ring {ctas}x{stages} {name} tile ranges")
        p.destroy()


def test_sms_left_to_a_collective_do_not_change_results():
    """spmv_b200_plan_set_comm_sms: the persistent form launches fewer CTAs (row blocks are dealt to fewer of them); every
    row must come out bitwise identical, for whole launches and row-block ranges. Plans not in the persistent form ignore
    the setting."""
    import torch
    for name, h in (("stencil3d_40", synth.stencil3d_numpy(40)), ("stencil2d_300", synth.stencil2d_numpy(300)),
                    ("rmat_12", synth.rmat_numpy(12, 16, seed=1))):
        x, y0 = synth.vector_numpy(h.cols, 2), synth.vector_numpy(h.rows, 3)
        d = synth.to_device(h)
        p = SpmvPlan(desc_of(d))
        nt = p.info().ntiles
        dx = torch.from_numpy(x).cuda()
        outs = []
        for sms in (0, 16, 140, 1000):
            p.set_comm_sms(sms)
            dy = torch.from_numpy(y0).cuda()
            p.execute(0.75, -0.5, dx, dy)
            if p.info().nsplit_rows == 0:
                dy2 = torch.from_numpy(y0).cuda()
                for lo, hi in ((0, nt // 2), (nt // 2, nt)):
                    p.execute_tiles(0.75, -0.5, dx, dy2, lo, hi)
                torch.cuda.synchronize()
                assert torch.equal(dy, dy2), (name, sms)
            torch.cuda.synchronize()
            outs.append(dy.cpu().numpy())
        assert_parity(h, x, y0, 0.75, -0.5, outs[0], what=f"comm_sms {name}")
        for o in outs[1:]:
            assert np.array_equal(o, outs[0]), name
        with pytest.raises(Exception):
            p.set_comm_sms(-1)
        p.destroy()


def test_kinds_are_exercised():
    """Each per-bin kernel, the staged-x form of both row kernels and the fix-up pass must actually run somewhere in
    this suite."""
    seen = np.zeros(3, dtype=np.int64)
    splits = 0
    staged = set()
    for name, h in _families():
        d = synth.to_device(h)
        p = SpmvPlan(desc_of(d))
        i = p.info()
        seen += np.array(list(i.tiles_per_kind))
        splits += i.nsplit_rows
        assert i.uses_tma == 1
        if i.xstage:
            staged.add(int(np.argmax(list(i.tiles_per_kind))))
        p.destroy()
    assert np.all(seen > 0) and splits > 0
    assert staged == {0, 1}, f"staged-x form seen for tile kinds {staged} (want SHORT and MEDIUM)"


def test_edge_shapes():
    import torch
    # m = 0
    p = SpmvPlan(CsrDesc(0, 5, 0, torch.zeros(1, dtype=torch.int32, device="cuda"),
                         torch.zeros(1, dtype=torch.int32, device="cuda"),
                         torch.zeros(1, dtype=torch.float64, device="cuda")))
    p.execute(1.0, 1.0, torch.zeros(5, dtype=torch.float64, device="cuda"),
              torch.zeros(1, dtype=torch.float64, device="cuda"))
    assert p.info().ntiles == 0
    p.destroy()
    # nnz = 0, m > 0: y = beta * y
    h = synth.Csr(300, 4, np.zeros(301, np.int32), np.zeros(0, np.int32), np.zeros(0))
    y0 = np.arange(300.0)
    y, _ = gpu_spmv(h, np.ones(4), y0, 3.0, -2.0)
    assert np.array_equal(y, -2.0 * y0)
    # exactly one tile boundary inside a row, nnz an exact multiple of the tile size
    h = _ragged(9, 8, 64, [512])
    x, y0 = synth.vector_numpy(64, 1), synth.vector_numpy(8, 2)
    y, info = gpu_spmv(h, x, y0, 1.0, 1.0, make_options(256, 8, 16))
    assert_parity(h, x, y0, 1.0, 1.0, y, what="exact multiple")
    assert info.nsplit_rows == 8


def test_beta_zero_semantics():
    """Default follows the oracle (0 * NaN = NaN, cli/verification.cpp:64); the opt-in flag skips the read of y."""
    h = synth.stencil2d_numpy(20)
    x = synth.vector_numpy(h.cols, 2)
    y0 = synth.vector_numpy(h.rows, 3)
    y0[7] = np.nan
    y, _ = gpu_spmv(h, x, y0, 1.0, 0.0)
    y_ref = oracle.best_host_spmv(1.0, 0.0, h.rowptr, h.col, h.val, x, y0)
    assert np.isnan(y[7]) and np.isnan(y_ref[7])
    assert np.array_equal(np.isnan(y), np.isnan(y_ref))
    y2, _ = gpu_spmv(h, x, y0, 1.0, 0.0, make_options(flags=FLAG_BETA0_SKIP_Y))
    assert not np.isnan(y2).any()
    y0[7] = 0.0
    assert_parity(h, x, y0, 1.0, 0.0, y2, what="beta0 skip")


def test_rowptr_view_and_misaligned_pointers():
    """rowptr may be a window of a larger matrix (rowptr[0] != 0) and value/colindex may be unaligned for TMA."""
    import torch
    h = _ragged(12, 4000, 2500, [0, 1, 2, 3, 5, 9, 17, 40, 300, 3000])
    x = synth.vector_numpy(h.cols, 2)
    d = synth.to_device(h)
    dx = torch.from_numpy(x).cuda()
    lo, hi = 1234, 3456
    y0 = synth.vector_numpy(hi - lo, 3)
    dy = torch.from_numpy(y0).cuda()
    view = CsrDesc(hi - lo, h.cols, int(h.rowptr[hi] - h.rowptr[lo]), d.rowptr[lo:hi + 1], d.col, d.val)
    p = SpmvPlan(view, make_options(256, 8, 16))
    p.execute(0.5, 2.0, dx, dy)
    torch.cuda.synchronize()
    sub = synth.Csr(hi - lo, h.cols, h.rowptr[lo:hi + 1] - h.rowptr[lo], h.col[h.rowptr[lo]:h.rowptr[hi]],
                    h.val[h.rowptr[lo]:h.rowptr[hi]])
    assert_parity(sub, x, y0, 0.5, 2.0, dy.cpu().numpy(), what="rowptr view")
    p.destroy()
    # misaligned value / colindex pointers -> the plan must fall back to plain loads by itself
    col_pad = torch.zeros(h.nnz + 1, dtype=torch.int32, device="cuda")
    val_pad = torch.zeros(h.nnz + 1, dtype=torch.float64, device="cuda")
    col_pad[1:] = d.col
    val_pad[1:] = d.val
    mis = CsrDesc(h.rows, h.cols, h.nnz, d.rowptr, col_pad[1:], val_pad[1:])
    p = SpmvPlan(mis)
    assert p.info().uses_tma == 0
    y0 = synth.vector_numpy(h.rows, 5)
    dy = torch.from_numpy(y0).cuda()
    p.execute(1.0, 1.0, dx, dy)
    torch.cuda.synchronize()
    assert_parity(h, x, y0, 1.0, 1.0, dy.cpu().numpy(), what="misaligned")
    p.destroy()


def test_staged_x_on_row_pointer_windows():
    """A row shard given as a window of the parent's row pointers (rowptr[lo : hi + 1], colindex / value of the parent):
    the staged-x form indexes its 16-bit column copy from the window's first element; windows whose first row pointer
    is not 16-byte aligned cannot take the ring form (its row pointers travel by TMA) and must still be right."""
    import torch
    h = synth.stencil3d_numpy(28)
    x = synth.vector_numpy(h.cols, 2)
    d = synth.to_device(h)
    dx = torch.from_numpy(x).cuda()
    for lo, hi in ((4000, 15000), (4001, 15003), (4002, 21952), (0, 7777)):
        y0 = synth.vector_numpy(hi - lo, 3)
        dy = torch.from_numpy(y0).cuda()
        view = CsrDesc(hi - lo, h.cols, int(h.rowptr[hi] - h.rowptr[lo]), d.rowptr[lo:hi + 1], d.col, d.val)
        p = SpmvPlan(view)
        assert p.info().xstage == 1
        p.execute(0.75, -0.5, dx, dy)
        torch.cuda.synchronize()
        sub = synth.Csr(hi - lo, h.cols, h.rowptr[lo:hi + 1] - h.rowptr[lo], h.col[h.rowptr[lo]:h.rowptr[hi]],
                        h.val[h.rowptr[lo]:h.rowptr[hi]])
        assert_parity(sub, x, y0, 0.75, -0.5, dy.cpu().numpy(), what=f"staged x on rows {lo}:{hi}")
        ref = oracle.port_xstage(h.col, p.export("tile_elem"))
        assert np.array_equal(p.export("lcol"), ref["lcol"])
        p.destroy()


def test_reference_shaped_entry_points_and_plan_cache():
    """sparse_csr_spmv / sparse_spmv keep the reference's argument lists (src/acc/api/spmv.h:20-28)."""
    import torch
    cache_invalidate()
    g = load_golden("c1_circuit")
    h = _golden_csr(g)
    d = synth.to_device(h)
    desc = desc_of(d)
    dx = torch.from_numpy(g["x"]).cuda()
    # the CLI's call pattern: alpha = beta = 1, y0 copied in before every call (cli/main.cpp:94-118)
    for _ in range(3):
        dy = torch.from_numpy(g["y0"]).cuda()
        sparse_csr_spmv(0, 1.0, 1.0, desc.as_const(), desc.as_const(), dx, dy)
    torch.cuda.synchronize()
    assert cache_size() == 1
    assert_parity(h, g["x"], g["y0"], 1.0, 1.0, dy.cpu().numpy(), g["y"][0], what="sparse_csr_spmv")
    dy = torch.from_numpy(g["y0"]).cuda()
    sparse_spmv(0, 0.75, -0.5, h.rows, h.cols, d.rowptr, d.col, d.val, dx, dy)
    torch.cuda.synchronize()
    assert cache_size() == 1
    assert_parity(h, g["x"], g["y0"], 0.75, -0.5, dy.cpu().numpy(), g["y"][1], what="sparse_spmv")
    with pytest.raises(SpmvB200Error):
        sparse_csr_spmv(1, 1.0, 1.0, desc, desc, dx, dy)  # operation_transpose is unsupported, as in the reference
    with pytest.raises(SpmvB200Error):
        sparse_csr_spmv(0, 1.0, 1.0, desc, desc, dx.cpu(), dy)  # host memory is refused: there is no CPU fallback
    cache_invalidate()
    assert cache_size() == 0


def test_plan_cache_detects_a_new_matrix_at_the_same_addresses():
    """A caching allocator hands the same addresses to the next matrix of the same shape: the stateless entry points
    must notice (fingerprint of the row pointers, re-read on the device on every call) and analyse again."""
    import torch
    cache_invalidate()
    m, n = 6000, 4000
    a = _ragged(21, m, n, [0, 1, 2, 3, 5, 9, 17, 40, 300, 2500])
    # a second matrix with the same number of rows AND non-zeros, other row lengths: reverse the row order
    lens = np.diff(a.rowptr)[::-1]
    rp = np.zeros(m + 1, np.int32)
    rp[1:] = np.cumsum(lens)
    rng = np.random.default_rng(22)
    b = synth.Csr(m, n, rp, rng.integers(0, n, a.nnz).astype(np.int32), rng.standard_normal(a.nnz))
    assert b.nnz == a.nnz and not np.array_equal(a.rowptr, b.rowptr)
    d = synth.to_device(a)
    desc = desc_of(d)
    x, y0 = synth.vector_numpy(n, 2), synth.vector_numpy(m, 3)
    dx = torch.from_numpy(x).cuda()
    before = cache_revalidations()
    for h in (a, b, b, a):
        d.rowptr.copy_(torch.from_numpy(h.rowptr))   # same device arrays, new contents
        d.col.copy_(torch.from_numpy(h.col))
        d.val.copy_(torch.from_numpy(h.val))
        dy = torch.from_numpy(y0).cuda()
        sparse_csr_spmv(0, 0.75, -0.5, desc, desc, dx, dy)
        torch.cuda.synchronize()
        assert_parity(h, x, y0, 0.75, -0.5, dy.cpu().numpy(), what="same address, new contents")
        dy = torch.from_numpy(y0).cuda()
        sparse_spmv(0, 1.0, 1.0, m, n, d.rowptr, d.col, d.val, dx, dy)  # nnz read on the device
        torch.cuda.synchronize()
        assert_parity(h, x, y0, 1.0, 1.0, dy.cpu().numpy(), what="same address, new contents (sparse_spmv)")
        assert cache_size() == 1
    assert cache_revalidations() - before == 2   # a -> b and b -> a; the repeated b is a plain hit
    cache_invalidate()


def test_host_buffer_path_on_a_device_resident_matrix():
    """spmv_b200_hostmat_create_device: the matrix already lives on the device (cli/utils.hpp:94-116 done by the caller),
    x / y travel per call; with beta == 0 and BETA0_SKIP_Y the y0 copy is skipped and NaN in h_y does not propagate."""
    h = synth.stencil3d_numpy(40)
    d = synth.to_device(h)
    x, y0 = synth.vector_numpy(h.cols, 2), synth.vector_numpy(h.rows, 3)
    hm = HostMatrix(h.rows, h.cols, d.rowptr, d.col, d.val)
    for a, b in AB[:3]:
        y = y0.copy()
        hm.spmv(a, b, x, y)
        assert_parity(h, x, y0, a, b, y, what="hostmat on device arrays")
    hm.destroy()
    hm = HostMatrix(h.rows, h.cols, d.rowptr, d.col, d.val, make_options(flags=FLAG_BETA0_SKIP_Y))
    y = np.full(h.rows, np.nan)
    hm.spmv(1.0, 0.0, x, y)
    assert_parity(h, x, np.zeros(h.rows), 1.0, 0.0, y, what="hostmat, y0 not sent")
    hm.destroy()
    with pytest.raises(SpmvB200Error):
        HostMatrix(h.rows, h.cols, h.rowptr, d.col, d.val)  # mixed host / device arrays


def test_host_buffer_path():
    g = load_golden("c4_rmat_s11")
    h = _golden_csr(g)
    hm = HostMatrix(h.rows, h.cols, h.rowptr, h.col, h.val)
    for (a, b), y_ref in zip(g["ab"], g["y"]):
        y = g["y0"].copy()
        hm.spmv(float(a), float(b), g["x"], y)
        assert_parity(h, g["x"], g["y0"], float(a), float(b), y, y_ref, what="hostmat")
    hm.destroy()
    y = g["y0"].copy()
    host_spmv(1.0, 1.0, h.rows, h.cols, h.rowptr, h.col, h.val, g["x"], y)
    assert_parity(h, g["x"], g["y0"], 1.0, 1.0, y, g["y"][0], what="host_spmv")


def test_host_buffer_path_pipelined_chunks():
    """Matrices without split rows and with enough tiles take the chunked, copy/compute-overlapped host path."""
    for name, h in (("stencil2d_300", synth.stencil2d_numpy(300)), ("stencil3d_40", synth.stencil3d_numpy(40)),
                    ("uniform", synth.uniform_numpy(4000, 6000, 32, seed=5))):
        x, y0 = synth.vector_numpy(h.cols, 2), synth.vector_numpy(h.rows, 3)
        hm = HostMatrix(h.rows, h.cols, h.rowptr, h.col, h.val)
        for a, b in AB[:3]:
            y = y0.copy()
            hm.spmv(a, b, x, y)
            assert_parity(h, x, y0, a, b, y, what=f"hostmat pipelined {name}")
            y2 = y0.copy()
            hm.spmv(a, b, x, y2)
            assert np.array_equal(y, y2)
        hm.destroy()


def test_host_buffer_path_copies_only_the_referenced_part_of_x():
    """A row shard references its own columns plus a halo: entries of h_x outside that range are neither copied nor
    read (they are poisoned with NaN here), in the pipelined and in the single-shot host path."""
    N = 40
    full_rows = N ** 3
    for lo, hi in ((full_rows // 3, 2 * full_rows // 3), (full_rows // 2, full_rows // 2 + 700)):
        h = synth.stencil3d_numpy(N, lo, hi)
        x, y0 = synth.vector_numpy(h.cols, 2), synth.vector_numpy(h.rows, 3)
        hm = HostMatrix(h.rows, h.cols, h.rowptr, h.col, h.val)
        c0, c1 = hm.x_range()
        assert (c0, c1) == (int(h.col.min()), int(h.col.max()) + 1) and c1 - c0 < h.cols
        xp = x.copy()
        xp[:c0] = np.nan
        xp[c1:] = np.nan
        y = y0.copy()
        hm.spmv(0.75, -0.5, xp, y)
        assert_parity(h, x, y0, 0.75, -0.5, y, what=f"hostmat shard rows {lo}:{hi}")
        hm.destroy()


def test_bad_options_are_rejected():
    d = synth.to_device(synth.stencil2d_numpy(8))
    for bad in (make_options(100), make_options(2048, 8, 4096), make_options(2048, 300, 128), make_options(2048, 8, 126)):
        with pytest.raises(SpmvB200Error):
            SpmvPlan(desc_of(d), bad)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json sizes: size-independent properties (the serial oracle would take too long for every test run)
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_c2_properties():
    """C2: 2D 5-point Laplacian on a 4096^2 grid (16.8M rows, 83.9M nnz)."""
    import torch
    N = 4096
    d = synth.stencil2d_device(N)
    p = SpmvPlan(desc_of(d))
    ones = torch.ones(d.cols, dtype=torch.float64, device="cuda")
    y = torch.zeros(d.rows, dtype=torch.float64, device="cuda")
    p.execute(1.0, 0.0, ones, y)
    # known answer: A*1 = 4 - (number of neighbours) -> 0 in the interior, 1 on edges, 2 in corners (exact in fp64)
    yy = y.view(N, N)
    assert float(yy[1:-1, 1:-1].abs().max()) == 0.0
    assert float(yy[0, 1:-1].min()) == 1.0 and float(yy[0, 1:-1].max()) == 1.0
    assert float(yy[0, 0]) == 2.0 and float(yy[-1, -1]) == 2.0
    assert float(y.sum()) == 4.0 * (N - 2) + 4 * 2.0
    # linearity + epilogue: A(2u + 3v) == 2Au + 3Av exactly is not guaranteed in fp64; check within the bound
    u, v = synth.vector_device(d.cols, 2), synth.vector_device(d.cols, 3)
    yu, yv, yw = (torch.zeros_like(y) for _ in range(3))
    p.execute(1.0, 0.0, u, yu)
    p.execute(1.0, 0.0, v, yv)
    p.execute(1.0, 0.0, 2.0 * u + 3.0 * v, yw)
    err = (yw - (2.0 * yu + 3.0 * yv)).abs().max()
    assert float(err) <= 64 * 8 * 2.2e-16 * 5
    # bitwise reproducibility and alpha/beta epilogue at full size
    y1 = synth.vector_device(d.rows, 4)
    y2 = y1.clone()
    p.execute(0.75, -0.5, u, y1)
    p.execute(0.75, -0.5, u, y2)
    assert torch.equal(y1, y2)
    assert float((y1 - (0.75 * yu - 0.5 * synth.vector_device(d.rows, 4))).abs().max()) <= 1e-14
    # a sampled block of rows against the oracle
    lo, hi = 4096 * 1000 + 17, 4096 * 1000 + 17 + 20000
    rp = d.rowptr[lo:hi + 1].cpu().numpy()
    sub = synth.Csr(hi - lo, d.cols, rp - rp[0], d.col[rp[0]:rp[-1]].cpu().numpy(), d.val[rp[0]:rp[-1]].cpu().numpy())
    y_ref = oracle.best_host_spmv(1.0, 0.0, sub.rowptr, sub.col, sub.val, u.cpu().numpy(), np.zeros(hi - lo))
    bound = oracle.port_row_bound(1.0, 0.0, sub.rowptr, sub.col, sub.val, u.cpu().numpy(), np.zeros(hi - lo))
    ok, worst, row = oracle.check_rows(yu[lo:hi].cpu().numpy(), y_ref, bound)
    assert ok, (worst, row)
    p.destroy()


def test_full_size_c4_rmat_row_sums():
    """C4: R-MAT scale 24 (16.8M rows, 268M nnz, longest row ~370k nnz): A*1 equals the per-row sums of the values,
    computed independently with a segmented reduction, within the fp64 bound; bitwise reproducible."""
    import torch
    d = synth.rmat_device(24, 16, seed=1)
    p = SpmvPlan(desc_of(d))
    info = p.info()
    assert info.nsplit_rows > 0 and info.bin_rows[3] > 0 and info.tiles_per_kind[2] > 0
    ones = torch.ones(d.cols, dtype=torch.float64, device="cuda")
    y = torch.zeros(d.rows, dtype=torch.float64, device="cuda")
    p.execute(1.0, 0.0, ones, y)
    y2 = torch.zeros_like(y)
    p.execute(1.0, 0.0, ones, y2)
    assert torch.equal(y, y2)
    lens = (d.rowptr[1:] - d.rowptr[:-1]).to(torch.int64)
    row_of = torch.repeat_interleave(torch.arange(d.rows, device="cuda"), lens)
    ref = torch.zeros_like(y).index_add_(0, row_of, d.val)
    absum = torch.zeros_like(y).index_add_(0, row_of, d.val.abs())
    # index_add_ is itself a floating-point reduction in another order: allow its error as well
    assert bool(((y - ref).abs() <= 2e-12 * absum + 1e-300).all())
    assert int((lens == 0).sum()) > 0 and float(y[lens == 0].abs().max()) == 0.0
    p.destroy()


def _check_sample(d, plan, x, alpha, beta, rows, what):
    """One SpMV at full size; the sampled rows against the reference's CPU SpMV (oracle/sampled.py); run twice."""
    import torch
    from oracle import sampled
    y0 = synth.vector_device(d.rows, 5)
    y = y0.clone()
    plan.execute(alpha, beta, x, y)
    y2 = y0.clone()
    plan.execute(alpha, beta, x, y2)
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int64), y2.view(torch.int64)), f"{what}: run-to-run results differ bitwise"
    res = sampled.check_sampled_rows(d.rowptr, d.col, d.val, x, y0, y, alpha, beta, rows)
    assert res["ok"], f"{what}: {res}"
    return res, y


def test_full_size_c3_sampled_rows_against_oracle():
    """C3: uniform random 10^7 x 10^7, 32 nnz per row (3.2e8 nnz): random x, 20 000 sampled rows + the ends."""
    from oracle import sampled
    d = synth.uniform_device(10_000_000, 10_000_000, 32, seed=1)
    p = SpmvPlan(desc_of(d))
    x = synth.vector_device(d.cols, 2)
    rows = sampled.sample_rows(d.rows, 20000, seed=11)
    for a, b in ((0.75, -0.5), (1.0, 1.0)):
        res, _ = _check_sample(d, p, x, a, b, rows, f"C3 a={a} b={b}")
        assert res["rows_checked"] >= 20000 and res["nnz_checked"] == 32 * res["rows_checked"]
    p.destroy()


def test_full_size_c4_sampled_rows_against_oracle():
    """C4: R-MAT scale 24 (2^28 nnz): random x; the sample holds the longest row, >= 100 rows that are split across
    row blocks (their partial sums are combined by the fix-up pass) and 20 000 random rows."""
    import torch
    from oracle import sampled
    d = synth.rmat_device(24, 16, seed=1)
    p = SpmvPlan(desc_of(d))
    info = p.info()
    assert info.direct == 1 and info.nsplit_rows >= 100
    split = p.export("split_rows")[:info.nsplit_rows]
    lens = (d.rowptr[1:] - d.rowptr[:-1])
    longest = int(torch.argmax(lens).item())
    assert longest in set(split.tolist())
    pick = np.unique(np.concatenate([split[:: max(1, split.size // 400)], [longest]]))
    rows = sampled.sample_rows(d.rows, 20000, seed=12, must_include=pick.tolist())
    x = synth.vector_device(d.cols, 2)
    res, y = _check_sample(d, p, x, 0.75, -0.5, rows, "C4")
    assert res["longest_row_checked"] == int(lens[longest].item()) and res["longest_row_checked"] > 100000
    # rows without elements: y = beta * y0 exactly
    y0 = synth.vector_device(d.rows, 5)
    empty = lens == 0
    assert int(empty.sum()) > 0 and torch.equal(y[empty], -0.5 * y0[empty])
    p.destroy()
    # the tiled (shared-memory) form on the same matrix
    p = SpmvPlan(desc_of(d), make_options(flags=FLAG_NO_DIRECT))
    assert p.info().direct == 0
    _check_sample(d, p, x, 1.0, 1.0, rows, "C4 tiled")
    p.destroy()


def test_full_size_c5_known_answer_and_sampled_rows():
    """C5: 27-point averaging stencil on 384^3 (56.6M rows, 1.52e9 nnz = 71 % of int32): A*1 = 1 on interior rows
    within the bound and (neighbours)/27 on the boundary; random x on 20 000 sampled rows and around z-plane
    boundaries and the very end of the arrays (where 32-bit byte offsets would have wrapped)."""
    import torch
    from oracle import sampled
    N = 384
    d = synth.stencil3d_device(N)
    assert d.nnz == (3 * N - 2) ** 3
    p = SpmvPlan(desc_of(d), make_options(flags=FLAG_BETA0_SKIP_Y))
    ones = torch.ones(d.cols, dtype=torch.float64, device="cuda")
    y = torch.full((d.rows,), float("nan"), dtype=torch.float64, device="cuda")
    p.execute(1.0, 0.0, ones, y)
    yy = y.view(N, N, N)
    assert float((yy[1:-1, 1:-1, 1:-1] - 1.0).abs().max()) <= 27 * 2.3e-16
    assert abs(float(yy[0, 0, 0]) - 8.0 / 27.0) <= 1e-15 and abs(float(yy[-1, -1, -1]) - 8.0 / 27.0) <= 1e-15
    assert abs(float(yy[0, 5, 5]) - 18.0 / 27.0) <= 1e-15 and abs(float(yy[-1, 0, 7]) - 12.0 / 27.0) <= 1e-15
    # the sum of all row sums = nnz / 27 (sum of pairwise sums on the device: allow its rounding)
    assert abs(float(y.sum()) - d.nnz / 27.0) <= 1e-9 * d.nnz / 27.0
    del ones, yy
    plane = N * N
    edge = [r for z in (1, 191, 192, 383) for r in range(z * plane - 40, z * plane + 40)]
    rows = sampled.sample_rows(d.rows, 20000, seed=13, must_include=edge)
    x = synth.vector_device(d.cols, 2)
    res, _ = _check_sample(d, p, x, 1.0, 0.0, rows, "C5")
    assert res["rows_checked"] >= 20000
    p.destroy()
    p = SpmvPlan(desc_of(d))      # default semantics (y read even when beta == 0), general alpha / beta
    _check_sample(d, p, x, 0.75, -0.5, rows, "C5 a=0.75 b=-0.5")
    p.destroy()


def test_plans_with_different_tile_sizes_coexist():
    """The shared-memory limit is a property of the kernel function: creating a plan with small tiles must not break a
    live plan with large tiles (regression: 'invalid argument' at launch)."""
    import torch
    h = synth.stencil3d_numpy(24)
    d = synth.to_device(h)
    x, y0 = synth.vector_numpy(h.cols, 2), synth.vector_numpy(h.rows, 3)
    big = SpmvPlan(desc_of(d), make_options(8192))
    small = SpmvPlan(desc_of(d), make_options(512))
    dx = torch.from_numpy(x).cuda()
    for plan in (big, small, big):
        dy = torch.from_numpy(y0).cuda()
        plan.execute(1.0, 1.0, dx, dy)
        torch.cuda.synchronize()
        assert_parity(h, x, y0, 1.0, 1.0, dy.cpu().numpy(), what="coexisting plans")
    big.destroy()
    small.destroy()
