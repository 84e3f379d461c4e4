"""CPU tests: the plain-C oracle against the reference's golden vectors (produced by the reference itself, see
tests/golden/make_golden.py) and, where /root/reference was compiled here, against the reference bit for bit."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN_CASES, load_golden
from spmv_acc_b200 import synth

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_matches_golden_bitwise(name):
    g = load_golden(name)
    for (a, b), y in zip(g["ab"], g["y"]):
        got = oracle.port_host_spmv(a, b, g["rowptr"], g["col"], g["val"], g["x"], g["y0"], n=int(g["cols"]))
        assert np.array_equal(got, y), f"{name} alpha={a} beta={b}"
    assert np.array_equal(oracle.port_host_spmv_ax(g["rowptr"], g["col"], g["val"], g["x"]), g["y_ax"])


def test_verify_y_golden():
    g = np.load(load_golden.__globals__["GOLDEN"] / "verify_y.npz")
    r = oracle.port_verify_y(g["dy"], g["hy"])
    assert r["max_error"] == float(g["max_error"])
    assert r["first_failed_at"] == int(g["first_failed_at"])
    assert r["failed_count"] == int(g["failed_count"])
    # known answers by hand: |hy|<=1e-12 -> abs 1e-14; else rel 1e-7 (cli/verification.cpp:24-25)
    assert r["failed_count"] == 2 and r["first_failed_at"] == 1


def test_rand_vector_golden():
    g = np.load(load_golden.__globals__["GOLDEN"] / "rand_vector_64.npy")
    v = oracle.port_generate_vector(64, seed=1)
    assert np.array_equal(v, g)
    assert v.min() >= -1.0 and v.max() <= 1.0 - 2.0 * 2 / 101  # 100-point lattice in [-1, 0.96]


@needs_ref
def test_port_equals_reference_on_random_matrices():
    rng = np.random.default_rng(11)
    for trial in range(40):
        m, n = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        lens = rng.integers(0, 40, m)
        rp = np.zeros(m + 1, np.int32)
        rp[1:] = np.cumsum(lens)
        col = rng.integers(0, n, int(rp[-1])).astype(np.int32)
        val = rng.standard_normal(int(rp[-1]))
        x, y0 = rng.standard_normal(n), rng.standard_normal(m)
        a, b = float(rng.standard_normal()), float(rng.standard_normal())
        assert np.array_equal(oracle.port_host_spmv(a, b, rp, col, val, x, y0),
                              oracle.ref_host_spmv(a, b, rp, col, val, x, y0))
        dy = oracle.port_host_spmv(a, b, rp, col, val, x, y0) * (1 + 1e-7 * (rng.random(m) < 0.1))
        hy = oracle.ref_host_spmv(a, b, rp, col, val, x, y0)
        assert oracle.port_verify_y(dy, hy) == oracle.ref_verify_y(dy, hy)


def test_beta_zero_propagates_nan_like_reference():
    # cli/verification.cpp:64 computes alpha*y0 + beta*y[i] unconditionally: 0 * NaN = NaN
    rp = np.array([0, 1, 2], np.int32)
    y = oracle.port_host_spmv(1.0, 0.0, rp, np.array([0, 1], np.int32), np.array([2.0, 3.0]), np.array([1.0, 1.0]),
                              np.array([np.nan, 5.0]))
    assert np.isnan(y[0]) and y[1] == 3.0


def test_merge_path_partition_port_semantics():
    rp = np.array([0, 0, 3, 3, 3, 10, 10, 12], np.int32)
    # S[t] = largest r with rowptr[r] <= t*ITEMS  (merge_path_partition.h:14, merge_path_utils.h:32-44)
    S = oracle.port_merge_path_partition(rp, 5, 3)
    assert S.tolist() == [1, 4, 4, 4, 7]
    bp = oracle.port_flat_break_points_v2(rp, 4)
    # element 0 -> row 1, element 4 -> row 4, element 8 -> row 4 (flat_imp.inl:134-152)
    assert bp[:3].tolist() == [1, 4, 4]


def _random_rowptr(rng, m, choices):
    lens = rng.choice(choices, size=m)
    rp = np.zeros(m + 1, np.int32)
    rp[1:] = np.cumsum(lens)
    return rp


@pytest.mark.parametrize("T,S,L", [(256, 8, 16), (256, 4, 64), (512, 8, 128), (2048, 8, 128)])
def test_analysis_port_invariants_and_tiled_sum(T, S, L):
    rng = np.random.default_rng(T + S + L)
    for trial in range(60):
        m, n = int(rng.integers(1, 80)), int(rng.integers(1, 60))
        rp = _random_rowptr(rng, m, [0, 0, 1, 2, 3, 5, 9, 17, 40, 300, 700, 2500])
        nnz = int(rp[-1])
        col = rng.integers(0, n, nnz).astype(np.int32)
        val = rng.standard_normal(nnz)
        x, y0 = rng.standard_normal(n), rng.standard_normal(m)
        ana = oracle.port_analysis(rp, T, S, L)
        nt = ana["ntiles"]
        assert nt == max(1, -(-(nnz + m) // T))                       # tiles balance nnz + rows
        tr, te = ana["tile_row"], ana["tile_elem"]
        assert tr[0] == 0 and tr[-1] == m and te[0] == 0 and te[-1] == nnz
        assert np.all(np.diff(tr) >= 0) and np.all(np.diff(te) >= 0)
        assert np.all(np.diff(te) <= T + L - 1)                       # nnz balance: the shared-memory tile bound
        assert np.all(np.diff(tr) <= T)                               # row balance, however many rows are empty
        assert ana["bin_rows"].sum() == m and ana["bin_nnz"].sum() == nnz
        # every row is summed exactly once and the result matches the serial oracle within the fp64 bound
        yt = oracle.port_tiled_spmv(1.5, 0.25, rp, col, val, x, y0, ana, T, L)
        yr = oracle.port_host_spmv(1.5, 0.25, rp, col, val, x, y0)
        ok, worst, row = oracle.check_rows(yt, yr, oracle.port_row_bound(1.5, 0.25, rp, col, val, x, y0))
        assert ok, (trial, worst, row)
        if ana["nsplit"] == 0:
            assert np.array_equal(yt, yr)


def test_shard_bounds_port():
    rp = np.arange(0, 1001, 10, dtype=np.int32)  # 100 rows x 10 nnz
    assert oracle.port_shard_bounds(rp, 4).tolist() == [0, 25, 50, 75, 100]
    rp2 = np.array([0, 1000, 1000, 1001, 1002], np.int32)
    assert oracle.port_shard_bounds(rp2, 2).tolist() == [0, 1, 4]
    assert oracle.port_shard_bounds(np.zeros(1, np.int32), 3).tolist() == [0, 0, 0, 0]


@needs_ref
def test_reference_adaptive_plus_blocks_are_nnz_balanced_like_ours():
    """Structure cross-check against the reference's own row-block analysis (csr_adaptive_plus_analyze.cpp:12-98):
    both cut the C1 stand-in into blocks of about 2048 nnz that start on row boundaries."""
    c1 = synth.circuit_numpy()
    bp, _ = oracle.ref_adaptive_plus_analyze(c1.rowptr, 2048, 8)
    g = np.load(load_golden.__globals__["GOLDEN"] / "c1_adaptive_plus_blocks.npz")
    assert np.array_equal(bp, g["break_points"])
    ours = oracle.port_analysis(c1.rowptr, 2048, 8, 128)
    ref_nnz = np.diff(c1.rowptr[bp])
    our_nnz = np.diff(ours["tile_elem"])
    assert bp[0] == 0 and bp[-1] == c1.rows and ours["tile_row"][-1] == c1.rows
    assert our_nnz.max() <= 2048 + 127 and abs(len(our_nnz) - len(ref_nnz)) <= len(ref_nnz)


def test_port_direct_arrays_small_known_answer():
    """Hand-checked: rows of length 2, 0, 3, 0, 0, 1 -> starts at elements 0, 2, 5; non-empty rows 0, 2, 5."""
    rowptr = np.array([0, 2, 2, 5, 5, 5, 6], np.int32)
    out = oracle.port_direct_arrays(rowptr, np.array([0, 1, 3, 6]))
    assert out["nz_rows"].tolist() == [0, 2, 5]
    assert out["row_start_bits"].tolist() == [(1 << 0) | (1 << 2) | (1 << 5)]
    assert out["tile_nzbase"].tolist() == [0, 1, 2]   # non-empty rows in front of rows 0, 1, 3
    # starts beyond bit 31 land in the next word
    rowptr = np.array([0, 40, 40, 70], np.int32)
    out = oracle.port_direct_arrays(rowptr, np.array([0, 3]))
    assert out["row_start_bits"].tolist() == [1, 1 << 8, 0]


def test_port_adaptive_choice_on_the_baseline_shapes():
    """The reference's 4-sample selector (src/acc/hip-adaptive/adaptive.cpp:16-67) on the row pointers of the
    BASELINE.json configs; SURVEY.md §8a lists the same picks."""
    g = load_golden("c1_circuit")
    assert oracle.port_adaptive_choice(g["rowptr"]) == "adaptive line"            # 32653 / 7602 = 4 (integer) <= 4
    N = 4096                                                                       # C2: 5-point stencil, avg 4.999 -> 4
    counts = np.full((N, N), 5, np.int32)
    counts[0, :] -= 1; counts[-1, :] -= 1; counts[:, 0] -= 1; counts[:, -1] -= 1
    rp = np.zeros(N * N + 1, np.int32)
    np.cumsum(counts.ravel(), out=rp[1:])
    assert int(rp[-1]) == 83869696 and oracle.port_adaptive_choice(rp) == "adaptive line"
    rp = (np.arange(10_000_001, dtype=np.int64) * 32).astype(np.int32)            # C3: avg 32, nnz 3.2e8 > 2^23
    assert oracle.port_adaptive_choice(rp) == "adaptive flat"
    rp = np.array([0, 0, 0, 1000, 9000], np.int32)                                 # halves 0 : 9000 -> two data blocks
    assert oracle.port_adaptive_choice(rp) == "vector-row, two data blocks"
    rp = (np.arange(1001, dtype=np.int32) * 9)                                     # small, avg 9 -> line-enhance (adaptive)
    assert oracle.port_adaptive_choice(rp) == "adaptive line-enhance"


def _decode_staged_columns(col, tile_elem, ref):
    """Inverse of the staged-x encoding: the column an lcol entry stands for, from the tile's segment table."""
    xd = ref["xdesc"].reshape(-1, 32)
    out = np.empty_like(col)
    for t in range(xd.shape[0]):
        e0, e1 = int(tile_elem[t]), int(tile_elem[t + 1])
        nseg, nlines = int(xd[t, 0]), int(xd[t, 1])
        line = xd[t, 2:2 + nseg].astype(np.int64)
        off = xd[t, 18:26].copy().view(np.uint16)[:nseg].astype(np.int64)
        assert nseg >= 1 and np.all(np.diff(line) > 0) and off[0] == 0 and np.all(np.diff(off) > 0) and off[-1] < nlines
        lc = ref["lcol"][e0:e1].astype(np.int64)
        rank = lc >> 4
        assert rank.max(initial=0) < nlines
        seg = np.searchsorted(off, rank, side="right") - 1
        out[e0:e1] = ((line[seg] + (rank - off[seg])) << 4) | (lc & 15)
    return out


def test_staged_x_port_encoding_decodes_back_to_the_columns():
    """oracle.port_xstage (the specification the GPU analysis is compared with bit for bit): for stencils, banded and
    mesh-like matrices -- including row blocks whose runs of x lines had to be merged across gaps -- every 16-bit local
    index, decoded through its row block's segment table, is the original column; segment tables are sorted, ranks fit
    the staged line count, and merged tables stage at least the referenced lines."""
    rng = np.random.default_rng(5)
    cases = [synth.stencil2d_numpy(40), synth.stencil3d_numpy(12)]
    for m, strides, reach in ((20000, (1, 200), 10), (30000, (1, 97, 9409), 6), (8000, (1, 64), 3)):
        offs = sorted({0} | {sg * k * st for st in strides for k in range(1, reach + 1) for sg in (1, -1)})
        r = np.arange(m, dtype=np.int64)[:, None] + np.asarray(offs, dtype=np.int64)[None, :]
        ok = (r >= 0) & (r < m)
        rp = np.zeros(m + 1, dtype=np.int32)
        rp[1:] = np.cumsum(ok.sum(1))
        cases.append(synth.Csr(m, m, rp, r[ok].astype(np.int32), rng.standard_normal(int(ok.sum()))))
    merged = 0
    for h in cases:
        for T in (512, 2048):
            te = [0]
            while te[-1] < h.nnz:  # row-aligned blocks of about T non-zeros
                r = int(np.searchsorted(h.rowptr, min(h.nnz, te[-1] + T), side="left"))
                nxt = int(h.rowptr[min(r, h.rows)])
                te.append(nxt if nxt > te[-1] else h.nnz)
            te = np.asarray(te, dtype=np.int32)
            ref = oracle.port_xstage(h.col, te)
            xd = ref["xdesc"].reshape(-1, 32)
            good = xd[:, 0] > 0
            assert ref["failed"] == int((~good).sum())
            if ref["failed"]:
                continue
            assert np.array_equal(_decode_staged_columns(h.col, te, ref), h.col)
            assert ref["max_lines"] == int(xd[:, 1].max()) <= 256 and int(xd[:, 0].max()) <= 16
            for t in range(0, xd.shape[0], max(1, xd.shape[0] // 40)):
                referenced = np.unique(h.col[te[t]:te[t + 1]] >> 4).size
                assert xd[t, 1] >= referenced
                merged += int(xd[t, 1] > referenced)
    assert merged > 0, "no row block needed the merge path"


@pytest.mark.skipif(not oracle.have_ref_selector(), reason="oracle/_ref/libref_selector.so not built (needs /root/reference)")
def test_selector_restatement_equals_the_reference_selector_compiled_in_place():
    """port_adaptive_choice (quoted by the selector study) against the reference's own src/acc/hip-adaptive/adaptive.cpp
    compiled unmodified (its five launchers replaced by recorders): row-pointer arrays that reach every branch -- two
    unbalanced halves in both directions, short rows, small and large matrices, empty leading rows -- and random ones.
    Only rowptr[m/4], [m/2], [3m/4], [m] matter to the selector, so large nnz are produced with scaled row pointers."""
    rng = np.random.default_rng(11)
    cases = []

    def rp_from_lens(lens):
        return np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.int64))]).astype(np.int32)

    cases.append(rp_from_lens(np.full(1000, 3)))                               # avg <= 4 -> line
    cases.append(rp_from_lens(np.full(1000, 5)))                               # small -> adaptive line-enhance
    cases.append(rp_from_lens(np.full(400_000, 32)))                           # 12.8 M nnz -> flat
    cases.append(rp_from_lens(np.full(300_000, 40)))                           # 12.0 M nnz -> line-enhance (adaptive)
    cases.append(rp_from_lens(np.r_[np.full(500, 1), np.full(500, 40)]))       # second half 40x heavier
    cases.append(rp_from_lens(np.r_[np.full(500, 40), np.full(500, 1)]))       # first half heavier
    cases.append(rp_from_lens(np.r_[np.full(500, 10), np.full(500, 39)]))      # ratio 3.9 by integer division -> no split
    cases.append(rp_from_lens(np.r_[np.full(500, 10), np.full(500, 40)]))      # ratio 4 -> split
    cases.append(rp_from_lens(np.full(2_200_000, 4)))                          # avg 4 exactly, 8.8 M nnz
    cases.append(rp_from_lens(np.full(1_700_000, 5)))                          # 8.5 M nnz: between 2^23 and 0xC00000
    for _ in range(60):
        m = int(rng.integers(8, 5000))
        scale = int(rng.choice([1, 7, 300, 3000]))
        lens = rng.integers(1, 12, m) * scale
        if rng.random() < 0.3:
            lens[: m // 2] = rng.integers(1, 3, m // 2)
        cases.append(rp_from_lens(lens))
    seen = set()
    for rp in cases:
        if int(rp[(rp.size - 1) // 2]) == 0 or int(rp[-1]) == int(rp[(rp.size - 1) // 2]):
            continue  # the reference divides by the non-zeros of a half: an empty half is undefined behaviour there
        a, b = oracle.ref_adaptive_choice(rp), oracle.port_adaptive_choice(rp)
        assert a == b, (a, b, rp.size - 1, int(rp[-1]))
        seen.add(a)
    # the selector's last arm (plain line-enhance) needs nnz > 0xC00000 and nnz <= 2^23 at once: dead code in the reference
    assert seen == set(oracle.ADAPTIVE_CHOICES) - {"line-enhance"}, seen
