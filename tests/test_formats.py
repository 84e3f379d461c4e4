"""CPU tests for the on-disk formats (next-row f.1): our readers/writers against the reference's own readers."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from spmv_acc_b200 import formats, synth

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def _small():
    csr = synth.uniform_numpy(40, 50, 7, seed=3)
    return csr, synth.vector_numpy(50, 4)


def test_csr_text_round_trip(tmp_path):
    csr, x = _small()
    p = tmp_path / "a.csr"
    formats.write_csr_text(p, csr, x)
    got, gx = formats.read_csr_text(p)
    assert (got.rows, got.cols) == (40, 50)
    assert np.array_equal(got.rowptr, csr.rowptr) and np.array_equal(got.col, csr.col)
    assert np.array_equal(got.val, csr.val) and np.array_equal(gx, x)


def test_golden_c1_file_matches_fixture():
    got, x = formats.read_csr_text(GOLDEN / "rajat03_standin.csr")
    g = np.load(GOLDEN / "c1_circuit.npz")
    assert (got.rows, got.cols, got.nnz) == (7602, 7602, 32653)
    assert np.array_equal(got.rowptr, g["rowptr"]) and np.array_equal(got.col, g["col"])
    assert np.array_equal(got.val, g["val"]) and np.array_equal(x, g["x"])


@pytest.mark.parametrize("vt", [formats.TP_FLOAT, formats.TP_BOOL])
def test_bin2_round_trip(tmp_path, vt):
    csr, _ = _small()
    p = tmp_path / "a.bin2"
    formats.write_bin2(p, csr, vt)
    got = formats.read_bin2(p)
    assert np.array_equal(got.rowptr, csr.rowptr) and np.array_equal(got.col, csr.col)
    assert np.array_equal(got.val, csr.val if vt == formats.TP_FLOAT else np.ones(csr.nnz))


def test_bin2_rejects_bad_magic(tmp_path):
    p = tmp_path / "bad.bin2"
    p.write_bytes(b"\0" * 64)
    with pytest.raises(ValueError):
        formats.read_bin2(p)


def test_mtx_round_trip_general_and_symmetric(tmp_path):
    csr, _ = _small()
    p = tmp_path / "a.mtx"
    formats.write_mtx(p, csr)
    got = formats.read_mtx(p)
    assert np.array_equal(got.rowptr, csr.rowptr) and np.array_equal(got.col, csr.col)
    assert np.array_equal(got.val, csr.val)
    sym = synth.stencil2d_numpy(6)
    formats.write_mtx(p, sym, symmetric_lower_only=True)
    got = formats.read_mtx(p)
    assert np.array_equal(got.rowptr, sym.rowptr) and np.array_equal(got.col, sym.col)
    assert np.array_equal(got.val, sym.val)


@needs_ref
def test_readers_agree_with_reference_readers(tmp_path):
    csr, x = _small()
    formats.write_csr_text(tmp_path / "a.csr", csr, x)
    formats.write_bin2(tmp_path / "a.bin2", csr)
    formats.write_mtx(tmp_path / "a.mtx", csr)
    for fmt in ("csr", "bin2", "mtx"):
        rp, col, val, rx, n = oracle.ref_read(tmp_path / f"a.{fmt}", fmt)
        ours, ox = formats.load(tmp_path / f"a.{fmt}", fmt)
        assert np.array_equal(rp, ours.rowptr) and np.array_equal(col, ours.col) and np.array_equal(val, ours.val)
        assert n == ours.cols
        if fmt == "csr":
            assert np.array_equal(rx, ox)


@needs_ref
def test_reference_reader_reads_the_committed_c1_file():
    rp, col, val, x, n = oracle.ref_read(GOLDEN / "rajat03_standin.csr", "csr")
    g = np.load(GOLDEN / "c1_circuit.npz")
    assert np.array_equal(rp, g["rowptr"]) and np.array_equal(col, g["col"]) and np.array_equal(val, g["val"])
    assert np.array_equal(x, g["x"]) and n == 7602


def test_csr_tool_output_follows_the_reference_tool():
    """Hand-checked against tools/main.cpp:125-181 on rows of length 2, 0, 3, 0, 0, 1, 4."""
    from spmv_acc_b200 import csr_tool
    rp = np.array([0, 2, 2, 5, 5, 5, 6, 10], np.int32)
    assert csr_tool.dist_lines(rp) == ["0 = 3", "1 = 1", "2 = 1", "3 = 1", "4 = 1"]
    # 3 parts -> ceil(7/3) = 3 rows per part: rows 0-2 (5 nnz), 3-5 (1 nnz), last part row 6 (4 nnz)
    assert csr_tool.part_nnz_lines(rp, 3) == ["[part ID] [part nnz] [avg-nnz/row]", "0 5 1.66667", "1 1 0.333333", "2 4 4"]
    # parts == 0: one part per row
    assert csr_tool.part_nnz_lines(rp, 0)[1:4] == ["0 2 2", "1 0 0", "2 3 3"]
    assert csr_tool.main(["dist", str(GOLDEN / "rajat03_standin.csr")]) == 0


def test_mtx_pattern_complex_and_comment_lines_against_reference_reader(tmp_path):
    """Field variants and the two parse paths (vectorised; line by line when a comment sits between the entries) give
    what the reference's MatrixMarket reader gives (cli/matrix_market_reader.hpp:50-303)."""
    cases = {
        "pattern_sym.mtx": "%%MatrixMarket matrix coordinate pattern symmetric\n% c\n4 4 4\n1 1\n3 1\n4 2\n4 4\n",
        "complex_gen.mtx": "%%MatrixMarket matrix coordinate complex general\n3 4 3\n1 4 1.5 -2\n3 1 0.25 7\n2 2 -1 0\n",
        "integer_herm.mtx": "%%MatrixMarket matrix coordinate integer Hermitian\n3 3 3\n2 1 5\n3 3 -4\n3 2 9\n",
        "real_comment_inside.mtx": "%%MatrixMarket matrix coordinate real general\n2 3 3\n1 1 1.0\n% a comment\n2 3 2.5\n1 2 -1\n",
    }
    for name, text in cases.items():
        path = tmp_path / name
        path.write_text(text)
        ours = formats.read_mtx(path)
        if oracle.have_ref():
            rp, col, val, _, cols = oracle.ref_read(str(path), "mtx")
            assert (ours.rows, ours.cols) == (rp.size - 1, cols), name
            assert np.array_equal(ours.rowptr, rp) and np.array_equal(ours.col, col), name
            assert np.array_equal(ours.val, val), name
    pat = formats.read_mtx(tmp_path / "pattern_sym.mtx")
    assert pat.nnz == 6 and np.all(pat.val == 1.0)                       # 2 diagonal + 2 mirrored pairs
    assert formats.read_mtx(tmp_path / "complex_gen.mtx").val.tolist() == [1.5, -1.0, 0.25]   # real parts, (row, col) order
    with pytest.raises(ValueError):
        (tmp_path / "bad.mtx").write_text("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n")
        formats.read_mtx(tmp_path / "bad.mtx")
