"""csr-tool diagnostics from the device: the `nnz` / `dist` tables computed on the GPU equal the
host tables line for line (which follow the reference's tools/main.cpp:117-182), and `plan` reports what the engine's
device analysis pass found, consistent with the row pointers."""
import numpy as np
import pytest

from spmv_acc_b200 import csr_tool, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,make", [("stencil2d_64", lambda: synth.stencil2d_numpy(64)),
                                       ("rmat_12", lambda: synth.rmat_numpy(12, 16, seed=1)),
                                       ("uniform_800", lambda: synth.uniform_numpy(800, 1200, 32, seed=1))])
def test_device_tables_equal_host_tables(name, make):
    h = make()
    assert csr_tool.dist_lines_device(h.rowptr) == csr_tool.dist_lines(h.rowptr)
    for parts in (0, 1, 7, 64):
        assert csr_tool.part_nnz_lines_device(h.rowptr, parts) == csr_tool.part_nnz_lines(h.rowptr, parts), (name, parts)


def test_plan_mode_reports_the_device_analysis(tmp_path, capsys):
    from spmv_acc_b200 import formats
    h = synth.rmat_numpy(12, 16, seed=1)
    lines = csr_tool.plan_lines(h)
    lens = np.diff(h.rowptr.astype(np.int64))
    assert lines[0].startswith(f"rows {h.rows} cols {h.cols} nnz {h.nnz} ")
    short = [ln for ln in lines if ln.startswith("short (<= 8):")][0].split(":")[1].split()
    assert int(short[0]) == int((lens <= 8).sum()) and int(short[1]) == int(lens[lens <= 8].sum())
    total_rows = sum(int(ln.split(":")[1].split()[0]) for ln in lines[2:6])
    total_nnz = sum(int(ln.split(":")[1].split()[1]) for ln in lines[2:6])
    assert total_rows == h.rows and total_nnz == h.nnz
    assert any(ln.startswith("form: ") for ln in lines)
    # through the command line, on a file in the reference's bin2 format
    path = tmp_path / "m.bin2"
    formats.write_bin2(path, h)
    assert csr_tool.main(["plan", str(path), "-f", "bin2"]) == 0
    out = capsys.readouterr().out
    assert "row blocks of" in out and "form:" in out
    assert csr_tool.main(["dist", str(path), "-f", "bin2", "--device"]) == 0
    assert capsys.readouterr().out.strip().splitlines() == csr_tool.dist_lines(h.rowptr)
