"""Generates the golden vectors under tests/golden/ by running THE REFERENCE ITSELF: host_spmv and verify_y from
/root/reference/cli/verification.cpp and csr_adaptive_plus_analyze_imp from src/acc/hip-csr-adaptive-plus, compiled
in place into oracle/_ref/libref_oracle.so (`make -C oracle ref`). Run in the build container only (the reference is
not present on the GPU box); the outputs are committed.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from spmv_acc_b200 import formats, synth  # noqa: E402

OUT = Path(__file__).resolve().parent
AB = [(1.0, 1.0), (0.75, -0.5), (1.0, 0.0), (0.0, 2.0), (-1.25, 1e-3)]


def case(name, csr, x, y0):
    ys = [oracle.ref_host_spmv(a, b, csr.rowptr, csr.col, csr.val, x, y0) for a, b in AB]
    yax = oracle.ref_host_spmv_ax(csr.rowptr, csr.col, csr.val, x)
    np.savez_compressed(OUT / f"{name}.npz", rows=csr.rows, cols=csr.cols, rowptr=csr.rowptr, col=csr.col,
                        val=csr.val, x=x, y0=y0, ab=np.array(AB), y=np.stack(ys), y_ax=yax)
    print(name, csr.rows, csr.cols, csr.nnz)


def main():
    assert oracle.have_ref(), "build oracle/_ref first: make -C oracle ref"
    # C1: rajat03-shaped stand-in, x / y0 exactly as the reference CLI would draw them (cli/utils.hpp:65-85:
    # unseeded rand(): hX[cols], temphY[rows], ... in that order; for .csr input x comes from the file instead)
    c1 = synth.circuit_numpy()
    oracle.ref_generate_vector(0, seed=1)            # srand(1) == never seeded
    _hx = oracle.ref_generate_vector(c1.cols, seed=None)
    y0 = oracle.ref_generate_vector(c1.rows, seed=None)
    x = synth.vector_numpy(c1.cols, 2)              # the "file" x
    formats.write_csr_text(OUT / "rajat03_standin.csr", c1, x, header="rajat03-shaped synthetic stand-in 7602x7602")
    case("c1_circuit", c1, x, y0)
    # small members of the other BASELINE families
    for name, csr in (("c2_stencil2d_48", synth.stencil2d_numpy(48)), ("c5_stencil3d_10", synth.stencil3d_numpy(10)),
                      ("c3_uniform_600x700_32", synth.uniform_numpy(600, 700, 32, seed=1)),
                      ("c4_rmat_s11", synth.rmat_numpy(11, 16, seed=1))):
        case(name, csr, synth.vector_numpy(csr.cols, 2), synth.vector_numpy(csr.rows, 3))
    # edge cases: empty rows, one giant row, a single row, no non-zeros
    rng = np.random.default_rng(7)
    lens = np.array([0, 3, 0, 0, 5000, 1, 0, 17, 300, 0], dtype=np.int64)
    rp = np.zeros(lens.size + 1, np.int32); rp[1:] = np.cumsum(lens)
    n = 977
    ragged = synth.Csr(lens.size, n, rp, rng.integers(0, n, int(rp[-1])).astype(np.int32),
                       rng.standard_normal(int(rp[-1])))
    case("edge_ragged", ragged, rng.standard_normal(n), rng.standard_normal(lens.size))
    one = synth.Csr(1, 5, np.array([0, 3], np.int32), np.array([4, 0, 2], np.int32), np.array([1.5, -2.0, 0.25]))
    case("edge_single_row", one, np.arange(1.0, 6.0), np.array([10.0]))
    empty = synth.Csr(6, 4, np.zeros(7, np.int32), np.zeros(0, np.int32), np.zeros(0))
    case("edge_no_nnz", empty, np.ones(4), np.arange(6.0))
    # verify_y known answers (cli/verification.cpp:15-38)
    hy = np.array([1.0, 0.0, 1e-13, -2.0, 5.0, 0.0])
    dy = np.array([1.0 + 5e-8, 2e-14, 1e-13, -2.0 - 3e-7, 5.0, 0.5e-14])
    r = oracle.ref_verify_y(dy, hy)
    np.savez(OUT / "verify_y.npz", hy=hy, dy=dy, max_error=r["max_error"], first_failed_at=r["first_failed_at"],
             failed_count=r["failed_count"])
    # the reference's own nnz-balanced row blocks for the C1 stand-in (structure cross-check, not our layout)
    bp, first = oracle.ref_adaptive_plus_analyze(c1.rowptr, 2048, 8)
    np.savez_compressed(OUT / "c1_adaptive_plus_blocks.npz", break_points=bp, first_block_of_row=first)
    # the reference's vector generator (first 64 draws of the unseeded stream)
    np.save(OUT / "rand_vector_64.npy", oracle.ref_generate_vector(64, seed=1))


if __name__ == "__main__":
    main()
