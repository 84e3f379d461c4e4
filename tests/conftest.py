import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    return np.load(GOLDEN / f"{name}.npz")


GOLDEN_CASES = ["c1_circuit", "c2_stencil2d_48", "c5_stencil3d_10", "c3_uniform_600x700_32", "c4_rmat_s11",
                "edge_ragged", "edge_single_row", "edge_no_nnz"]
