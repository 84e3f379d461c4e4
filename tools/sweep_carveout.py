"""Shared-memory carve-out sensitivity of the MIXED / MEDIUM kernels (development tool): the unified L1/shared array
is 256 KB per SM; what is not carved out for shared memory is L1, and the number of x gathers in flight scales with it.

    python tools/sweep_carveout.py c4 --tiles 1024,2048 --carve -1,30,45,57,72,86,100 [--flags F]
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import torch  # noqa: E402

from spmv_acc_b200 import make_options, synth  # noqa: E402
from sweep import make, time_plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--tiles", default="1024,2048")
ap.add_argument("--carve", default="-1,30,45,57,72,86,100")
ap.add_argument("--flags", default="0")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
csr = make(a.workload)
balg = synth.algorithmic_bytes(csr.rows, csr.cols, csr.nnz)
for T in [int(t) for t in a.tiles.split(",")]:
    for fl in [int(f) for f in a.flags.split(",")]:
        for c in [int(c) for c in a.carve.split(",")]:
            if c < 0:
                os.environ.pop("SPMV_B200_CARVEOUT", None)
            else:
                os.environ["SPMV_B200_CARVEOUT"] = str(c)
            ms, info = time_plan(csr, make_options(T, 0, 0, 0, fl), a.reps)
            print(json.dumps({"workload": a.workload, "tile": info.tile_nnz, "flags": fl, "carveout_pct": c,
                              "ms": round(ms, 4), "gbs": round(balg / ms / 1e6, 1), "smem": info.smem_bytes,
                              "kinds": list(info.tiles_per_kind)}), flush=True)
