"""Experiment: column-blocked SpMV for matrices whose x does not stay in L2 (C3): split A into K column blocks
(each a CSR matrix over all rows) and run K SpMVs, y accumulating. Uses the production kernels unchanged."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch  # noqa: E402

from spmv_acc_b200 import CsrDesc, SpmvPlan, make_options, synth  # noqa: E402
from sweep import make  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "c3"
csr = make(w)
x = synth.vector_device(csr.cols, 2)
lens = (csr.rowptr[1:] - csr.rowptr[:-1]).to(torch.int64)
row_of = torch.repeat_interleave(torch.arange(csr.rows, device="cuda"), lens)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


for K in (1, 2, 3, 4):
    plans, keep = [], []
    for b in range(K):
        lo, hi = csr.cols * b // K, csr.cols * (b + 1) // K
        mask = (csr.col >= lo) & (csr.col < hi)
        cnt = torch.zeros(csr.rows, dtype=torch.int64, device="cuda").index_add_(0, row_of[mask], torch.ones(int(mask.sum()), dtype=torch.int64, device="cuda"))
        rp = torch.zeros(csr.rows + 1, dtype=torch.int64, device="cuda")
        torch.cumsum(cnt, 0, out=rp[1:])
        rp = rp.to(torch.int32)
        col, val = csr.col[mask].contiguous(), csr.val[mask].contiguous()
        keep.append((rp, col, val))
        plans.append(SpmvPlan(CsrDesc(csr.rows, csr.cols, int(col.numel()), rp, col, val)))
    y = synth.vector_device(csr.rows, 3)

    def run():
        for i, p in enumerate(plans):
            p.execute(1.0, 1.0, x, y)

    ms = timeit(run)
    infos = [p.info() for p in plans]
    print(f"{w} K={K}: {ms:.4f} ms total; tiles {[i.tile_nnz for i in infos]} kinds {[list(i.tiles_per_kind) for i in infos]}", flush=True)
    for p in plans:
        p.destroy()
    del plans, keep
    torch.cuda.empty_cache()
