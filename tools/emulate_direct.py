"""Lane-by-lane CPU emulation of the direct kernel (k_spmv_warp, spmv_acc_b200/csrc/kernels.cu) in numpy.

Development aid, not product code: there is no GPU in the build container, so the segmented-sum logic of the kernel
(flag-word masks, the bit-test predicate of the segmented scan, ordinals into nz_rows, head / tail fragments of split
rows, the fix-up pass) is replayed here with 32-element vectors standing in for the lanes of a warp and compared with
a plain CSR loop. `python tools/emulate_direct.py` runs a set of small matrices through it.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

LANES = np.arange(32)


def warp_sum(v):
    v = v.copy()
    off = 16
    while off > 0:
        v = v + v[LANES ^ off]
        off >>= 1
    return v


def shfl_up(v, off):
    out = v.copy()
    out[off:] = v[:-off]
    return out


def emulate_tile(a, ti):
    """One warp = one row block. `a` is a dict with the kernel arguments."""
    r0, r1, e0, e1, nzi, _tail, t, flags = (int(x) for x in a["desc"][ti])
    y, partials = a["y"], a["partials"]
    col, val, x, bits, nz_rows, rowptr = a["col"], a["val"], a["x"], a["bits"], a["nz_rows"], a["rowptr"]
    alpha, beta = a["alpha"], a["beta"]

    def finish_row(row, s):
        y[row] = alpha * s + beta * y[row]

    acc = np.zeros(32)
    started = False
    rb = e0 & ~31
    span = e1 - e0
    while rb < e1:
        i0 = rb + LANES
        p = np.zeros((4, 32))
        for j in range(4):
            i = i0 + 32 * j
            ok = ((i - e0) >= 0) & ((i - e0) < span)
            idx = np.where(ok, i, 0)
            c = np.where(ok, col[idx], -1)
            xv = np.where(c >= 0, x[np.where(c >= 0, c, 0)], 0.0)
            vv = np.where(ok, val[idx], 0.0)
            p[j] = xv * vv
        # flag word per lane (lane & 3), masked
        wq = np.zeros(32, dtype=np.uint64)
        for l in range(32):
            w = int(bits[(rb >> 5) + (l & 3)]) if (rb >> 5) + (l & 3) < bits.size else 0
            lo = e0 - (rb + 32 * (l & 3))
            hi = e1 - (rb + 32 * (l & 3))
            keep = 0 if hi <= 0 else (0xFFFFFFFF if hi >= 32 else (1 << hi) - 1)
            if lo > 0:
                keep &= 0 if lo >= 32 else (~((1 << lo) - 1)) & 0xFFFFFFFF
            wq[l] = w & keep
        if not np.any(wq != 0):
            acc = acc + ((p[0] + p[1]) + (p[2] + p[3]))
            rb += 128
            continue
        for j in range(4):
            wj = int(wq[j])
            if wj == 0:
                acc = acc + p[j]
                continue
            first = (wj & -wj).bit_length() - 1
            open_ = warp_sum(acc + np.where(LANES < first, p[j], 0.0))
            if started:
                finish_row(int(nz_rows[nzi - 1]), open_[0])
            elif flags & 1:
                partials[2 * t] = open_[0]
            cnt = bin(wj).count("1")
            last = first
            if cnt > 1:
                s = p[j].copy()
                off = 1
                while off < 32:
                    o = shfl_up(s, off)
                    for l in range(32):
                        if l >= off and ((wj >> (l - off + 1)) & ((1 << off) - 1)) == 0:
                            s[l] = s[l] + o[l]
                    off <<= 1
                last = wj.bit_length() - 1
                for l in range(32):
                    if l >= first and l < last and ((wj >> (l + 1)) & 1):
                        ordinal = bin(wj & ((2 << l) - 1)).count("1")
                        finish_row(int(nz_rows[nzi + ordinal - 1]), s[l])
            acc = np.where(LANES >= last, p[j], 0.0)
            nzi += cnt
            started = True
        rb += 128
    for r in range(r0, r1):
        if rowptr[r] == rowptr[r + 1]:
            finish_row(r, 0.0)
    open_ = warp_sum(acc)[0]
    if not started:
        if flags & 1:
            partials[2 * t] = open_
    elif flags & 2:
        partials[2 * t + 1] = open_
    else:
        finish_row(int(nz_rows[nzi - 1]), open_)


def direct_spmv(alpha, beta, rowptr, col, val, x, y0, T=2048, medium_max=128):
    import oracle
    ana = oracle.port_analysis(rowptr, T, 8, medium_max)
    nt = ana["ntiles"]
    tr, te, ts = ana["tile_row"], ana["tile_elem"], ana["tile_split"]
    d = oracle.port_direct_arrays(rowptr, tr)
    desc = np.zeros((nt, 8), dtype=np.int64)
    for t in range(nt):
        fl = (1 if ts[t] else 0) | (2 if ts[t + 1] else 0)
        desc[t] = (tr[t], tr[t + 1], te[t], te[t + 1], d["tile_nzbase"][t], 0, t, fl)
    bits = np.concatenate([d["row_start_bits"], np.zeros(16, np.uint32)])
    y = np.array(y0, dtype=np.float64, copy=True)
    partials = np.zeros(2 * max(nt, 1))
    a = dict(desc=desc, y=y, partials=partials, col=col, val=val, x=x, bits=bits, nz_rows=d["nz_rows"], rowptr=rowptr,
             alpha=alpha, beta=beta)
    for t in range(nt):
        emulate_tile(a, t)
    sr = ana["split_rows"]
    ns = ana["nsplit"]
    for k in range(ns):
        row, t0, t1 = int(sr[k]), int(sr[ns + k]), int(sr[2 * ns + k])
        s = partials[2 * t0 + 1]
        for tt in range(t0 + 1, t1 + 1):
            s += partials[2 * tt]
        y[row] = alpha * s + beta * y[row]
    return y, ana


def main():
    import oracle
    from spmv_acc_b200 import synth
    rng = np.random.default_rng(7)
    cases = [("rmat11", synth.rmat_numpy(11, 16, seed=1)), ("uniform", synth.uniform_numpy(300, 500, 32, seed=1)),
             ("stencil2d", synth.stencil2d_numpy(24)), ("circuit", synth.circuit_numpy(700, 3200, seed=3))]
    # ragged: empty rows, rows of one, a few very long rows
    lens = rng.integers(0, 4, size=400)
    lens[[17, 120, 300]] = [900, 3000, 260]
    lens[350:] = 0
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nnz = int(rp[-1])
    ragged = synth.Csr(400, 1000, rp, np.sort(rng.integers(0, 1000, size=nnz)).astype(np.int32),
                       rng.uniform(-1, 1, size=nnz))
    cases.append(("ragged", ragged))
    worst = 0.0
    for name, h in cases:
        for T in (256, 512, 2048):
            x, y0 = synth.vector_numpy(h.cols, 2), synth.vector_numpy(h.rows, 3)
            y, ana = direct_spmv(0.75, -0.5, h.rowptr, h.col, h.val, x, y0, T=T)
            y_ref = oracle.port_host_spmv(0.75, -0.5, h.rowptr, h.col, h.val, x, y0)
            bound = oracle.port_row_bound(0.75, -0.5, h.rowptr, h.col, h.val, x, y0)
            ok, ratio, row = oracle.check_rows(y, y_ref, bound, 1e-12)
            print(f"{name:10s} T={T:5d} tiles={ana['ntiles']:4d} split={ana['nsplit']:3d} ok={ok} ratio={ratio:.3g}")
            assert ok, (name, T, row)
            worst = max(worst, ratio)
    print("emulation OK, worst error / bound =", worst)


if __name__ == "__main__":
    main()
