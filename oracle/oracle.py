"""ctypes doors onto ``oracle/liboracle.so`` (plain-C restatement) and ``oracle/_ref/libref_oracle.so`` (the
reference's own functions compiled in place from /root/reference by ``oracle/Makefile``).

ORACLE — test infrastructure only; see ``oracle/host_spmv_port.c`` for the reference file:line each function follows.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE = Path("/root/reference")

__all__ = [
    "build", "have_ref", "port_host_spmv", "port_host_spmv_ax", "port_verify_y", "port_verify", "port_row_bound",
    "port_generate_vector", "port_merge_path_partition", "port_flat_break_points_v2", "port_analysis",
    "port_shard_bounds", "port_gather_stat", "port_tiled_spmv", "port_direct_arrays", "port_adaptive_choice", "ref_adaptive_choice", "have_ref_selector", "ADAPTIVE_CHOICES", "port_xstage", "ref_host_spmv", "ref_host_spmv_ax", "ref_verify_y",
    "ref_adaptive_plus_analyze", "ref_read", "ref_generate_vector", "best_host_spmv", "check_rows",
]

_f64 = np.float64
_i32 = np.int32


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def build(with_ref: bool = True) -> None:
    """Compile the C restatement and, when /root/reference is present, the reference itself (oracle/_ref)."""
    targets = ["all"]
    if with_ref and REFERENCE.exists():
        targets += ["ref", "ref-gpu", "ref-selector"]
        if (HERE.parent / "spmv_acc_b200" / "lib" / "libspmv_b200.so").exists():
            targets += ["ref-cli", "ref-harness"]
    res = subprocess.run(["make", "-s", "-C", str(HERE), *targets], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{res.stdout}\n{res.stderr}")


_port = None
_ref = None


def _port_lib():
    global _port
    if _port is None:
        so = HERE / "liboracle.so"
        if not so.exists():
            build(with_ref=False)
        _port = C.CDLL(str(so))
    return _port


def have_ref() -> bool:
    return (HERE / "_ref" / "libref_oracle.so").exists()


def _ref_lib():
    global _ref
    if _ref is None:
        so = HERE / "_ref" / "libref_oracle.so"
        if not so.exists():
            raise FileNotFoundError("oracle/_ref/libref_oracle.so is missing: run `make -C oracle ref` where "
                                    "/root/reference is mounted")
        _ref = C.CDLL(str(so))
        _ref.ref_read_csr_text.restype = C.c_void_p
        _ref.ref_read_bin2.restype = C.c_void_p
        _ref.ref_read_mtx.restype = C.c_void_p
    return _ref


class _VerifyResult(C.Structure):
    _fields_ = [("max_error", C.c_double), ("first_failed_at", C.c_int), ("failed_count", C.c_int)]


def _spmv_args(rowptr, col, val, x):
    return _c(rowptr, _i32), _c(col, _i32), _c(val, _f64), _c(x, _f64)


def _host_spmv(lib, name, alpha, beta, rowptr, col, val, x, y0, n):
    rowptr, col, val, x = _spmv_args(rowptr, col, val, x)
    m = rowptr.size - 1
    y = np.array(y0, dtype=_f64, copy=True)
    getattr(lib, name)(C.c_double(alpha), C.c_double(beta), _p(val, C.c_double), _p(rowptr, C.c_int),
                       _p(col, C.c_int), C.c_int(m), C.c_int(n), C.c_int(int(val.size)), _p(x, C.c_double),
                       _p(y, C.c_double))
    return y


def port_host_spmv(alpha, beta, rowptr, col, val, x, y0, n=None):
    """y = alpha*A*x + beta*y0 — restatement of cli/verification.cpp:56-66."""
    return _host_spmv(_port_lib(), "port_host_spmv_axpby", alpha, beta, rowptr, col, val, x, y0,
                      len(x) if n is None else n)


def ref_host_spmv(alpha, beta, rowptr, col, val, x, y0, n=None):
    """The reference's own host_spmv (cli/verification.cpp:56-66), compiled in place."""
    return _host_spmv(_ref_lib(), "ref_host_spmv_axpby", alpha, beta, rowptr, col, val, x, y0,
                      len(x) if n is None else n)


def best_host_spmv(alpha, beta, rowptr, col, val, x, y0):
    """The reference when it was compiled here, else the restatement."""
    return (ref_host_spmv if have_ref() else port_host_spmv)(alpha, beta, rowptr, col, val, x, y0)


def _host_spmv_ax(lib, name, rowptr, col, val, x):
    rowptr, col, val, x = _spmv_args(rowptr, col, val, x)
    m = rowptr.size - 1
    y = np.zeros(m, dtype=_f64)
    getattr(lib, name)(_p(val, C.c_double), _p(rowptr, C.c_int), _p(col, C.c_int), C.c_int(m), C.c_int(x.size),
                       C.c_int(int(val.size)), _p(x, C.c_double), _p(y, C.c_double))
    return y


def port_host_spmv_ax(rowptr, col, val, x):
    return _host_spmv_ax(_port_lib(), "port_host_spmv_ax", rowptr, col, val, x)


def ref_host_spmv_ax(rowptr, col, val, x):
    return _host_spmv_ax(_ref_lib(), "ref_host_spmv_ax", rowptr, col, val, x)


def _verify_y(lib, name, dy, hy):
    dy, hy = _c(dy, _f64), _c(hy, _f64)
    out = _VerifyResult()
    getattr(lib, name)(_p(dy, C.c_double), _p(hy, C.c_double), C.c_int(dy.size), C.byref(out))
    return {"max_error": out.max_error, "first_failed_at": out.first_failed_at, "failed_count": out.failed_count}


def port_verify_y(dy, hy):
    return _verify_y(_port_lib(), "port_verify_y", dy, hy)


def ref_verify_y(dy, hy):
    return _verify_y(_ref_lib(), "ref_verify_y", dy, hy)


def port_verify(dy, hy) -> int:
    dy, hy = _c(dy, _f64), _c(hy, _f64)
    return int(_port_lib().port_verify(_p(dy, C.c_double), _p(hy, C.c_double), C.c_int(dy.size)))


def port_row_bound(alpha, beta, rowptr, col, val, x, y0):
    rowptr, col, val, x = _spmv_args(rowptr, col, val, x)
    y0 = _c(y0, _f64)
    m = rowptr.size - 1
    out = np.zeros(m, dtype=_f64)
    _port_lib().port_row_bound(C.c_double(alpha), C.c_double(beta), _p(val, C.c_double), _p(rowptr, C.c_int),
                               _p(col, C.c_int), C.c_int(m), _p(x, C.c_double), _p(y0, C.c_double),
                               _p(out, C.c_double))
    return out


def check_rows(y, y_ref, bound, tol=1e-12):
    """North-star per-row test |y - y_ref| <= tol * bound. Returns (ok, worst_ratio, worst_row)."""
    y, y_ref, bound = _c(y, _f64), _c(y_ref, _f64), _c(bound, _f64)
    err = np.abs(y - y_ref)
    lim = tol * bound
    bad = ~(err <= lim)  # NaN counts as a failure
    ratio = np.where(lim > 0, err / np.where(lim > 0, lim, 1.0), np.where(err == 0, 0.0, np.inf))
    worst = int(np.argmax(ratio)) if ratio.size else -1
    return (not bool(bad.any())), (float(ratio[worst]) if ratio.size else 0.0), worst


def _gen(lib, prefix, n, seed):
    out = np.zeros(n, dtype=_f64)
    if seed is not None:
        getattr(lib, prefix + "srand")(C.c_uint(seed))
    getattr(lib, prefix + "generate_vector")(C.c_int(n), _p(out, C.c_double))
    return out


def port_generate_vector(n, seed=1):
    """cli/utils.hpp:46-56 — the reference never seeds, which equals srand(1)."""
    return _gen(_port_lib(), "port_", n, seed)


def ref_generate_vector(n, seed=1):
    return _gen(_ref_lib(), "ref_", n, seed)


def port_merge_path_partition(rowptr, count, items_per_block):
    rowptr = _c(rowptr, _i32)
    S = np.zeros(count, dtype=_i32)
    _port_lib().port_merge_path_partition(_p(rowptr, C.c_int), C.c_int(rowptr.size - 1), C.c_int(count),
                                          C.c_int(items_per_block), _p(S, C.c_int))
    return S


def port_flat_break_points_v2(rowptr, stride):
    rowptr = _c(rowptr, _i32)
    nnz = int(rowptr[-1])
    bp = np.full((nnz + stride - 1) // stride + 1, -1, dtype=_i32)
    _port_lib().port_flat_break_points_v2(_p(rowptr, C.c_int), C.c_int(rowptr.size - 1), C.c_int(stride),
                                          _p(bp, C.c_int))
    return bp


def port_analysis(rowptr, tile_nnz=2048, short_max=8, medium_max=128):
    """CPU restatement of the CUDA row analysis; returns a dict of numpy arrays with the export layout."""
    rowptr = _c(rowptr, _i32)
    m = rowptr.size - 1
    lib = _port_lib()
    nt = int(lib.port_analysis_ntiles(_p(rowptr, C.c_int), C.c_int(m), C.c_int(tile_nnz))) if m > 0 else 0
    n1 = nt + 1 if m > 0 else 0
    tile_row = np.zeros(n1, _i32)
    tile_elem = np.zeros(n1, _i32)
    tile_part = np.zeros(n1, _i32)
    tile_split = np.zeros(n1, np.uint8)
    tile_maxlen = np.zeros(nt, _i32)
    tile_kind = np.zeros(nt, np.uint8)
    row_bin = np.zeros(m, np.uint8)
    bin_rows = np.zeros(4, np.int64)
    bin_nnz = np.zeros(4, np.int64)
    cap = max(nt, 1)
    split = np.zeros(3 * cap, _i32)
    ns = lib.port_analysis(_p(rowptr, C.c_int), C.c_int(m), C.c_int(tile_nnz), C.c_int(short_max),
                           C.c_int(medium_max), _p(tile_row, C.c_int), _p(tile_elem, C.c_int),
                           _p(tile_split, C.c_ubyte), _p(tile_part, C.c_int), _p(tile_maxlen, C.c_int),
                           _p(tile_kind, C.c_ubyte), _p(row_bin, C.c_ubyte), _p(bin_rows, C.c_longlong),
                           _p(bin_nnz, C.c_longlong), _p(split, C.c_int), C.c_int(cap))
    ns = int(ns)
    split_rows = np.concatenate([split[0:ns], split[cap:cap + ns], split[2 * cap:2 * cap + ns]]).astype(_i32)
    return {
        "ntiles": nt, "tile_row": tile_row, "tile_elem": tile_elem, "tile_split": tile_split,
        "tile_part": tile_part, "tile_maxlen": tile_maxlen, "tile_kind": tile_kind, "row_bin": row_bin,
        "bin_rows": bin_rows, "bin_nnz": bin_nnz, "nsplit": ns, "split_rows": split_rows,
    }


def port_direct_arrays(rowptr, tile_row):
    """CPU restatement (numpy, integer only) of the extra analysis arrays of the direct form
    (spmv_acc_b200/csrc/analysis.cu: k_row_start_bits, k_nz_rows, k_desc_direct): row-start bit flags over the absolute
    element index, the ascending list of non-empty rows, and the number of non-empty rows in front of every tile.
    These play the role of the reference's per-block row bookkeeping (hip-flat break points, flat_imp.inl:107-152;
    merge-path S[t], merge_path_partition.h:7-17) for a kernel that never searches row pointers per element."""
    rowptr = _c(rowptr, _i32).astype(np.int64)
    lens = np.diff(rowptr)
    nz_rows = np.flatnonzero(lens > 0).astype(_i32)
    end = int(rowptr[-1]) if rowptr.size else 0
    bits = np.zeros((end + 31) // 32, np.uint32)
    starts = rowptr[:-1][lens > 0]
    np.bitwise_or.at(bits, starts >> 5, (np.uint32(1) << (starts & 31).astype(np.uint32)))
    tile_row = np.asarray(tile_row, dtype=np.int64)
    nzbase = np.searchsorted(nz_rows, tile_row[:-1], side="left").astype(_i32)
    return {"row_start_bits": bits, "nz_rows": nz_rows, "tile_nzbase": nzbase}


def port_xstage(col, tile_elem, span_max=65536, lines_max=256, seg_max=16):
    """CPU restatement of the staged-x analysis (spmv_acc_b200/csrc/analysis.cu:k_xstage_build): returns
    dict(failed, max_lines, lcol uint16 [nnz], xdesc int32 [32*ntiles]) with the export layout."""
    col, tile_elem = _c(col, _i32), _c(tile_elem, _i32)
    nt = tile_elem.size - 1
    base = int(tile_elem[0]) if tile_elem.size else 0
    nnz = int(tile_elem[-1]) - base if tile_elem.size else 0
    lcol = np.zeros(max(nnz, 1), np.uint16)
    xdesc = np.zeros(32 * max(nt, 1), _i32)
    best = C.c_int(0)
    failed = _port_lib().port_xstage(_p(col, C.c_int), _p(tile_elem, C.c_int), C.c_int(nt), C.c_int(base),
                                     C.c_int(span_max), C.c_int(lines_max), C.c_int(seg_max),
                                     lcol.ctypes.data_as(C.POINTER(C.c_ushort)), _p(xdesc, C.c_int), C.byref(best))
    return {"failed": int(failed), "max_lines": int(best.value), "lcol": lcol[:nnz], "xdesc": xdesc[:32 * nt]}


ADAPTIVE_CHOICES = ["vector-row, two data blocks", "adaptive line", "adaptive line-enhance", "adaptive flat",
                    "line-enhance"]


def port_adaptive_choice(rowptr) -> str:
    """Which kernel the reference's run-time selector would pick (src/acc/hip-adaptive/adaptive.cpp:16-67)."""
    rowptr = _c(rowptr, _i32)
    lib = _port_lib()
    lib.port_adaptive_choice.restype = C.c_int
    return ADAPTIVE_CHOICES[int(lib.port_adaptive_choice(_p(rowptr, C.c_int), C.c_int(rowptr.size - 1)))]


def have_ref_selector() -> bool:
    return (HERE / "_ref" / "libref_selector.so").exists()


def ref_adaptive_choice(rowptr) -> str:
    """The same decision taken by the reference's own adaptive.cpp, compiled in place (oracle/_ref/libref_selector.so)."""
    so = HERE / "_ref" / "libref_selector.so"
    if not so.exists():
        raise FileNotFoundError("oracle/_ref/libref_selector.so is missing: run `make -C oracle ref-selector` where "
                                "/root/reference is mounted")
    rowptr = _c(rowptr, _i32)
    lib = C.CDLL(str(so))
    lib.ref_adaptive_choice.restype = C.c_int
    return ADAPTIVE_CHOICES[int(lib.ref_adaptive_choice(_p(rowptr, C.c_int), C.c_int(rowptr.size - 1)))]


def port_gather_stat(rowptr, col, medium_max=128):
    """(active lanes, distinct 128-byte lines, sampled nnz, sampled nnz in long rows) of the sampling pass."""
    rowptr, col = _c(rowptr, _i32), _c(col, _i32)
    out = np.zeros(4, np.int64)
    _port_lib().port_gather_stat(_p(rowptr, C.c_int), _p(col, C.c_int), C.c_int(rowptr.size - 1),
                                 C.c_int(medium_max), _p(out, C.c_longlong))
    return tuple(int(v) for v in out)


def port_shard_bounds(rowptr, nshards):
    rowptr = _c(rowptr, _i32)
    out = np.zeros(nshards + 1, _i32)
    _port_lib().port_shard_bounds(_p(rowptr, C.c_int), C.c_int(rowptr.size - 1), C.c_int(nshards), _p(out, C.c_int))
    return out


def port_tiled_spmv(alpha, beta, rowptr, col, val, x, y0, ana, tile_nnz=2048, medium_max=128):
    """CPU emulation of the tile decomposition used by the CUDA kernels (structure check, see analysis_port.c)."""
    rowptr, col, val, x = _spmv_args(rowptr, col, val, x)
    m = rowptr.size - 1
    y = np.array(y0, dtype=_f64, copy=True)
    nt = ana["ntiles"]
    partials = np.zeros(max(2 * nt, 1), _f64)
    done = np.zeros(max(m, 1), np.uint8)
    tr, te, ts = _c(ana["tile_row"], _i32), _c(ana["tile_elem"], _i32), _c(ana["tile_split"], np.uint8)
    rc = _port_lib().port_tiled_spmv(C.c_double(alpha), C.c_double(beta), _p(val, C.c_double), _p(rowptr, C.c_int),
                                     _p(col, C.c_int), C.c_int(m), C.c_int(tile_nnz), C.c_int(medium_max),
                                     _p(tr, C.c_int), _p(te, C.c_int), _p(ts, C.c_ubyte), C.c_int(nt),
                                     _p(x, C.c_double), _p(y, C.c_double), _p(partials, C.c_double),
                                     _p(done, C.c_ubyte))
    if rc != 0:
        raise AssertionError(f"tile decomposition violates invariant {rc}")
    return y


def ref_adaptive_plus_analyze(rowptr, min_nnz_per_block=2048, vec=1):
    """The reference's csr_adaptive_plus_analyze_imp<int,512,vec> (csr_adaptive_plus_analyze.cpp:12-98)."""
    rowptr = _c(rowptr, _i32)
    m = rowptr.size - 1
    nnz = int(rowptr[-1])
    cap = nnz // max(min_nnz_per_block, 1) + m + 16
    bp = np.zeros(cap, _i32)
    first = np.zeros(m + 1, _i32)
    blocks = _ref_lib().ref_adaptive_plus_analyze(C.c_int(m), C.c_int(nnz), C.c_int(min_nnz_per_block), C.c_int(vec),
                                                  _p(rowptr, C.c_int), _p(bp, C.c_int), C.c_int(cap),
                                                  _p(first, C.c_int))
    if blocks < 0:
        raise RuntimeError("ref_adaptive_plus_analyze failed")
    return bp[:blocks + 1].copy(), first


def ref_read(path, fmt):
    """Read a matrix with the reference's own reader (fmt: csr | bin2 | mtx). Returns (rowptr, col, val, x|None, n)."""
    lib = _ref_lib()
    fn = {"csr": lib.ref_read_csr_text, "bin2": lib.ref_read_bin2, "mtx": lib.ref_read_mtx}[fmt]
    h = fn(str(path).encode())
    if not h:
        raise RuntimeError(f"reference reader failed on {path}")
    h = C.c_void_p(h)
    rows, cols, nnz, has_x = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.ref_csr_sizes(h, C.byref(rows), C.byref(cols), C.byref(nnz), C.byref(has_x))
    rowptr = np.zeros(rows.value + 1, _i32)
    col = np.zeros(nnz.value, _i32)
    val = np.zeros(nnz.value, _f64)
    x = np.zeros(cols.value, _f64) if has_x.value else None
    lib.ref_csr_copy(h, _p(rowptr, C.c_int), _p(col, C.c_int), _p(val, C.c_double),
                     _p(x, C.c_double) if x is not None else None)
    lib.ref_csr_free(h)
    return rowptr, col, val, x, cols.value
