"""ctypes bindings of the C ABI declared in ``include/spmv_b200.h``.

The CUDA library must exist: there is no CPU fallback anywhere in this package. If ``lib/libspmv_b200.so`` is
missing, importing the bindings raises with the build command to run.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_DIR = PKG / "lib"

# every symbol include/spmv_b200.h declares (tests/test_abi.py checks the header against this list)
ABI_SYMBOLS = [
    "spmv_b200_abi_version", "spmv_b200_last_error", "spmv_b200_plan_create", "spmv_b200_execute",
    "spmv_b200_execute_tiles", "spmv_b200_execute_push", "spmv_b200_execute_tiles_push", "spmv_b200_plan_tile_col_range",
    "spmv_b200_plan_set_comm_sms",
    "spmv_b200_halo_loop_create", "spmv_b200_halo_loop_run", "spmv_b200_halo_loop_sync", "spmv_b200_halo_loop_get_info",
    "spmv_b200_halo_loop_destroy", "spmv_b200_enable_peer_access", "spmv_b200_peer_alloc", "spmv_b200_peer_open",
    "spmv_b200_peer_close", "spmv_b200_peer_free", "spmv_b200_cache_revalidations",
    "spmv_b200_plan_destroy", "spmv_b200_plan_get_info", "spmv_b200_plan_export", "spmv_b200_csr_spmv",
    "spmv_b200_sparse_spmv", "spmv_b200_cache_invalidate", "spmv_b200_cache_size", "spmv_b200_hostmat_create", "spmv_b200_hostmat_create_device",
    "spmv_b200_hostmat_spmv", "spmv_b200_hostmat_x_range", "spmv_b200_hostmat_destroy", "spmv_b200_host_spmv", "spmv_b200_coo_to_csr", "spmv_b200_shard_bounds",
    "spmv_b200_col_block_bitmap",
]


class Options(C.Structure):
    _fields_ = [("tile_nnz", C.c_int32), ("short_max", C.c_int32), ("medium_max", C.c_int32),
                ("vec_div", C.c_int32), ("flags", C.c_uint32)]


MAX_PUSH = 8
IPC_HANDLE_BYTES = 64


class Push(C.Structure):
    _fields_ = [("count", C.c_int32), ("multicast_mask", C.c_uint32), ("row_lo", C.c_int32 * MAX_PUSH),
                ("row_hi", C.c_int32 * MAX_PUSH), ("dst", C.c_void_p * MAX_PUSH)]


MAX_RANGES = 16


class HaloLoopDesc(C.Structure):
    _fields_ = [("plan", C.c_void_p), ("buf", C.c_void_p * 2), ("row_lo", C.c_int32), ("row_hi", C.c_int32),
                ("n_neigh", C.c_int32), ("wait_flags", C.c_void_p * MAX_PUSH), ("signal_flags", C.c_void_p * MAX_PUSH),
                ("push", Push * 2), ("n_boundary", C.c_int32), ("n_interior", C.c_int32),
                ("boundary", C.c_int32 * (2 * MAX_RANGES)), ("interior", C.c_int32 * (2 * MAX_RANGES)),
                ("flags", C.c_uint32)]


class HaloLoopInfo(C.Structure):
    _fields_ = [("iterations_enqueued", C.c_int64), ("single_launch", C.c_int32), ("uses_graph", C.c_int32),
                ("launches_per_iteration", C.c_int32), ("boundary_row_blocks", C.c_int32)]


HALO_NO_GRAPH = 1
HALO_MULTI_LAUNCH = 2
HALO_ALIGN_PUSH = 4
ERR_TIMEOUT = 4


class PlanInfo(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("nnz", C.c_int64), ("tile_nnz", C.c_int32),
                ("short_max", C.c_int32), ("medium_max", C.c_int32), ("vec_div", C.c_int32), ("flags", C.c_uint32),
                ("uses_tma", C.c_int32), ("ntiles", C.c_int32), ("tiles_per_kind", C.c_int32 * 3),
                ("nsplit_rows", C.c_int32), ("launches_per_execute", C.c_int32), ("direct", C.c_int32), ("bin_rows", C.c_int64 * 4),
                ("bin_nnz", C.c_int64 * 4), ("gather_active", C.c_int64), ("gather_lines", C.c_int64),
                ("smem_bytes", C.c_int64), ("workspace_bytes", C.c_int64), ("xstage", C.c_int32),
                ("xstage_lines", C.c_int32), ("ring_ctas", C.c_int32), ("ring_stages", C.c_int32)]


FLAG_NO_TMA = 1
FLAG_BETA0_SKIP_Y = 2
FLAG_L2_PERSIST_X = 4
FLAG_DIRECT = 0x40
FLAG_NO_DIRECT = 0x80
FLAG_NO_XSTAGE = 0x40000

EXPORT_IDS = {"tile_row": 0, "tile_elem": 1, "tile_split": 2, "tile_kind": 3, "tile_part": 4, "row_bin": 5,
              "split_rows": 6, "tile_maxlen": 7, "row_start_bits": 8, "nz_rows": 9, "tile_nzbase": 10, "lcol": 11,
              "xdesc": 12}
EXPORT_DTYPES = {"tile_row": "int32", "tile_elem": "int32", "tile_split": "uint8", "tile_kind": "uint8",
                 "tile_part": "int32", "row_bin": "uint8", "split_rows": "int32", "tile_maxlen": "int32",
                 "row_start_bits": "uint32", "nz_rows": "int32", "tile_nzbase": "int32", "lcol": "uint16",
                 "xdesc": "int32"}

_lib = None
_gen = None
_ctx = None


def _load(name: str) -> C.CDLL:
    so = LIB_DIR / name
    if not so.exists():
        raise RuntimeError(
            f"{so} is missing: the CUDA extension has not been built. Run `python -m spmv_acc_b200.build` "
            f"(needs nvcc; cross-compiles for sm_100a without a GPU). There is no CPU fallback.")
    return C.CDLL(str(so))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = _load("libspmv_b200.so")
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.spmv_b200_abi_version.restype = C.c_int
        L.spmv_b200_last_error.restype = C.c_char_p
        L.spmv_b200_plan_create.argtypes = [C.POINTER(vp), i32, i32, i64, vp, vp, vp, C.POINTER(Options), vp]
        L.spmv_b200_execute.argtypes = [vp, dbl, dbl, vp, vp, vp]
        L.spmv_b200_execute_tiles.argtypes = [vp, dbl, dbl, vp, vp, i32, i32, vp]
        L.spmv_b200_execute_push.argtypes = [vp, dbl, dbl, vp, vp, C.POINTER(Push), vp]
        L.spmv_b200_execute_tiles_push.argtypes = [vp, dbl, dbl, vp, vp, i32, i32, C.POINTER(Push), vp]
        L.spmv_b200_plan_tile_col_range.argtypes = [vp, vp, vp, vp]
        L.spmv_b200_plan_set_comm_sms.argtypes = [vp, i32]
        L.spmv_b200_halo_loop_create.argtypes = [C.POINTER(vp), C.POINTER(HaloLoopDesc)]
        L.spmv_b200_halo_loop_run.argtypes = [vp, i32, vp]
        L.spmv_b200_halo_loop_sync.argtypes = [vp, vp]
        L.spmv_b200_halo_loop_get_info.argtypes = [vp, C.POINTER(HaloLoopInfo)]
        L.spmv_b200_halo_loop_destroy.argtypes = [vp]
        L.spmv_b200_enable_peer_access.argtypes = [i32]
        L.spmv_b200_peer_alloc.argtypes = [C.POINTER(vp), i64, C.c_char_p]
        L.spmv_b200_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.spmv_b200_peer_close.argtypes = [vp]
        L.spmv_b200_peer_free.argtypes = [vp]
        L.spmv_b200_plan_destroy.argtypes = [vp]
        L.spmv_b200_plan_get_info.argtypes = [vp, C.POINTER(PlanInfo)]
        L.spmv_b200_plan_export.argtypes = [vp, i32, vp, i64, C.POINTER(i64)]
        L.spmv_b200_csr_spmv.argtypes = [i32, dbl, dbl, i32, i32, i32, vp, vp, vp, vp, vp]
        L.spmv_b200_sparse_spmv.argtypes = [i32, dbl, dbl, i32, i32, vp, vp, vp, vp, vp]
        L.spmv_b200_hostmat_create.argtypes = [C.POINTER(vp), i32, i32, i64, vp, vp, vp, C.POINTER(Options)]
        L.spmv_b200_hostmat_create_device.argtypes = [C.POINTER(vp), i32, i32, i64, vp, vp, vp, C.POINTER(Options)]
        L.spmv_b200_hostmat_spmv.argtypes = [vp, dbl, dbl, vp, vp]
        L.spmv_b200_hostmat_destroy.argtypes = [vp]
        L.spmv_b200_hostmat_x_range.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        L.spmv_b200_host_spmv.argtypes = [dbl, dbl, i32, i32, i64, vp, vp, vp, vp, vp]
        L.spmv_b200_coo_to_csr.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp]
        L.spmv_b200_shard_bounds.argtypes = [i32, i64, vp, i32, vp, vp]
        L.spmv_b200_col_block_bitmap.argtypes = [i64, vp, i32, i32, vp, vp]
        for s in ABI_SYMBOLS:
            if s not in ("spmv_b200_last_error",):
                getattr(L, s).restype = C.c_int
        L.spmv_b200_cache_revalidations.restype = C.c_int64
        L.spmv_b200_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def gen() -> C.CDLL:
    global _gen
    if _gen is None:
        G = _load("libspmv_b200_gen.so")
        vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
        G.spmv_b200_gen_vector.argtypes = [i64, u64, vp, vp]
        G.spmv_b200_gen_stencil2d_counts.argtypes = [i32, i32, i64, i64, vp, vp]
        G.spmv_b200_gen_stencil2d_fill.argtypes = [i32, i32, i64, i64, vp, vp, vp, vp]
        G.spmv_b200_gen_stencil3d_counts.argtypes = [i32, i64, i64, vp, vp]
        G.spmv_b200_gen_stencil3d_fill.argtypes = [i32, i64, i64, vp, vp, vp, vp]
        G.spmv_b200_gen_uniform_fill.argtypes = [i64, i64, i32, i32, u64, vp, vp, vp]
        G.spmv_b200_gen_rmat_edges.argtypes = [i32, i64, dbl, dbl, dbl, u64, vp, vp]
        G.spmv_b200_gen_rmat_finish.argtypes = [i32, i64, vp, u64, vp, vp, vp, vp]
        _gen = G
    return _gen


def ctx() -> C.CDLL:
    global _ctx
    if _ctx is None:
        X = _load("libspmv_b200_ctx.so")
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        X.spmv_b200_ctx_cusparse_create.argtypes = [C.POINTER(vp), i32, i32, i64, vp, vp, vp, vp, vp, i32]
        X.spmv_b200_ctx_cusparse_spmv.argtypes = [vp, dbl, dbl, vp]
        X.spmv_b200_ctx_cusparse_destroy.argtypes = [vp]
        X.spmv_b200_ctx_cub_create.argtypes = [C.POINTER(vp), i32, i32, i32, vp, vp, vp, vp, vp]
        X.spmv_b200_ctx_cub_spmv.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
        X.spmv_b200_ctx_cub_destroy.argtypes = [vp]
        X.spmv_b200_ctx_gather_bound.argtypes = [i64, vp, vp, vp, vp, i32, i32, i32, vp]
        X.spmv_b200_ctx_gather4_bound.argtypes = [i64, vp, vp, i64, vp, i32, i32, vp]
        X.spmv_b200_ctx_gather_affine.argtypes = [i64, vp, vp, vp, vp, vp, i32, i32, vp]
        _ctx = X
    return _ctx


class SpmvB200Error(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().spmv_b200_last_error()
        raise SpmvB200Error(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")
