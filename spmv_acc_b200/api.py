"""Host-side mirror of the reference's kernel API for the `cuda-b200` strategy.

Names and argument meaning follow the reference (all citations relative to the reference tree):
  * ``CsrDesc``            <-> ``csr_desc<int,double>``  (src/acc/api/types.h:25-41)
  * ``sparse_csr_spmv``    <-> ``sparse_csr_spmv(trans, alpha, beta, h_csr_desc, d_csr_desc, dx, dy)``
                               (src/acc/api/spmv.h:20-21, dispatch in src/acc/strategy_picker.cpp:19-65)
  * ``sparse_spmv``        <-> deprecated 10-argument entry (src/acc/api/spmv.h:27-28, spmv_imp.cpp:10-18)
  * ``SpmvPlan``           <-> analyze / kernel / destroy of csr-adaptive-plus
                               (src/acc/hip-csr-adaptive-plus/csr_adaptive_plus_spmv.cpp:16-73)
  * ``HostMatrix``         <-> the CLI's host-buffer pattern (cli/utils.hpp:94-116, cli/main.cpp:99-118)

Everything calls the C ABI of ``libspmv_b200.so`` through ctypes with raw device pointers; torch is used only to
own device memory and streams. y is updated in place, like the reference. Errors raise ``SpmvB200Error`` (the C++
launcher raises ``std::runtime_error``, which the reference harness catches: benchmark/csr_spmv.hpp:52-62).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

from . import _lib
from ._lib import FLAG_BETA0_SKIP_Y, FLAG_DIRECT, FLAG_L2_PERSIST_X, FLAG_NO_DIRECT, FLAG_NO_TMA, FLAG_NO_XSTAGE, Options, PlanInfo, SpmvB200Error, check  # noqa: F401

operation_none = 0       # src/acc/api/types.h:8
operation_transpose = 1


def _ptr(t: Any) -> int:
    """Raw address of a torch tensor / numpy array / int."""
    if t is None:
        return 0
    if isinstance(t, int):
        return t
    if hasattr(t, "data_ptr"):
        return int(t.data_ptr())
    if isinstance(t, np.ndarray):
        return int(t.ctypes.data)
    raise TypeError(f"cannot take the address of {type(t)}")


def _current_stream() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)


@dataclass
class CsrDesc:
    """Non-owning CSR view: rows, cols, nnz and three arrays (int32 row_ptr[rows+1], int32 col_index[nnz], fp64 values[nnz])."""
    rows: int
    cols: int
    nnz: int
    row_ptr: Any
    col_index: Any
    values: Any

    def as_const(self) -> "CsrDesc":  # var_csr_desc::as_const, src/acc/api/types.h:22
        return self


def _require_device(t: Any, name: str, dtype: str) -> None:
    if hasattr(t, "is_cuda"):
        if not t.is_cuda:
            raise SpmvB200Error(f"{name} must live in device memory (got a CPU tensor); there is no CPU fallback")
        if str(t.dtype) != f"torch.{dtype}":
            raise SpmvB200Error(f"{name} must be {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise SpmvB200Error(f"{name} must be contiguous")


def sparse_csr_spmv(trans: int, alpha: float, beta: float, h_csr_desc: Optional[CsrDesc], d_csr_desc: CsrDesc,
                    dx: Any, dy: Any) -> None:
    """y = alpha*A*x + beta*y on the current device, in place on ``dy``; returns without synchronising.

    ``h_csr_desc`` is accepted for signature compatibility and never dereferenced (it may alias device memory when
    entered through ``sparse_spmv``, see src/acc/api/spmv_imp.cpp:14-17)."""
    d = d_csr_desc
    _require_device(d.row_ptr, "d_csr_desc.row_ptr", "int32")
    _require_device(d.col_index, "d_csr_desc.col_index", "int32")
    _require_device(d.values, "d_csr_desc.values", "float64")
    _require_device(dx, "dx", "float64")
    _require_device(dy, "dy", "float64")
    rc = _lib.lib().spmv_b200_csr_spmv(int(trans), float(alpha), float(beta), int(d.rows), int(d.cols), int(d.nnz),
                                       _ptr(d.row_ptr), _ptr(d.col_index), _ptr(d.values), _ptr(dx), _ptr(dy))
    check(rc, "sparse_csr_spmv")


def sparse_spmv(htrans: int, halpha: float, hbeta: float, hm: int, hn: int, rowptr: Any, colindex: Any, value: Any,
                x: Any, y: Any) -> None:
    """The deprecated C-style entry of the reference with its exact argument list (device pointers)."""
    for t, nme, dt in ((rowptr, "rowptr", "int32"), (colindex, "colindex", "int32"), (value, "value", "float64"),
                       (x, "x", "float64"), (y, "y", "float64")):
        _require_device(t, nme, dt)
    rc = _lib.lib().spmv_b200_sparse_spmv(int(htrans), float(halpha), float(hbeta), int(hm), int(hn), _ptr(rowptr),
                                          _ptr(colindex), _ptr(value), _ptr(x), _ptr(y))
    check(rc, "sparse_spmv")


def cache_invalidate() -> None:
    check(_lib.lib().spmv_b200_cache_invalidate(), "cache_invalidate")


def cache_size() -> int:
    return int(_lib.lib().spmv_b200_cache_size())


def make_options(tile_nnz: int = 0, short_max: int = 0, medium_max: int = 0, vec_div: int = 0,
                 flags: int = 0) -> Options:
    return Options(tile_nnz, short_max, medium_max, vec_div, flags)


class SpmvPlan:
    """One-time row analysis + repeated execution (analyze / kernel / destroy)."""

    def __init__(self, d_csr_desc: CsrDesc, options: Optional[Options] = None, stream: Optional[int] = None):
        d = d_csr_desc
        _require_device(d.row_ptr, "row_ptr", "int32")
        _require_device(d.col_index, "col_index", "int32")
        _require_device(d.values, "values", "float64")
        self._keep = (d.row_ptr, d.col_index, d.values)  # the plan borrows these device arrays
        self._h = C.c_void_p()
        self.desc = d
        rc = _lib.lib().spmv_b200_plan_create(C.byref(self._h), int(d.rows), int(d.cols), int(d.nnz),
                                              _ptr(d.row_ptr), _ptr(d.col_index), _ptr(d.values),
                                              C.byref(options) if options is not None else None,
                                              _current_stream() if stream is None else stream)
        check(rc, "plan_create")

    def execute(self, alpha: float, beta: float, dx: Any, dy: Any, stream: Optional[int] = None) -> None:
        if not self._h:
            raise SpmvB200Error("plan was destroyed")
        _require_device(dx, "dx", "float64")
        _require_device(dy, "dy", "float64")
        rc = _lib.lib().spmv_b200_execute(self._h, float(alpha), float(beta), _ptr(dx), _ptr(dy),
                                          _current_stream() if stream is None else stream)
        check(rc, "execute")

    def set_comm_sms(self, sms: int) -> None:
        """Leave ``sms`` SMs free for a collective running beside the SpMV launches (persistent form only)."""
        if not self._h:
            raise SpmvB200Error("plan was destroyed")
        check(_lib.lib().spmv_b200_plan_set_comm_sms(self._h, int(sms)), "plan_set_comm_sms")

    def execute_tiles(self, alpha: float, beta: float, dx: Any, dy: Any, tile_lo: int, tile_hi: int,
                      stream: Optional[int] = None) -> None:
        """Row blocks [tile_lo, tile_hi) only (``export("tile_row")`` gives their row ranges); no split rows allowed."""
        if not self._h:
            raise SpmvB200Error("plan was destroyed")
        _require_device(dx, "dx", "float64")
        _require_device(dy, "dy", "float64")
        rc = _lib.lib().spmv_b200_execute_tiles(self._h, float(alpha), float(beta), _ptr(dx), _ptr(dy), int(tile_lo),
                                                int(tile_hi), _current_stream() if stream is None else stream)
        check(rc, "execute_tiles")

    def execute_push(self, alpha: float, beta: float, dx: Any, dy: Any, push: list, stream: Optional[int] = None) -> None:
        """execute + fused halo push: ``push`` is a list of (row_lo, row_hi, dst_address); rows in [row_lo, row_hi) are
        also stored to ``dst_address[row]`` (8-byte elements; typically another GPU's memory mapped through CUDA IPC)."""
        if not self._h:
            raise SpmvB200Error("plan was destroyed")
        if len(push) > _lib.MAX_PUSH:
            raise SpmvB200Error(f"at most {_lib.MAX_PUSH} push ranges are supported")
        _require_device(dx, "dx", "float64")
        _require_device(dy, "dy", "float64")
        ps = _fill_push(_lib.Push(), push)
        rc = _lib.lib().spmv_b200_execute_push(self._h, float(alpha), float(beta), _ptr(dx), _ptr(dy), C.byref(ps),
                                               _current_stream() if stream is None else stream)
        check(rc, "execute_push")

    def execute_tiles_push(self, alpha: float, beta: float, dx: Any, dy: Any, tile_lo: int, tile_hi: int, push: list,
                           stream: Optional[int] = None) -> None:
        """Row blocks [tile_lo, tile_hi) with the fused halo push of ``execute_push``."""
        if not self._h:
            raise SpmvB200Error("plan was destroyed")
        if len(push) > _lib.MAX_PUSH:
            raise SpmvB200Error(f"at most {_lib.MAX_PUSH} push ranges are supported")
        _require_device(dx, "dx", "float64")
        _require_device(dy, "dy", "float64")
        ps = _fill_push(_lib.Push(), push)
        rc = _lib.lib().spmv_b200_execute_tiles_push(self._h, float(alpha), float(beta), _ptr(dx), _ptr(dy),
                                                     int(tile_lo), int(tile_hi), C.byref(ps),
                                                     _current_stream() if stream is None else stream)
        check(rc, "execute_tiles_push")

    def tile_col_range(self):
        """(min, max) column index referenced by every row block (int32 arrays of length ntiles)."""
        nt = self.info().ntiles
        lo, hi = np.empty(nt, np.int32), np.empty(nt, np.int32)
        if nt:
            check(_lib.lib().spmv_b200_plan_tile_col_range(self._h, lo.ctypes.data, hi.ctypes.data, _current_stream()),
                  "plan_tile_col_range")
        return lo, hi

    def info(self) -> PlanInfo:
        out = PlanInfo()
        check(_lib.lib().spmv_b200_plan_get_info(self._h, C.byref(out)), "plan_get_info")
        return out

    def export(self, name: str) -> np.ndarray:
        what = _lib.EXPORT_IDS[name]
        nbytes = C.c_int64(0)
        check(_lib.lib().spmv_b200_plan_export(self._h, what, None, 0, C.byref(nbytes)), "plan_export(size)")
        dt = np.dtype(_lib.EXPORT_DTYPES[name])
        out = np.zeros(nbytes.value // dt.itemsize, dtype=dt)
        if nbytes.value:
            check(_lib.lib().spmv_b200_plan_export(self._h, what, out.ctypes.data, nbytes.value, None), "plan_export")
        return out

    def destroy(self) -> None:
        if self._h:
            h, self._h = self._h, C.c_void_p()
            check(_lib.lib().spmv_b200_plan_destroy(h), "plan_destroy")

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class HostMatrix:
    """Matrix given in host memory: uploaded + analysed once; every ``spmv`` copies x, y0 in and y out."""

    def __init__(self, rows: int, cols: int, rowptr: Any, colindex: Any, value: Any,
                 options: Optional[Options] = None):
        """Host arrays are uploaded; device tensors (the CLI's create_device_data done by the caller) are used in
        place and must outlive this object."""
        self._h = C.c_void_p()
        self.rows, self.cols, self.nnz = int(rows), int(cols), int(len(value))
        opt = C.byref(options) if options is not None else None
        on_device = [bool(getattr(t, "is_cuda", False)) for t in (rowptr, colindex, value)]
        if all(on_device):
            self._keep = (rowptr, colindex, value)
            rc = _lib.lib().spmv_b200_hostmat_create_device(C.byref(self._h), self.rows, self.cols, self.nnz,
                                                            _ptr(rowptr), _ptr(colindex), _ptr(value), opt)
            check(rc, "hostmat_create_device")
        elif any(on_device):
            raise SpmvB200Error("HostMatrix: rowptr / colindex / value must be all host or all device arrays")
        else:
            rc = _lib.lib().spmv_b200_hostmat_create(C.byref(self._h), self.rows, self.cols, self.nnz, _ptr(rowptr),
                                                     _ptr(colindex), _ptr(value), opt)
            check(rc, "hostmat_create")

    def spmv(self, alpha: float, beta: float, h_x: Any, h_y: Any) -> None:
        check(_lib.lib().spmv_b200_hostmat_spmv(self._h, float(alpha), float(beta), _ptr(h_x), _ptr(h_y)),
              "hostmat_spmv")

    def x_range(self):
        """Columns [lo, hi) the matrix references: the only part of ``h_x`` that ``spmv`` reads and copies."""
        lo, hi = C.c_int32(), C.c_int32()
        check(_lib.lib().spmv_b200_hostmat_x_range(self._h, C.byref(lo), C.byref(hi)), "hostmat_x_range")
        return int(lo.value), int(hi.value)

    def destroy(self) -> None:
        if self._h:
            h, self._h = self._h, C.c_void_p()
            check(_lib.lib().spmv_b200_hostmat_destroy(h), "hostmat_destroy")

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def host_spmv(alpha: float, beta: float, rows: int, cols: int, rowptr: np.ndarray, colindex: np.ndarray,
              value: np.ndarray, x: np.ndarray, y: np.ndarray) -> None:
    """One-shot host-buffer SpMV on the GPU (upload, analyse, multiply, download). In place on ``y``."""
    rc = _lib.lib().spmv_b200_host_spmv(float(alpha), float(beta), int(rows), int(cols), int(len(value)),
                                        _ptr(rowptr), _ptr(colindex), _ptr(value), _ptr(x), _ptr(y))
    check(rc, "host_spmv")


class HaloLoop:
    """x <- A*x on one row shard with the halo exchange fused into the SpMV kernels (``spmv_b200_halo_loop_*``).
    ``desc`` is a filled ``_lib.HaloLoopDesc``; ``run`` enqueues iterations on the current stream, ``sync`` waits for
    them and raises ``SpmvB200Error`` if a neighbour's flag timed out."""

    def __init__(self, desc: "_lib.HaloLoopDesc"):
        self._h = C.c_void_p()
        self._desc = desc  # keeps the plan / buffers named by the descriptor alive with the loop
        check(_lib.lib().spmv_b200_halo_loop_create(C.byref(self._h), C.byref(desc)), "halo_loop_create")

    def run(self, iterations: int, stream: Optional[int] = None) -> None:
        check(_lib.lib().spmv_b200_halo_loop_run(self._h, int(iterations),
                                                 _current_stream() if stream is None else stream), "halo_loop_run")

    def sync(self, stream: Optional[int] = None) -> None:
        check(_lib.lib().spmv_b200_halo_loop_sync(self._h, _current_stream() if stream is None else stream),
              "halo_loop_sync")

    def info(self) -> "_lib.HaloLoopInfo":
        info = _lib.HaloLoopInfo()
        check(_lib.lib().spmv_b200_halo_loop_get_info(self._h, C.byref(info)), "halo_loop_get_info")
        return info

    def destroy(self) -> None:
        if self._h:
            _lib.lib().spmv_b200_halo_loop_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def _fill_push(ps, push):
    """push = [(row_lo, row_hi, dst_address[, is_multicast_address])]"""
    ps.count = len(push)
    ps.multicast_mask = 0
    for j, entry in enumerate(push):
        lo, hi, dst = entry[:3]
        ps.row_lo[j], ps.row_hi[j], ps.dst[j] = int(lo), int(hi), int(dst)
        if len(entry) > 3 and entry[3]:
            ps.multicast_mask |= 1 << j
    return ps


def cache_revalidations() -> int:
    """Stateless calls whose cached plan did not match the matrix found at the same addresses (analysed again)."""
    return int(_lib.lib().spmv_b200_cache_revalidations())


def enable_peer_access(peer_device: int) -> None:
    check(_lib.lib().spmv_b200_enable_peer_access(int(peer_device)), "enable_peer_access")


class PeerBuffer:
    """Device buffer other GPUs (other processes) can store into. ``PeerBuffer.alloc`` owns cudaMalloc'ed memory on the
    current device and exposes its 64-byte CUDA IPC handle; ``PeerBuffer.open`` maps another process' buffer for the
    kernels of the *current* device (cudaIpcOpenMemHandle with lazy peer access, NVLink on an NVSwitch box)."""

    def __init__(self, address: int, nbytes: int, handle: bytes, owner: bool):
        self.address, self.nbytes, self.handle, self.owner = address, nbytes, handle, owner

    @classmethod
    def alloc(cls, nbytes: int) -> "PeerBuffer":
        out = C.c_void_p()
        handle = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        check(_lib.lib().spmv_b200_peer_alloc(C.byref(out), int(nbytes), handle), "peer_alloc")
        return cls(int(out.value), int(nbytes), handle.raw, True)

    @classmethod
    def open(cls, handle: bytes, nbytes: int) -> "PeerBuffer":
        out = C.c_void_p()
        check(_lib.lib().spmv_b200_peer_open(C.create_string_buffer(handle, _lib.IPC_HANDLE_BYTES), C.byref(out)),
              "peer_open")
        return cls(int(out.value), int(nbytes), handle, False)

    def tensor(self, dtype: str, count: int, offset_bytes: int = 0):
        """torch view of the buffer (no copy); only meaningful for buffers owned by this process."""
        import torch
        item = np.dtype(dtype).itemsize
        assert offset_bytes + count * item <= self.nbytes

        class _View:
            __cuda_array_interface__ = {"shape": (int(count),), "typestr": np.dtype(dtype).str,
                                        "data": (self.address + offset_bytes, False), "version": 2}
        t = torch.as_tensor(_View(), device="cuda")
        t._spmv_b200_keepalive = self
        return t

    def release(self) -> None:
        if self.address:
            fn = _lib.lib().spmv_b200_peer_free if self.owner else _lib.lib().spmv_b200_peer_close
            check(fn(self.address), "peer_free" if self.owner else "peer_close")
            self.address = 0


def coo_to_csr(rows: int, cols: int, d_row: Any, d_col: Any, d_val: Any) -> CsrDesc:
    """COO on the device (int32 row / col, fp64 values, any order, duplicates kept) -> CsrDesc with new device arrays,
    entries sorted by (row, col) like the reference's matrix_market::to_csr (cli/sparse_format.h:100-128)."""
    import torch
    _require_device(d_row, "row", "int32")
    _require_device(d_col, "col", "int32")
    _require_device(d_val, "val", "float64")
    nnz = int(d_val.numel())
    rowptr = torch.empty(rows + 1, dtype=torch.int32, device=d_val.device)
    col = torch.empty(max(nnz, 1), dtype=torch.int32, device=d_val.device)
    val = torch.empty(max(nnz, 1), dtype=torch.float64, device=d_val.device)
    check(_lib.lib().spmv_b200_coo_to_csr(int(rows), int(cols), nnz, _ptr(d_row), _ptr(d_col), _ptr(d_val),
                                          _ptr(rowptr), _ptr(col), _ptr(val), _current_stream()), "coo_to_csr")
    return CsrDesc(int(rows), int(cols), nnz, rowptr, col[:nnz], val[:nnz])


def shard_bounds(d_rowptr: Any, rows: int, nshards: int) -> np.ndarray:
    """nnz-balanced contiguous row shards: bounds[g] = lower_bound(rowptr, g*nnz/nshards) (int32, length nshards+1)."""
    _require_device(d_rowptr, "rowptr", "int32")
    out = np.zeros(nshards + 1, dtype=np.int32)
    check(_lib.lib().spmv_b200_shard_bounds(int(rows), -1, _ptr(d_rowptr), int(nshards), out.ctypes.data,
                                            _current_stream()), "shard_bounds")
    return out


def col_block_bitmap(d_colindex: Any, nnz: int, cols: int, block_shift: int) -> np.ndarray:
    """uint8 bitmap over blocks of 2^block_shift entries of x: 1 if the matrix references the block."""
    nblocks = (cols + (1 << block_shift) - 1) >> block_shift
    out = np.zeros(nblocks, dtype=np.uint8)
    if nnz > 0:
        _require_device(d_colindex, "colindex", "int32")
    check(_lib.lib().spmv_b200_col_block_bitmap(int(nnz), _ptr(d_colindex), int(cols), int(block_shift),
                                                out.ctypes.data, _current_stream()), "col_block_bitmap")
    return out
