"""ctypes binding of ``libdmb200.so`` (C ABI: ``include/dmb200.h``).

The library is the product: if it cannot be loaded this module raises -- there is no
CPU fallback.  ``load_library(path)`` accepts an explicit path only so that the build
container's unit tests can inject the thread-by-thread CPU emulation of the kernels
(``tests/emu``); nothing in the package ever passes one.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

MAX_TILE_DIGITS = 6
MAX_OPS = 16
MAX_QUBITS = 32

OP_MATS, OP_CX, OP_CX_TSP, OP_DIAG2, OP_SWAP = 0, 1, 2, 3, 4
HAS_PA, HAS_PB = 1, 2

OP_DTYPE = np.dtype([("kind", "<i4"), ("flags", "<i4"), ("a", "i1"), ("b", "i1"), ("fd", "i1", (4,)),
                     ("reserved_", "i1", (2,)), ("pa", "<f8", (12,)), ("pb", "<f8", (12,)),
                     ("coef", "<f8", (16,))])
PASS_DTYPE = np.dtype([("n_tile_digits", "<i4"), ("n_ops", "<i4"), ("tile_digit", "<i4", (MAX_TILE_DIGITS,)),
                       ("ops", OP_DTYPE, (MAX_OPS,))])


QOP_DTYPE = np.dtype([("kind", "<i4"), ("flags", "<i4"), ("qa", "<i4"), ("qb", "<i4"), ("pa", "<f8", (12,)),
                      ("pb", "<f8", (12,)), ("coef", "<f8", (16,))])

ABI_VERSION = 5          # include/dmb200.h DMB_ABI_VERSION


class Stats(ctypes.Structure):
    _fields_ = [("tile_pass_launches", ctypes.c_uint64), ("other_launches", ctypes.c_uint64),
                ("fused_ops", ctypes.c_uint64), ("state_bytes_moved", ctypes.c_uint64),
                ("folded_swaps", ctypes.c_uint64), ("small_plan_launches", ctypes.c_uint64),
                ("chained_ops", ctypes.c_uint64)]


class DmbError(RuntimeError):
    pass


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_PKG_DIR, "libdmb200.so")

_c = ctypes
_vp, _i, _u64, _sz, _d = _c.c_void_p, _c.c_int, _c.c_uint64, _c.c_size_t, _c.c_double
_SIGNATURES = {
    "dmb_abi_version": (_i, []),
    "dmb_sizeof_op": (_sz, []),
    "dmb_sizeof_pass": (_sz, []),
    "dmb_last_error": (_c.c_char_p, []),
    "dmb_sizeof_qop": (_sz, []),
    "dmb_schedule": (_i, [_vp, _sz, _vp, _i, _i, _i, _i, _i, _i, _sz, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "dmb_create": (_i, [_i, _c.POINTER(_vp)]),
    "dmb_destroy": (_i, [_vp]),
    "dmb_set_stream": (_i, [_vp, _vp]),
    "dmb_sync": (_i, [_vp]),
    "dmb_get_stats": (_i, [_vp, _c.POINTER(Stats)]),
    "dmb_reset_stats": (_i, [_vp]),
    "dmb_set_tile_variant": (_i, [_vp, _i]),
    "dmb_init_product": (_i, [_vp, _vp, _i, _u64, _i, _vp, _vp, _vp, _d]),
    "dmb_apply_passes": (_i, [_vp, _vp, _i, _vp, _sz]),
    "dmb_apply_pass_remote": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i]),
    "dmb_apply_pass_remote_sel": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _i]),
    "dmb_ipc_export": (_i, [_vp, _vp, _vp, _vp]),
    "dmb_ipc_open": (_i, [_vp, _vp, _u64, _vp]),
    "dmb_ipc_close": (_i, [_vp, _vp]),
    "dmb_marginal": (_i, [_vp, _vp, _i, _u64, _i, _vp, _vp, _vp, _vp]),
    "dmb_fwht": (_i, [_vp, _vp, _i]),
    "dmb_contract_digit": (_i, [_vp, _vp, _vp, _u64, _u64, _vp]),
    "dmb_read_coeffs": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "dmb_to_matrix": (_i, [_vp, _vp, _i, _vp, _vp]),
    "dmb_dot": (_i, [_vp, _vp, _vp, _u64, _vp]),
    "dmb_chop": (_i, [_vp, _vp, _u64, _d]),
    "dmb_upload": (_i, [_vp, _vp, _vp, _u64, _u64]),
    "dmb_download": (_i, [_vp, _vp, _vp, _u64, _u64]),
    "dmb_download_async": (_i, [_vp, _vp, _vp, _u64, _u64]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


_LIBRARIES = {}


def load_library(path=None):
    """dlopen the CUDA library and type its entry points (once per path: engines share the
    handle, and with it the per-device ``dmb_ctx``).  Raises if it is missing."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _LIBRARIES:
        return _LIBRARIES[path]
    if not os.path.exists(path):
        raise DmbError("CUDA library %s not found -- build it with `python -c 'import __graft_entry__ as g; "
                       "g.build()'` (nvcc, sm_100a); this package has no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.dmb_abi_version() != ABI_VERSION:
        raise DmbError("ABI version mismatch")
    if lib.dmb_sizeof_op() != OP_DTYPE.itemsize or lib.dmb_sizeof_pass() != PASS_DTYPE.itemsize:
        raise DmbError("struct layout mismatch between capi.py and libdmb200.so")
    if lib.dmb_sizeof_qop() != QOP_DTYPE.itemsize:
        raise DmbError("struct layout mismatch between capi.py and libdmb200.so (dmb_qop)")
    _LIBRARIES[path] = lib
    return lib


def _ptr(arr):
    return arr.ctypes.data_as(ctypes.c_void_p)


SCHED_PROGRAM_ORDER, SCHED_TILE_SEARCH = 0, 1


def schedule(lib, qops, pos, n_digits, max_tile=MAX_TILE_DIGITS, max_ops=MAX_OPS, window=256, min_tail=0,
             final_moves=None, strategy=SCHED_PROGRAM_ORDER):
    """``dmb_schedule``: QOP_DTYPE array (program order, qubit ids) -> (PASS_DTYPE array, indices of
    the ops left unscheduled, leftover final moves or None).  ``pos`` (list) is advanced in place."""
    assert qops.dtype == QOP_DTYPE and qops.flags["C_CONTIGUOUS"]
    n_ops = len(qops)
    pos_arr = np.ascontiguousarray(pos, dtype=np.int32)
    out = np.empty(max(n_ops, 1), dtype=PASS_DTYPE)
    left = np.empty(max(n_ops, 1), dtype=np.int32)
    n_out, n_left = ctypes.c_size_t(), ctypes.c_size_t()
    if final_moves is not None:
        mv = np.ascontiguousarray(list(final_moves) or [(0, 0)], dtype=np.int32).reshape(-1, 2)
        n_mv = ctypes.c_int32(len(final_moves))
        mv_p, n_mv_p = _ptr(mv), ctypes.cast(ctypes.byref(n_mv), ctypes.c_void_p)
    else:
        mv_p = n_mv_p = None
    rc = lib.dmb_schedule(_ptr(qops), n_ops, _ptr(pos_arr), len(pos), int(n_digits), int(max_tile), int(max_ops),
                          int(window), int(strategy), int(min_tail), mv_p, n_mv_p, _ptr(out), len(out),
                          ctypes.cast(ctypes.byref(n_out), ctypes.c_void_p), _ptr(left),
                          ctypes.cast(ctypes.byref(n_left), ctypes.c_void_p))
    if rc != 0:
        raise DmbError(lib.dmb_last_error().decode("utf-8", "replace"))
    pos[:] = pos_arr.tolist()
    moves_left = None
    if final_moves is not None:
        moves_left = [(int(q), int(t)) for q, t in mv[:n_mv.value]]
    return out[:n_out.value], left[:n_left.value], moves_left


class Context:
    """One ``dmb_ctx``: a device binding + stream + counters."""

    def __init__(self, lib, device=0):
        self.lib = lib
        self._h = ctypes.c_void_p()
        self._check(lib.dmb_create(int(device), ctypes.byref(self._h)))

    def _check(self, rc):
        if rc != 0:
            raise DmbError(self.lib.dmb_last_error().decode("utf-8", "replace"))

    def close(self):
        if self._h:
            self.lib.dmb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ----------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self.lib.dmb_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def sync(self):
        self._check(self.lib.dmb_sync(self._h))

    def stats(self):
        s = Stats()
        self._check(self.lib.dmb_get_stats(self._h, ctypes.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in Stats._fields_}

    def reset_stats(self):
        self._check(self.lib.dmb_reset_stats(self._h))

    def set_tile_variant(self, v):
        self._check(self.lib.dmb_set_tile_variant(self._h, int(v)))

    # -- state -------------------------------------------------------------------------
    def init_product(self, state_ptr, n_bits, rank_bits, hi, lo, v, scale):
        hi = np.ascontiguousarray(hi, dtype=np.int32)
        lo = np.ascontiguousarray(lo, dtype=np.int32)
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(len(hi), 4)
        self._check(self.lib.dmb_init_product(self._h, state_ptr, int(n_bits), int(rank_bits), len(hi),
                                              _ptr(hi), _ptr(lo), _ptr(v), float(scale)))

    def apply_passes(self, state_ptr, n_bits, passes):
        if len(passes) == 0:
            return
        assert passes.dtype == PASS_DTYPE and passes.flags["C_CONTIGUOUS"]
        self._check(self.lib.dmb_apply_passes(self._h, state_ptr, int(n_bits), _ptr(passes), len(passes)))

    def apply_pass_remote(self, dst_ptr, n_bits, one_pass, src_tab, block_shift, push=False):
        """One tile pass fused with the exchange: pulls its input from peer buffers, or (push)
        runs in place and stores its output into peer buffers."""
        assert one_pass.dtype == PASS_DTYPE and len(one_pass) == 1
        tab = np.ascontiguousarray(src_tab, dtype=np.uint64)
        tab_bits = int(len(tab)).bit_length() - 1
        assert (1 << tab_bits) == len(tab)
        self._check(self.lib.dmb_apply_pass_remote(self._h, dst_ptr, int(n_bits), _ptr(one_pass), _ptr(tab),
                                                   tab_bits, int(block_shift), 1 if push else 0))

    def apply_pass_remote_sel(self, dst_ptr, n_bits, one_pass, src_tab, sel_bits, push=False):
        """The same with the table index gathered from the selected bits ``sel_bits`` of the element index (a global slot
        swapped with an arbitrary local slot)."""
        assert one_pass.dtype == PASS_DTYPE and len(one_pass) == 1
        tab = np.ascontiguousarray(src_tab, dtype=np.uint64)
        sel = np.ascontiguousarray(sel_bits, dtype=np.int32)
        assert len(tab) == 1 << len(sel)
        self._check(self.lib.dmb_apply_pass_remote_sel(self._h, dst_ptr, int(n_bits), _ptr(one_pass), _ptr(tab),
                                                       len(sel), _ptr(sel), 1 if push else 0))

    def ipc_export(self, dev_ptr):
        handle = (ctypes.c_ubyte * 64)()
        off = ctypes.c_uint64()
        self._check(self.lib.dmb_ipc_export(self._h, ctypes.c_void_p(dev_ptr), handle, ctypes.byref(off)))
        return bytes(handle), int(off.value)

    def ipc_open(self, handle, offset):
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        out = ctypes.c_void_p()
        self._check(self.lib.dmb_ipc_open(self._h, buf, int(offset), ctypes.byref(out)))
        return int(out.value)

    def ipc_close(self, handle):
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self.lib.dmb_ipc_close(self._h, buf))

    def marginal(self, state_ptr, n_bits, rank_bits, hi, lo, wt, out_ptr):
        hi = np.ascontiguousarray(hi, dtype=np.int32)
        lo = np.ascontiguousarray(lo, dtype=np.int32)
        wt = np.ascontiguousarray(wt, dtype=np.float64).reshape(len(hi), 2, 4)
        self._check(self.lib.dmb_marginal(self._h, state_ptr, int(n_bits), int(rank_bits), len(hi),
                                          _ptr(hi), _ptr(lo), _ptr(wt), out_ptr))

    def fwht(self, vec_ptr, n_qubits):
        self._check(self.lib.dmb_fwht(self._h, vec_ptr, int(n_qubits)))

    def contract_digit(self, in_ptr, out_ptr, H, L, nv):
        nv = np.ascontiguousarray(nv, dtype=np.float64)
        self._check(self.lib.dmb_contract_digit(self._h, in_ptr, out_ptr, int(H), int(L), _ptr(nv)))

    READ_CHUNK = 1 << 16         # the library's gather scratch holds 4^8 coefficients per call

    def read_coeffs(self, state_ptr, idx):
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        out = np.empty(len(idx), dtype=np.float64)
        for lo in range(0, len(idx), self.READ_CHUNK):
            part, dst = idx[lo:lo + self.READ_CHUNK], out[lo:lo + self.READ_CHUNK]
            self._check(self.lib.dmb_read_coeffs(self._h, state_ptr, _ptr(part), len(part), _ptr(dst)))
        return out

    def to_matrix(self, state_ptr, n_qubits, work_ptr, out_ptr):
        self._check(self.lib.dmb_to_matrix(self._h, state_ptr, int(n_qubits), work_ptr, out_ptr))

    def dot(self, a_ptr, b_ptr, count):
        out = ctypes.c_double()
        self._check(self.lib.dmb_dot(self._h, a_ptr, b_ptr, int(count), ctypes.byref(out)))
        return out.value

    def chop(self, state_ptr, count, thr):
        self._check(self.lib.dmb_chop(self._h, state_ptr, int(count), float(thr)))

    def upload(self, state_ptr, host, offset=0):
        host = np.ascontiguousarray(host, dtype=np.float64)
        self._check(self.lib.dmb_upload(self._h, state_ptr, _ptr(host), int(offset), host.size))
        self.sync()      # the host array may be a temporary

    def download_async(self, state_ptr, host, offset=0):
        """D2H copy without waiting: ``host`` (a C-contiguous float64 array in pinned memory) is valid after ``sync()``."""
        assert host.dtype == np.float64 and host.flags["C_CONTIGUOUS"]
        self._check(self.lib.dmb_download_async(self._h, state_ptr, _ptr(host), int(offset), host.size))

    def download(self, state_ptr, host, offset=0):
        assert host.dtype == np.float64 and host.flags["C_CONTIGUOUS"]
        self._check(self.lib.dmb_download(self._h, state_ptr, _ptr(host), int(offset), host.size))
