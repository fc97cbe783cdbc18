"""ctypes binding of libsparray_b200.so (include/sparray_b200.h).

The shared object is hand-written sm_100a CUDA behind a C ABI; this module only
declares prototypes, turns status codes into exceptions and hands out the
per-stream workspace.  There is NO CPU fallback: if the library cannot be
loaded, or no CUDA device is present when a kernel is requested, it raises.
"""
import ctypes
import os
import threading

import torch

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsparray_b200.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_p = ctypes.c_void_p
c_f64 = ctypes.c_double


class SparrayB200Error(RuntimeError):
    pass


def _load():
    override = os.environ.get("SPB_LIB_PATH")       # kernel A/B experiments (profiles/build_variant.py): a prebuilt variant
    if override:
        return ctypes.CDLL(override)
    if _build.needs_build():
        # normally built in-tree by __graft_entry__.build(); rebuild here when the sources changed
        try:
            _build.build()
        except Exception as exc:  # pragma: no cover
            raise ImportError("libsparray_b200.so is missing/stale and could not be built with nvcc: %s" % exc)
    return ctypes.CDLL(LIB_PATH)


lib = _load()

# prototypes: name -> (restype, argtypes).  Every symbol include/sparray_b200.h declares.
PROTOTYPES = {
    "spb_version": (ctypes.c_char_p, []),
    "spb_error_string": (ctypes.c_char_p, [c_int]),
    "spb_device_count": (c_int, []),
    "spb_launch_count": (ctypes.c_ulonglong, []),
    "spb_debug_fastdiv": (ctypes.c_uint64, [ctypes.c_uint64, ctypes.c_uint64]),
    "spb_workspace_create": (c_int, [ctypes.POINTER(c_p)]),
    "spb_workspace_destroy": (c_int, [c_p]),
    "spb_union1d_sorted": (c_int, [c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "spb_intersect1d_sorted": (c_int, [c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p]),
    "spb_combine_ranges": (c_int, [c_p, c_p, c_int, c_p, c_i64, c_int, c_p]),
    "spb_len_range": (c_i64, [c_i64, c_i64, c_i64]),
    "spb_host_union1d_sorted": (c_int, [c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_host_intersect1d_sorted": (c_int, [c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p]),
    "spb_host_combine_ranges": (c_int, [c_p, c_p, c_int, c_p, c_i64, c_int]),
    "spb_union_binop": (c_int, [c_int, c_int, c_p, c_p, c_i64, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_intersect_binop": (c_int, [c_int, c_int, c_p, c_p, c_i64, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_unary_map": (c_int, [c_int, c_int, c_p, c_p, c_i64, c_f64, c_i64, c_p]),
    "spb_cast": (c_int, [c_int, c_int, c_p, c_p, c_i64, c_p]),
    "spb_binary_map": (c_int, [c_int, c_int, c_p, c_p, c_p, c_i64, c_p]),
    "spb_broadcast_expand": (c_int, [c_int, c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_p, c_p]),
    "spb_scatter_dense": (c_int, [c_int, c_p, c_p, c_i64, c_p, c_p]),
    "spb_gather": (c_int, [c_int, c_p, c_p, c_i64, c_p, c_p]),
    "spb_lower_bound": (c_int, [c_p, c_i64, c_p, c_i64, c_p, c_p, c_p]),
    "spb_select_nonzero": (c_int, [c_int, c_p, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_sum_all": (c_int, [c_int, c_p, c_i64, c_p, c_p, c_p]),
    "spb_select_flagged": (c_int, [c_int, c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_min_max": (c_int, [c_int, c_p, c_i64, c_int, c_p, c_p, c_p]),
    "spb_reduce_by_key": (c_int, [c_int, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_spmv": (c_int, [c_int, c_p, c_p, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_p]),
    "spb_axis_sum_dense": (c_int, [c_int, c_p, c_p, c_i64, c_p, c_int, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_axis_sum_window": (c_int, [c_int, c_p, c_p, c_i64, c_p, c_int, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_axis_accumulate": (c_int, [c_int, c_p, c_p, c_i64, c_int, c_p, c_p, c_i64, c_p, c_p, c_p]),
    "spb_emit_touched": (c_int, [c_p, c_p, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_axis_sum_slab_ok": (c_int, [c_i64, c_i64, c_i64, c_i64]),
    "spb_complex_split": (c_int, [c_int, c_p, c_i64, c_p, c_p, c_p]),
    "spb_complex_join": (c_int, [c_int, c_p, c_p, c_i64, c_p, c_p]),
    "spb_axis_sum_slab": (c_int, [c_int, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_getitem_box": (c_int, [c_int, c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "spb_outer_sum": (c_int, [c_p, c_int, c_p, c_p, c_i64, c_p]),
    "spb_lookup_compact": (c_int, [c_int, c_p, c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "spb_spgemm_rowlen": (c_int, [c_p, c_i64, c_i64, c_p, c_p, c_p]),
    "spb_exclusive_scan_i64": (c_int, [c_p, c_i64, c_p, c_p, c_p, c_p]),
    "spb_spgemm_expand": (c_int, [c_int, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p]),
    "spb_rekey_sort": (c_int, [c_int, c_p, c_p, c_i64, c_int, c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_p, c_p, c_p]),
}


def _declare(extra=None):
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


_declare()


def check(rc):
    if rc != 0:
        msg = lib.spb_error_string(int(rc))
        raise SparrayB200Error("libsparray_b200: %s (code %d)" % (msg.decode() if msg else "?", rc))


_cuda_ok = False


def require_cuda():
    global _cuda_ok
    if _cuda_ok:
        return
    if not torch.cuda.is_available():
        raise SparrayB200Error("sparray_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _cuda_ok = True


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _raw_stream():
    # torch.cuda.current_stream() costs ~10 us of Python per call; this is the same handle
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def stream():
    return ctypes.c_void_p(_raw_stream())


_ws_lock = threading.Lock()
_workspaces = {}


def workspace():
    """One scratch workspace per (device, stream)."""
    key = (torch._C._cuda_getDevice(), _raw_stream())
    ws = _workspaces.get(key)
    if ws is None:
        with _ws_lock:
            ws = _workspaces.get(key)
            if ws is None:
                handle = c_p()
                check(lib.spb_workspace_create(ctypes.byref(handle)))
                ws = _workspaces[key] = handle
    return ws
