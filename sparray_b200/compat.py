"""The kernel seam: the four names the reference's flat.py:7-10 imports from
``sparray/compat.py`` (selected at compat.py:104-120 between _merge.pyx and the
numpy fallbacks), here backed by sm_100a CUDA kernels through the C ABI.

Inputs may be CUDA int64 tensors (zero-copy) or host arrays (uploaded); results
are CUDA tensors.  Like the Cython kernels the inputs
must be int64 and C-contiguous -- anything else raises ValueError -- and are
assumed sorted and duplicate-free.  ``impl`` names the backend; there is no
``_impl`` fallback.
"""
import numpy as np
import torch

from . import _kernels as _k

__all__ = ["intersect1d_sorted", "union1d_sorted", "combine_ranges", "len_range"]

impl = "cuda-sm_100a"


def _dev(x, name):
    _k.require_cuda()
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.int64:
            raise ValueError("Buffer dtype mismatch for %s, expected 'int64' but got %s" % (name, x.dtype))
        if not x.is_contiguous():
            raise ValueError("%s is not C-contiguous" % name)
        return x.cuda()
    arr = np.asarray(x)
    if arr.size == 0:
        arr = arr.astype(np.int64)
    if arr.dtype != np.int64:
        raise ValueError("Buffer dtype mismatch for %s, expected 'int64' but got %s" % (name, arr.dtype))
    if arr.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % arr.ndim)
    if not arr.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return torch.from_numpy(arr).cuda()


def intersect1d_sorted(a, b, return_inds=False):
    """_merge.pyx:81-89: c, or (c, a_inds, b_inds)."""
    return _k.intersect1d_sorted(_dev(a, "a"), _dev(b, "b"), return_inds=bool(return_inds))


def union1d_sorted(a, b, return_masks=False):
    """_merge.pyx:92-102: c, or (c, lut, a_only, b_only)."""
    return _k.union1d_sorted(_dev(a, "a"), _dev(b, "b"), return_masks=bool(return_masks))


def combine_ranges(ranges, shape, result_size, inner=False):
    """_merge.pyx:7-12: flat indices of the box given by (start, stop, step) rows."""
    return _k.combine_ranges(ranges, shape, result_size, inner=inner)


def len_range(start, stop, step):
    """_merge.pyx:70-78."""
    return _k.len_range(start, stop, step)
