"""oracle/flat_oracle.py -- TEST INFRASTRUCTURE ONLY. Not part of the product.

CPU restatement (numpy + the C port in merge_oracle.c) of the reference's
FlatSparray hot path -- perimosocordiae/sparray, sparray/flat.py, base.py,
compat.py, _merge.pyx.  Arrays are passed as plain ``(indices, data, shape)``
triples of numpy arrays; nothing here touches torch or CUDA.

Only ``tests/``, ``bench.py``'s cpu_baseline / ``--impl reference`` leg and
``__graft_entry__.smoke()`` may import this module.  The shipped package
(``sparray_b200``) never does and has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this module against
(1) the reference's own known-answer vectors (sparray/tests/test_compat.py:10-53),
(2) fixtures produced by running the reference itself (both its Cython and its
numpy kernel paths) through ``tests/golden/make_golden.py`` and (3) in the build
container, live against the reference for randomised inputs.

Two deliberate divergences from the reference, both documented in DESIGN.md:
  * ``transpose`` follows numpy semantics when the permuted shape equals the
    original shape; the reference returns ``self`` there (flat.py:106-107,
    SURVEY.md F5).  ``transpose(..., reference_bug=True)`` reproduces the bug.
  * ``dot`` with a dense vector is the O(nnz) restatement
    ``bincount(idx // ncols, data * x[idx % ncols])`` instead of the reference's
    ``toarray().dot`` (flat.py:568), which cannot run at config sizes (F7); both
    are checked against each other at small sizes.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    """The C port (merge_oracle.c), built on demand with gcc."""
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libmerge_oracle.so")
        src = os.path.join(_HERE, "merge_oracle.c")
        if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = ctypes.CDLL(so)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.orc_len_range.restype = i64
        lib.orc_len_range.argtypes = [i64, i64, i64]
        lib.orc_merge_unique.restype = i64
        lib.orc_merge_unique.argtypes = [p, i64, p, i64, p, p, p, p]
        lib.orc_merge_intersect.restype = i64
        lib.orc_merge_intersect.argtypes = [p, i64, p, i64, p, p, p]
        lib.orc_combine_ranges.restype = None
        lib.orc_combine_ranges.argtypes = [p, p, i64, p, i64, ctypes.c_int]
        _LIB = lib
    return _LIB


def _i64(x):
    return np.ascontiguousarray(x, dtype=np.int64)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
# L1/L2: the four kernel names of the compat seam (compat.py:104-120)
# ---------------------------------------------------------------------------

def len_range(start, stop, step):
    """_merge.pyx:70-78 / compat.py:78-86."""
    return int(_lib().orc_len_range(int(start), int(stop), int(step)))


def union1d_sorted(a, b, return_masks=False):
    """_merge.pyx:92-102 (wrapper) over merge_unique :107-140."""
    a, b = _i64(a), _i64(b)
    c = np.empty(len(a) + len(b), dtype=np.int64)
    lut = np.empty(len(a) + len(b), dtype=np.uint8)
    a_only = np.zeros(len(a), dtype=np.uint8)
    b_only = np.zeros(len(b), dtype=np.uint8)
    n = _lib().orc_merge_unique(_ptr(a), len(a), _ptr(b), len(b), _ptr(c),
                                _ptr(lut), _ptr(a_only), _ptr(b_only))
    c = c[:n]
    if not return_masks:
        return c
    return c, lut[:n], a_only.view(bool), b_only.view(bool)


def intersect1d_sorted(a, b, return_inds=False):
    """_merge.pyx:81-89 (wrapper) over merge_intersect :145-157."""
    a, b = _i64(a), _i64(b)
    m = min(len(a), len(b))
    c = np.empty(m, dtype=np.int64)
    if not return_inds:
        n = _lib().orc_merge_intersect(_ptr(a), len(a), _ptr(b), len(b), _ptr(c), None, None)
        return c[:n]
    ai = np.empty(m, dtype=np.int64)
    bi = np.empty(m, dtype=np.int64)
    n = _lib().orc_merge_intersect(_ptr(a), len(a), _ptr(b), len(b), _ptr(c), _ptr(ai), _ptr(bi))
    return c[:n], ai[:n], bi[:n]


def combine_ranges(ranges, shape, result_size, inner=False):
    """_merge.pyx:7-12 (wrapper) over combine_ranges_fast :18-66."""
    ranges = _i64(ranges)
    shape = _i64(shape)
    out = np.zeros(int(result_size), dtype=np.int64)
    _lib().orc_combine_ranges(_ptr(ranges), _ptr(shape), len(shape), _ptr(out),
                              int(result_size), int(bool(inner)))
    return out


# ---------------------------------------------------------------------------
# L3: FlatSparray semantics on (indices, data, shape) triples
# ---------------------------------------------------------------------------

def canonicalize(indices, data):
    """flat.py:32-35 -- sort, merge duplicates by summation (in float64, cast
    back), keep explicit zeros."""
    indices = np.asarray(indices, dtype=np.int64).ravel()
    data = np.asarray(data).ravel()
    uniq, inv = np.unique(indices, return_inverse=True)
    summed = np.bincount(inv, weights=data, minlength=len(uniq))
    return uniq, summed.astype(data.dtype, copy=False)


def pairwise_union(a_idx, a_data, b_idx, b_data, ufunc, dtype=None):
    """flat.py:409-422 (_pairwise_sparray): f(a or 0, b or 0) on the union."""
    if dtype is None:
        dtype = np.promote_types(a_data.dtype, b_data.dtype)
    idx, lut, a_only, b_only = union1d_sorted(a_idx, b_idx, return_masks=True)
    out = np.empty(len(idx), dtype=dtype)
    out[lut == 0] = ufunc(a_data[a_only], 0)
    out[lut == 1] = ufunc(0, b_data[b_only])
    out[lut == 2] = ufunc(a_data[~a_only], b_data[~b_only])
    return idx, out


def pairwise_intersect(a_idx, a_data, b_idx, b_data, ufunc):
    """flat.py:424-434 (_pairwise_sparray_fixed_zero): f(a, b) on the
    intersection (valid when f(x, 0) == 0)."""
    idx, ai, bi = intersect1d_sorted(a_idx, b_idx, return_inds=True)
    return idx, ufunc(a_data[ai], b_data[bi])


def transpose(indices, data, shape, axes=None, reference_bug=False):
    """flat.py:97-112 + ctor :32-35.  Returns (indices, data, new_shape)."""
    ndim = len(shape)
    if ndim < 2:
        return indices, data, tuple(shape)
    if axes is None:
        axes = tuple(range(ndim - 1, -1, -1))
    axes = tuple(int(a) for a in axes)
    new_shape = tuple(shape[a] for a in axes)
    if reference_bug and tuple(shape) == new_shape:   # flat.py:106-107 (F5)
        return indices, data, tuple(shape)
    if axes == tuple(range(ndim)):
        return indices, data, tuple(shape)
    if len(indices) == 0:
        return indices, data, new_shape
    old_multi = np.unravel_index(indices, shape)
    new_flat = np.ravel_multi_index(tuple(old_multi[a] for a in axes), new_shape)
    # a permutation has no duplicates, so canonicalising == stable sort
    order = np.argsort(new_flat, kind="stable")
    return new_flat[order], data[order], new_shape


def sum_axis(indices, data, shape, axis, dtype=None):
    """flat.py:607-626.  Returns a scalar (axis None / 1-D) or
    (indices, float64 data, new_shape) -- bincount makes the output float64
    whatever the input dtype."""
    if dtype is None:
        dtype = data.dtype
    if axis is None:
        return data.sum(dtype=dtype)
    axis = int(axis)
    new_shape = tuple(shape[:axis]) + tuple(shape[axis + 1:])
    if not new_shape:
        return data.sum(dtype=dtype)
    if len(indices) == 0:
        return (np.zeros(0, np.int64), np.zeros(0, np.float64), new_shape)
    multi = np.unravel_index(indices, shape)
    multi = multi[:axis] + multi[axis + 1:]
    flat = np.ravel_multi_index(multi, new_shape)
    new_idx, inv = np.unique(flat, return_inverse=True)
    new_data = np.bincount(inv, data.astype(dtype, copy=False), minlength=len(new_idx))
    return new_idx, new_data, new_shape


def mean_axis(indices, data, shape, axis):
    """base.py:84-100 restricted to real dtypes."""
    out = sum_axis(indices, data, shape, axis, dtype=np.float64)
    n = np.prod(shape) if axis is None else shape[axis]
    if not isinstance(out, tuple):
        return out / n if n != 1 else out
    idx, dat, shp = out
    if n != 1:
        dat = dat * (1.0 / n)       # flat.py:698-704 (__itruediv__ by scalar)
    return idx, dat, shp


def dot_dense_vector(indices, data, shape, x):
    """O(nnz) restatement of ``a.dot(x)`` for a 2-D (or N-D, last axis
    contracted) sparse ``a`` and a dense vector ``x`` -- SURVEY.md section 8(c);
    equals the reference's dense route flat.py:568 and its sparse route
    :545-561 at sizes where those can run."""
    ncols = shape[-1]
    nrows = int(np.prod(shape[:-1]))
    rows = indices // ncols
    cols = indices - rows * ncols
    dt = np.result_type(data.dtype, x.dtype)
    y = np.bincount(rows, weights=(data.astype(dt) * x[cols].astype(dt)).astype(np.float64),
                    minlength=nrows)
    return y.astype(dt).reshape(tuple(shape[:-1]))


def dot_1d_dense(indices, data, x):
    """flat.py:564-566: 1-D sparse . 1-D dense."""
    return x[indices].dot(data)


def getitem_flatidx(indices, data, flat_idx, is_sorted=True):
    """flat.py:383-393 (_getitem_flatidx): result index = position in query."""
    flat_idx = np.asarray(flat_idx, dtype=np.int64)
    if not is_sorted:
        order = np.argsort(flat_idx, kind="mergesort")
        flat_idx = flat_idx[order]
    _, a_pos, q_pos = intersect1d_sorted(indices, flat_idx, return_inds=True)
    if not is_sorted:
        q_pos = order[q_pos]
    return q_pos, data[a_pos]


def indices_to_ranges(index, shape):
    """flat.py:366-381 (_indices_to_ranges) for an all int/slice index tuple
    already padded to ndim."""
    ranges = np.zeros((len(index), 3), dtype=np.int64)
    new_shape = []
    for ax, ix in enumerate(index):
        if isinstance(ix, slice):
            row = ix.indices(shape[ax])
            ranges[ax] = row
            new_shape.append(len_range(*row))
        else:
            ranges[ax] = (ix, ix + 1, 1)
    return ranges, new_shape


def getitem_basic(indices, data, shape, index):
    """flat.py:270-274: ints + slices -> materialise the box, intersect."""
    ranges, new_shape = indices_to_ranges(index, shape)
    box = combine_ranges(ranges, shape, int(np.prod(new_shape)), inner=False)
    q_pos, vals = getitem_flatidx(indices, data, box)
    return q_pos, vals, tuple(new_shape)


def getitem_scalar(indices, data, shape, index):
    """flat.py:262-267: all-integer index -> stored value or python 0."""
    flat = np.ravel_multi_index(index, shape)
    i = np.searchsorted(indices, flat)
    if i >= len(indices) or indices[i] != flat:
        return 0
    return data[i]


def toarray(indices, data, shape):
    """flat.py:74-77."""
    out = np.zeros(shape, dtype=data.dtype)
    out.flat[indices] = data
    return out


def from_ndarray(arr):
    """flat.py:48-54."""
    arr = np.asarray(arr)
    mask = arr.flat != 0
    idx, = np.nonzero(mask)
    return idx.astype(np.int64), arr.flat[mask], arr.shape


# ---------------------------------------------------------------------------
# Seeded synthetic operands shared by tests and bench
# ---------------------------------------------------------------------------

def random_operand(shape, density, dtype, seed):
    """indices = unique(rng.integers(0, prod(shape), round(density*prod))),
    values = standard_normal -- the generator SURVEY.md section 8(d) fixes."""
    rng = np.random.default_rng(seed)
    size = int(np.prod([int(s) for s in shape], dtype=object))
    n = int(round(density * size))
    idx = np.unique(rng.integers(0, size, size=n, dtype=np.int64))
    val = rng.standard_normal(len(idx)).astype(dtype)
    return idx, val


# ---------------------------------------------------------------------------
# A thin object wrapper so tests can evaluate the same expressions on the
# oracle, on the reference and on the product ("a + b", "a[1:, 2]", "a.sin()").
# ---------------------------------------------------------------------------

_FIXED_ZERO_UFUNCS = ("sin tan arcsin arctan sinh tanh arcsinh arctanh rint sign expm1 log1p "
                      "deg2rad rad2deg floor ceil trunc sqrt").split()   # compat.py:96-102


class OracleSparray:
    """CPU mirror of the FlatSparray surface that lies on the hot path.  Only
    sparse(op)sparse, scalar maps, reductions, transpose/reshape, dot with a
    vector and __getitem__ are provided -- densifying branches are not."""

    def __init__(self, indices, data, shape=None, is_canonical=False):
        indices = np.asarray(indices, dtype=np.int64).ravel()
        data = np.asarray(data).ravel()
        assert len(indices) == len(data)
        if not is_canonical:
            indices, data = canonicalize(indices, data)
        self.shape = (int(indices[-1]) + 1,) if shape is None else tuple(int(s) for s in shape)
        self.indices, self.data = indices, data

    dtype = property(lambda self: self.data.dtype)
    nnz = property(lambda self: len(self.indices))
    ndim = property(lambda self: len(self.shape))
    T = property(lambda self: self.transpose())

    @staticmethod
    def from_ndarray(arr):
        return OracleSparray(*from_ndarray(arr), is_canonical=True)

    def toarray(self):
        return toarray(self.indices, self.data, self.shape)

    def _with_data(self, data):                      # flat.py:580-581
        return OracleSparray(self.indices.copy(), data, self.shape, is_canonical=True)

    def _union(self, other, ufunc, dtype=None):
        assert self.shape == other.shape
        idx, dat = pairwise_union(self.indices, self.data, other.indices, other.data, ufunc, dtype)
        return OracleSparray(idx, dat, self.shape, is_canonical=True)

    def _cmp(self, other, ufunc):                    # flat.py:474-488
        if np.isscalar(other):
            assert not ufunc(0, other), "densifying comparison is off the hot path"
            return self._with_data(ufunc(self.data, other))
        return self._union(other, ufunc, dtype=bool)

    def __lt__(self, o): return self._cmp(o, np.less)
    def __le__(self, o): return self._cmp(o, np.less_equal)
    def __gt__(self, o): return self._cmp(o, np.greater)
    def __ge__(self, o): return self._cmp(o, np.greater_equal)
    def __eq__(self, o): return self._cmp(o, np.equal)
    def __ne__(self, o): return self._cmp(o, np.not_equal)
    __hash__ = None

    def __add__(self, o):                            # flat.py:490-505
        if np.isscalar(o):
            assert o == 0
            return self.copy()
        return self._union(o, np.add)

    def __sub__(self, o): return self.__add__(-o)    # base.py:45-46
    def __neg__(self): return self._with_data(-self.data)
    def __abs__(self): return self._with_data(abs(self.data))

    def __mul__(self, o):                            # flat.py:507-518
        if np.isscalar(o):
            return self._with_data(self.data * o)
        assert self.shape == o.shape
        idx, dat = pairwise_intersect(self.indices, self.data, o.indices, o.data, np.multiply)
        return OracleSparray(idx, dat, self.shape, is_canonical=True)
    __rmul__ = __mul__

    def __truediv__(self, o):                        # flat.py:528-529
        return self.__mul__(1. / o)

    def __floordiv__(self, o):                       # flat.py:531-533
        return self._with_data(np.floor_divide(self.data, o))

    def __pow__(self, e):                            # flat.py:570-578
        assert e > 0
        return self._with_data(self.data ** e)

    def minimum(self, o):                            # flat.py:583-593
        if np.isscalar(o):
            assert o >= 0
            return self._with_data(np.minimum(self.data, o))
        return self._union(o, np.minimum)

    def maximum(self, o):                            # flat.py:595-605
        if np.isscalar(o):
            assert o <= 0
            return self._with_data(np.maximum(self.data, o))
        return self._union(o, np.maximum)

    def astype(self, t): return self._with_data(self.data.astype(t))
    def copy(self): return self._with_data(self.data.copy())

    def __getattr__(self, name):                     # flat.py:722-739
        if name in _FIXED_ZERO_UFUNCS:
            return lambda: self._with_data(getattr(np, name)(self.data))
        raise AttributeError(name)

    def transpose(self, *axes):
        if len(axes) == 1 and not np.isscalar(axes[0]):
            axes = tuple(axes[0])
        i, d, s = transpose(self.indices, self.data, self.shape, axes or None)
        return OracleSparray(i, d, s, is_canonical=True)

    def reshape(self, new_shape):                    # flat.py:173-183
        new_shape = list(new_shape)
        if -1 in new_shape:
            k = new_shape.index(-1)
            new_shape[k] = int(np.prod(self.shape)) // -int(np.prod(new_shape))
        return OracleSparray(self.indices, self.data, new_shape, is_canonical=True)

    def sum(self, axis=None, dtype=None):
        out = sum_axis(self.indices, self.data, self.shape, axis, dtype)
        return OracleSparray(*out, is_canonical=True) if isinstance(out, tuple) else out

    def mean(self, axis=None):
        out = mean_axis(self.indices, self.data, self.shape, axis)
        return OracleSparray(*out, is_canonical=True) if isinstance(out, tuple) else out

    def dot(self, x):
        x = np.asarray(x)
        if self.ndim == 1:
            return dot_1d_dense(self.indices, self.data, x)
        return dot_dense_vector(self.indices, self.data, self.shape, x)

    def __getitem__(self, index):
        """numpy semantics on the stored elements; equals flat.py:254-325 where
        the reference is defined (ascending, duplicate-free fancy queries)."""
        if not isinstance(index, tuple):
            index = (index,)
        grid = np.arange(int(np.prod(self.shape)), dtype=np.int64).reshape(self.shape)[index]
        if np.ndim(grid) == 0:
            return getitem_scalar(self.indices, self.data, (int(np.prod(self.shape)),), (int(grid),))
        q = np.asarray(grid).ravel()
        pos, vals = getitem_flatidx(self.indices, self.data, q, is_sorted=bool(np.all(np.diff(q) > 0)))
        order = np.argsort(pos, kind="stable")
        return OracleSparray(pos[order], vals[order], np.shape(grid), is_canonical=True)
