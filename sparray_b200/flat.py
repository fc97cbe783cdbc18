"""FlatSparray on a B200: the host-side mirror of the reference's
``sparray/flat.py``.  Same class name, attributes (``indices``, ``data``,
``shape``), method names, argument meaning and error behaviour -- but
``indices`` (sorted unique int64 C-order flat indices) and ``data`` are
device-resident ``torch`` tensors and every piece of arithmetic is a
hand-written sm_100a CUDA kernel reached through the C ABI
(include/sparray_b200.h).  There is no CPU fallback: without the CUDA library
and a GPU, construction raises.

Deliberate divergences from the reference:
  * ``transpose`` is correct when the permuted shape equals the shape
    (flat.py:106-107 returns ``self`` there -- SURVEY.md F5);
  * ``swapaxes`` exists (absent upstream, F2);
  * ``dot`` with a dense vector is an O(nnz) SpMV (flat.py:568 densifies, F7);
  * ``__getitem__`` returns sorted (canonical) results for unsorted fancy
    queries and numpy semantics for repeated ones;
  * broadcasting between sparse operands stays sparse (index-expansion kernel + sort)
    instead of densifying (flat.py:470-472);
  * results whose size depends on the data are lazily counted: the element count is read back from
    the device when ``indices`` / ``data`` / ``nnz`` are first looked at, not when the op returns
    (invisible to callers; it keeps chains such as ``(a * b).sum(axis=2)`` free of host waits);
  * branches whose RESULT is dense by nature in the reference (sparse + nonzero scalar, sparse / sparse,
    sparse ** 0, comparisons that are true at zero, ...) return the same dense ndarray with the same
    warning; the dense array is built and combined on the device (there is still no CPU arithmetic);
    fancy assignment raises like the reference.
"""
import numbers
import warnings

import numpy as np
import torch

from . import _dtypes as dt
from . import _indexing as ix
from . import _kernels as _k
from .base import _BaseSparray

try:  # only for the warning category and spmatrix interop; never for arithmetic
    import scipy.sparse as ss
    _EfficiencyWarning = ss.SparseEfficiencyWarning
except Exception:  # pragma: no cover
    ss = None
    _EfficiencyWarning = UserWarning

_FIXED_POINT_AT_ZERO = ("sin", "tan", "arcsin", "arctan", "sinh", "tanh", "arcsinh", "arctanh", "rint", "sign",
                        "expm1", "log1p", "deg2rad", "rad2deg", "floor", "ceil", "trunc", "sqrt")   # compat.py:96-102


def _is_complex_data(x):
    if isinstance(x, torch.Tensor):
        return x.is_complex()
    if isinstance(x, np.ndarray):
        return x.dtype.kind == "c"
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], (complex, np.complexfloating)):
        return True
    return False


def _device():
    _k.require_cuda()
    return torch.device("cuda", torch._C._cuda_getDevice())


def _to_device(x, dtype=None):
    """numpy / list / torch (any device) -> contiguous 1-D tensor on the current CUDA device."""
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 1 and x.is_contiguous() and (dtype is None or x.dtype == dtype):
        return x                                   # the common internal case: already a device vector
    dev = _device()
    if isinstance(x, torch.Tensor):
        t = x.to(dev)
    else:
        arr = np.asarray(x)
        if arr.dtype == object or arr.dtype.kind in "US":
            raise TypeError("unsupported data type %s" % arr.dtype)
        if arr.dtype.kind == "f" and arr.size == 0 and dtype is None:
            arr = arr.astype(np.float64)
        dt.torch_dtype(arr.dtype)      # raises TypeError for dtypes with no device form
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    if dtype is not None and t.dtype != dtype:
        t = _k.cast(t.contiguous().reshape(-1), dtype) if t.numel() else t.to(dtype)
    return t.contiguous().reshape(-1)


def _to_device_nd(x):
    """numpy / torch array of any rank -> contiguous tensor on the current CUDA device."""
    if isinstance(x, torch.Tensor):
        return x.to(_device()).contiguous()
    arr = np.asarray(x)
    dt.torch_dtype(arr.dtype)
    return torch.from_numpy(np.ascontiguousarray(arr)).to(_device())


def _is_scalar(x):
    return np.isscalar(x) or (isinstance(x, (np.ndarray, torch.Tensor)) and x.ndim == 0)


class FlatSparray(_BaseSparray):
    """Simple sparse ndarray-like, similar to scipy.sparse matrices.
    Defined by three member variables (flat.py:19-25):
      self.data : device tensor of nonzero values (may include zeros)
      self.indices : sorted int64 device tensor of nonzero flat indices
      self.shape : tuple of integers, ala ndarray shape
    """

    def __new__(cls, indices=None, data=None, shape=None, is_canonical=False):
        # complex values: the subclass that composes every operation from the real kernels (sparray_b200/_complex.py)
        if cls is FlatSparray and data is not None and _is_complex_data(data):
            from ._complex import ComplexFlatSparray
            return object.__new__(ComplexFlatSparray)
        return object.__new__(cls)

    def __init__(self, indices, data, shape=None, is_canonical=False):    # flat.py:27-42
        indices = _to_device(indices, torch.int64)
        data = _to_device(data)
        assert len(indices) == len(data), "# inds (%d) != # data (%d)" % (len(indices), len(data))
        if not is_canonical:
            # sort and sum duplicates, but allow explicit zeros (flat.py:32-35)
            indices, data = _k.canonicalize(indices, data)
        if shape is None:
            self.shape = (int(indices[-1].item()) + 1,)
        else:
            self.shape = tuple(int(s) for s in shape)
            assert np.prod(self.shape, dtype=object) >= len(data)
        self._pending = None
        self._indices = indices
        self._data = data

    # -- lazily counted results ---------------------------------------------------------------
    # A kernel with a data-dependent output size writes worst-case-sized buffers and leaves the
    # element count on the device.  Reading the count is the op's one host synchronisation, so it
    # is put off until somebody looks at `indices` / `data` / `nnz`: a consumer that can take the
    # count from the device (sum over an axis through the dense accumulator) chains on the stream
    # without the host ever waiting.
    @classmethod
    def _from_kernel(cls, idx, val, cnt, shape):
        """(indices, values, device count or None) of a kernel wrapper -> canonical FlatSparray."""
        if cnt is None:
            return cls(idx, val, shape=shape, is_canonical=True)
        self = cls.__new__(cls)
        self.shape = shape if type(shape) is tuple else tuple(int(s) for s in shape)   # callers pass self.shape slices
        self._pending = (idx, val, cnt)
        self._indices = self._data = None
        return self

    def _resolve(self):
        idx, val, cnt = self._pending
        n = _k._count(cnt)
        self._indices, self._data = _k._trim(idx, n), _k._trim(val, n)
        self._pending = None

    @property
    def indices(self):
        if self._pending is not None:
            self._resolve()
        return self._indices

    @indices.setter
    def indices(self, value):
        if self._pending is not None:
            self._resolve()
        self._indices = value

    @property
    def data(self):
        if self._pending is not None:
            self._resolve()
        return self._data

    @data.setter
    def data(self, value):
        if self._pending is not None:
            self._resolve()
        self._data = value

    # -- basic properties ---------------------------------------------------------------------
    @property
    def dtype(self):
        if self._pending is not None:
            return dt.np_dtype(self._pending[1].dtype)
        return dt.np_dtype(self._data.dtype)

    def getnnz(self):
        """Get the count of explicitly-stored values"""
        return int(self.indices.numel())

    nnz = property(fget=getnnz, doc=getnnz.__doc__)

    def _size(self):
        return int(np.prod(self.shape, dtype=object))

    # -- converters (flat.py:48-82) -------------------------------------------------------------
    @staticmethod
    def from_ndarray(arr):
        """Converts an array-like (host or device) to a FlatSparray object."""
        shape = tuple(arr.shape) if hasattr(arr, "shape") else np.shape(arr)
        if _is_complex_data(arr):
            from ._complex import ComplexFlatSparray
            return ComplexFlatSparray._from_dense(arr)
        dense = _to_device(arr)
        idx, val = _k.select_nonzero(dense)
        return FlatSparray(idx, val, shape=shape, is_canonical=True)

    @staticmethod
    def from_spmatrix(mat):
        """Converts a scipy.sparse matrix to a FlatSparray object"""
        try:
            mat.sum_duplicates()
        except AttributeError:
            pass
        mat = mat.tocoo()
        inds = mat.row.astype(np.int64) * mat.shape[1] + mat.col.astype(np.int64)
        return FlatSparray(inds, mat.data, shape=mat.shape, is_canonical=False)

    def toarray(self):
        """Dense numpy array (device scatter, then one device-to-host copy)."""
        return self.todense_device().cpu().numpy()

    def todense_device(self):
        return _k.scatter_dense(self.indices, self.data, self._size()).reshape(self.shape)

    def tocoo(self):
        assert len(self.shape) == 2
        idx = self.indices.cpu().numpy()
        row, col = np.divmod(idx, self.shape[1])
        return ss.coo_matrix((self.data.cpu().numpy(), (row, col)), shape=self.shape)

    def to_numpy(self):
        """(indices, data) as host numpy arrays."""
        return self.indices.cpu().numpy(), self.data.cpu().numpy()

    def nonzero(self):
        """Tuple of host index arrays of the non-zero stored elements (flat.py:91-95)."""
        keep = _k.unary_map("ne_s", self.data, 0)
        idx, _ = _k.select_pairs(self.indices, self.data, keep)
        return np.unravel_index(idx.cpu().numpy(), self.shape)

    # -- axis permutation (flat.py:97-112) -------------------------------------------------------
    def transpose(self, *axes):
        ndim = len(self.shape)
        if ndim < 2:
            return self
        if not axes:
            axes = tuple(range(ndim - 1, -1, -1))
        elif len(axes) == 1 and ndim > 1:
            axes = tuple(axes[0])
        axes = tuple(int(a) + ndim if int(a) < 0 else int(a) for a in axes)
        if sorted(axes) != list(range(ndim)):
            raise ValueError("axes don't match array")
        if axes == tuple(range(ndim)):
            return self
        new_shape = tuple(self.shape[a] for a in axes)
        new_idx, new_val = _k.transpose(self.indices, self.data, self.shape, axes)
        return FlatSparray(new_idx, new_val, new_shape, is_canonical=True)

    def swapaxes(self, axis1, axis2):
        """numpy.swapaxes; sugar over transpose (absent upstream)."""
        ndim = len(self.shape)
        axes = list(range(ndim))
        a1, a2 = int(axis1) % ndim, int(axis2) % ndim
        axes[a1], axes[a2] = axes[a2], axes[a1]
        return self.transpose(*axes)

    # -- diagonal (flat.py:114-137) ----------------------------------------------------------------
    def diagonal(self, offset=0, axis1=0, axis2=1):
        if axis1 == axis2:
            raise ValueError("axis1 and axis2 cannot be the same")
        if len(self.shape) < 2:
            raise ValueError("diagonal requires at least two dimensions")
        if len(self.shape) > 2:
            raise NotImplementedError("diagonal() is NYI for ndim > 2")
        if axis1 != 0 or axis2 != 1:
            raise NotImplementedError("diagonal() is NYI for non-default axes")
        if offset >= 0:
            n = min(self.shape[0], self.shape[1] - offset)
            ranges = [[0, n, 1], [offset, n + offset, 1]]
        else:
            n = min(self.shape[0] + offset, self.shape[1])
            ranges = [[-offset, n - offset, 1], [0, n, 1]]
        if n <= 0:
            return FlatSparray([], np.zeros(0, self.dtype), shape=(0,), is_canonical=True)
        flat_idx = _k.combine_ranges(ranges, self.shape, n, inner=True, device=self.indices.device)
        return self._getitem_flatidx(flat_idx, (n,))

    def __repr__(self):
        return "<%s-FlatSparray of type %s\n\twith %d stored elements>" % (self.shape, self.dtype, self.getnnz())

    def __str__(self):
        idx, data = self.to_numpy()
        multi = np.unravel_index(idx, self.shape) if len(idx) else [[] for _ in self.shape]
        return "\n".join("  %s\t%s" % (x[1:], x[0]) for x in zip(data, *multi))

    # -- reshape family: metadata only, buffers shared (flat.py:173-191) ---------------------------
    def reshape(self, new_shape):
        new_shape = list(new_shape) if not isinstance(new_shape, numbers.Integral) else [new_shape]
        if -1 in new_shape:
            assert sum(d == -1 for d in new_shape) == 1, "Only one -1 allowed"
            k = new_shape.index(-1)
            new_shape[k] = self._size() // -int(np.prod(new_shape, dtype=object))
        else:
            assert np.prod(new_shape, dtype=object) >= len(self.data)
        return FlatSparray(self.indices, self.data, shape=new_shape, is_canonical=True)

    def resize(self, new_shape):
        assert np.prod(new_shape, dtype=object) >= len(self.data)
        self.shape = tuple(int(s) for s in new_shape)

    def ravel(self):
        return FlatSparray(self.indices, self.data, shape=(self._size(),), is_canonical=True)

    # -- indexing (flat.py:193-325, 366-393) --------------------------------------------------------
    def __getitem__(self, indices):
        items, kind = ix.prepare_indices(indices, self.shape)
        if kind == ix.EMPTY_SLICE_INDEX_MASK:                         # all colons
            return self
        if kind == ix.INTEGER_INDEX_MASK:                             # scalar lookup
            flat = int(np.ravel_multi_index(items, self.shape))
            q = torch.tensor([flat], dtype=torch.int64, device=self.indices.device)
            pos, found = _k.lower_bound(self.indices, q)
            if not bool(found.item()):
                return 0
            return self.data[pos[0]].cpu().numpy()[()]
        if not (kind & ix.ARRAY_INDEX_MASK):                          # ints + slices: a box
            ranges, new_shape = ix.indices_to_ranges(items, self.shape)
            if any(s == 0 for s in new_shape):
                return FlatSparray([], np.zeros(0, self.dtype), tuple(new_shape), is_canonical=True)
            if np.any(ranges[:, 2] < 0):
                raise NotImplementedError("negative slice steps are not supported")
            new_idx, new_val = _k.getitem_box(self.indices, self.data, self.shape, ranges, new_shape)
            return FlatSparray(new_idx, new_val, tuple(new_shape), is_canonical=True)
        # fancy: per-output-axis offset arrays, outer-summed on the device
        offsets, new_shape = ix.fancy_plan(items, self.shape)
        if any(s == 0 for s in new_shape):
            return FlatSparray([], np.zeros(0, self.dtype), tuple(new_shape), is_canonical=True)
        query = _k.outer_sum(offsets, new_shape, self.indices.device)
        return self._getitem_flatidx(query, new_shape)

    def _getitem_flatidx(self, flat_idx, new_shape, is_sorted=True):     # flat.py:383-393
        """Result index = position in the query; one binary search per query cell
        (O(q log nnz), no sort of the query, repeated / unsorted queries fine)."""
        new_idx, new_val = _k.lookup_compact(self.indices, self.data, flat_idx)
        return FlatSparray(new_idx, new_val, tuple(new_shape), is_canonical=True)

    def __setitem__(self, indices, val):                                  # flat.py:327-364
        items, kind = ix.prepare_indices(indices, self.shape)
        if kind == ix.EMPTY_SLICE_INDEX_MASK:
            raise ValueError("Assigning to entire FlatSparray would densify.")
        dev = self.indices.device
        if kind == ix.INTEGER_INDEX_MASK:
            flat = int(np.ravel_multi_index(items, self.shape))
            self._setitem_flatidx(torch.tensor([flat], dtype=torch.int64, device=dev), val)
            return
        if not (kind & ix.ARRAY_INDEX_MASK):                          # ints + slices: the whole box gets stored
            ranges, idx_shape = ix.indices_to_ranges(items, self.shape)
            if any(s == 0 for s in idx_shape):
                return
            if np.any(ranges[:, 2] < 0):
                raise NotImplementedError("negative slice steps are not supported")
            n = int(np.prod(idx_shape, dtype=object)) if idx_shape else 1
            box = _k.combine_ranges(ranges, self.shape, n, inner=False, device=dev)
            self._setitem_flatidx(box, val)
            return
        raise NotImplementedError("Fancy assignment is still NYI")

    def setdiag(self, values, offset=0):                                  # flat.py:139-160
        if len(self.shape) < 2:
            raise ValueError("setdiag() requires at least two dimensions")
        if len(self.shape) > 2:
            raise NotImplementedError("setdiag() is NYI for ndim > 2")
        if offset >= 0:
            n = min(self.shape[0], self.shape[1] - offset)
            ranges = [[0, n, 1], [offset, n + offset, 1]]
        else:
            n = min(self.shape[0] + offset, self.shape[1])
            ranges = [[-offset, n - offset, 1], [0, n, 1]]
        if n <= 0:
            return self
        diag = _k.combine_ranges(ranges, self.shape, n, inner=True, device=self.indices.device)
        self._setitem_flatidx(diag, values)

    def _setitem_flatidx(self, flat_idx, values):                         # flat.py:395-407
        """Store `values` (scalar or one per index) at sorted flat indices: one fused union merge
        in which the assigned value wins and untouched elements survive; the structure grows as
        needed.  Mutates this array like the reference does."""
        n = flat_idx.numel()
        if np.isscalar(values) or (hasattr(values, "ndim") and values.ndim == 0):
            vals = torch.full((n,), values.item() if hasattr(values, "item") else values,
                              dtype=self.data.dtype, device=flat_idx.device)
        else:
            vals = _to_device(values, self.data.dtype)
            if vals.numel() == 1:
                vals = vals.expand(n).contiguous()
            if vals.numel() != n:
                raise ValueError("shape mismatch: value array of size %d could not be broadcast to indexing "
                                 "result of size %d" % (vals.numel(), n))
        idx, data, _ = _k.setop_binop("union", "override", self.indices, self.data, flat_idx, vals)
        self.indices, self.data = idx, data

    # -- sparse (op) sparse helpers (flat.py:409-434) ---------------------------------------------
    def _promoted(self, other, op):
        """Values of both operands in np.promote_types(a, b) (flat.py:415)."""
        a_val, b_val = self.data, other.data
        if a_val.dtype != b_val.dtype:
            common = dt.promote(a_val.dtype, b_val.dtype)
            if a_val.dtype != common:
                a_val = _k.cast(a_val, common)
            if b_val.dtype != common:
                b_val = _k.cast(b_val, common)
        if a_val.dtype == torch.bool and op == "sub":
            raise TypeError("numpy boolean subtract, the `-` operator, is not supported")
        return a_val, b_val

    def _pairwise_sparray(self, other, op):
        """op(sparse, sparse) -> sparse on the UNION of the index sets; one fused kernel."""
        a_val, b_val = self._promoted(other, op)
        idx, val, cnt = _k.setop_binop("union", op, self.indices, a_val, other.indices, b_val, lazy=True)
        return FlatSparray._from_kernel(idx, val, cnt, self.shape)

    def _pairwise_sparray_fixed_zero(self, other, op):
        """op(sparse, sparse) -> sparse on the INTERSECTION (op(x, 0) == 0); one fused kernel."""
        a_val, b_val = self._promoted(other, op)
        idx, val, cnt = _k.setop_binop("intersect", op, self.indices, a_val, other.indices, b_val, lazy=True)
        return FlatSparray._from_kernel(idx, val, cnt, self.shape)

    # -- broadcasting and dense operands (flat.py:436-472) ------------------------------------------
    def _broadcast(self, shape):
        """This array broadcast to `shape`, WITHOUT densifying (the reference densifies here,
        flat.py:470-472 "TODO: fix this hack"): an index-expansion kernel replicates the stored
        elements along the broadcast axes, then one radix sort makes the result canonical.  Like
        the reference's dense detour (from_ndarray), explicit zeros do not survive."""
        shape = tuple(int(s) for s in shape)
        padded = (1,) * (len(shape) - len(self.shape)) + tuple(self.shape)
        for i, o in zip(padded, shape):
            if i != o and i != 1:
                raise ValueError("shape mismatch: objects cannot be broadcast to a single shape")
        idx, val = _k.broadcast_expand(self.indices, self.data, padded, shape)
        bits = max(int(np.prod(shape, dtype=object)) - 1, 1).bit_length()
        idx, val = _k.rekey_sort(idx, val, [], [], 1, bits)
        keep = _k.unary_map("ne_s", val, 0)
        idx, val = _k.select_pairs(idx, val, keep)
        return FlatSparray(idx, val, shape, is_canonical=True)

    def _handle_broadcasting(self, other):
        """(lhs, rhs) with a common shape; rhs is a FlatSparray or a C-contiguous device tensor."""
        if ss is not None and ss.issparse(other):
            other = FlatSparray.from_spmatrix(other)
        if isinstance(other, FlatSparray):
            if other.shape == self.shape:
                return self, other
            bshape = tuple(int(s) for s in np.broadcast_shapes(self.shape, other.shape))
            lhs = self if self.shape == bshape else self._broadcast(bshape)
            rhs = other if other.shape == bshape else other._broadcast(bshape)
            return lhs, rhs
        if isinstance(other, torch.Tensor):
            dense = other.to(self.indices.device)
            oshape = tuple(dense.shape)
        else:
            arr = np.asarray(other)
            dt.torch_dtype(arr.dtype)
            oshape = arr.shape
            dense = None
        bshape = tuple(int(s) for s in np.broadcast_shapes(self.shape, oshape))
        lhs = self if self.shape == bshape else self._broadcast(bshape)
        if dense is None:
            dense = torch.from_numpy(np.array(np.broadcast_to(arr, bshape), order="C")).to(self.indices.device)
        else:
            dense = dense.expand(bshape).contiguous()
        return lhs, dense

    def _pairwise_dense2dense(self, other, op):
        """op(dense, sparse) -> dense ndarray with the dense operand's dtype (flat.py:436-442):
        gather the touched cells, combine, scatter back into a copy."""
        flat = other.reshape(-1)
        cells = _k.gather(flat, self.indices)
        common = dt.promote(cells.dtype, self.data.dtype)
        a = cells if cells.dtype == common else _k.cast(cells, common)
        b = self.data if self.data.dtype == common else _k.cast(self.data, common)
        r = _k.binary_map(op, a, b)
        if r.dtype != flat.dtype:
            r = _k.cast(r, flat.dtype)
        out = flat.clone()
        _k.scatter_into(out, self.indices, r)
        return out.reshape(other.shape).cpu().numpy()

    def _pairwise_dense2sparse(self, other, op):
        """op(sparse, dense) -> sparse on this array's structure (flat.py:444-449)."""
        cells = _k.gather(other.reshape(-1), self.indices)
        common = dt.promote(cells.dtype, self.data.dtype)
        a = self.data if self.data.dtype == common else _k.cast(self.data, common)
        b = cells if cells.dtype == common else _k.cast(cells, common)
        return self._with_data(_k.binary_map(op, a, b))

    def _dense_elementwise(self, other, op):
        """op(self.toarray(), dense) computed on the device -> dense ndarray (the reference's
        `ufunc(self.toarray(), other)` branches, e.g. flat.py:488,593,605)."""
        mine = self.todense_device().reshape(-1)
        flat = other.reshape(-1)
        common = dt.promote(mine.dtype, flat.dtype)
        a = mine if mine.dtype == common else _k.cast(mine, common)
        b = flat if flat.dtype == common else _k.cast(flat, common)
        return _k.binary_map(op, a, b).reshape(other.shape).cpu().numpy()

    def _comparison(self, other, method_name, op, op_symbol):              # flat.py:474-488
        if _is_scalar(other):
            np_op = getattr(np, {"lt": "less", "le": "less_equal", "gt": "greater", "ge": "greater_equal",
                                 "eq": "equal", "ne": "not_equal"}[op])
            if not np_op(0, other):
                return self._scalar_map(op + "_s", other, compare=True)
            kind = "nonzero scalar" if other != 0 else "0"
            warnings.warn("FlatSparray %s %s densifies" % (op_symbol, kind), _EfficiencyWarning)
            return self._dense_scalar_map(op + "_s", other, compare=True)           # ufunc(self.toarray(), other)
        lhs, rhs = self._handle_broadcasting(other)
        if isinstance(rhs, FlatSparray):
            return lhs._pairwise_sparray(rhs, op)
        return lhs._dense_elementwise(rhs, op)

    def __add__(self, other):                                              # flat.py:490-505
        if _is_scalar(other):
            if other == 0:
                return self.copy()
            warnings.warn("FlatSparray + nonzero scalar densifies", _EfficiencyWarning)
            return self._dense_add_scalar(other)                                     # self.toarray() + other
        lhs, rhs = self._handle_broadcasting(other)
        if isinstance(rhs, FlatSparray):
            return lhs._pairwise_sparray(rhs, "add")
        return lhs._pairwise_dense2dense(rhs, "add")

    def __sub__(self, other):                                              # base.py:45-46
        # a - b == a + (-b) bit for bit (IEEE / two's complement); fused so -b is never materialised
        if _is_scalar(other):
            return self.__add__(-other)
        lhs, rhs = self._handle_broadcasting(other)
        if isinstance(rhs, FlatSparray):
            return lhs._pairwise_sparray(rhs, "sub")
        # dense: a + (-b) with the dense operand negated on the device
        neg = _k.unary_map("neg", rhs.reshape(-1)).reshape(rhs.shape)
        return lhs._pairwise_dense2dense(neg, "add")

    def __mul__(self, other):                                              # flat.py:507-518
        if _is_scalar(other):
            if isinstance(other, (complex, np.complexfloating)) and complex(other).imag != 0:
                return self.astype(np.complex64 if self.dtype == np.float32 else np.complex128) * other
            return self._scalar_map("mul_s", other)
        lhs, rhs = self._handle_broadcasting(other)
        if isinstance(rhs, FlatSparray):
            return lhs._pairwise_sparray_fixed_zero(rhs, "mul")
        return lhs._pairwise_dense2sparse(rhs, "mul")

    def _divide(self, other, div_func="divide", rdivide=False):            # flat.py:520-536
        if ss is not None and ss.issparse(other):
            other = FlatSparray.from_spmatrix(other)
        if isinstance(other, FlatSparray) or rdivide:
            # a sparse divisor (or a reflected division) cannot keep sparsity: dense result, computed on the
            # device (flat.py:520-527)
            return self._dense_divide(other, div_func, rdivide)
        if _is_scalar(other):
            if div_func in ("true_divide", "divide"):
                return self.__mul__(1. / other)                            # flat.py:528-529
            return self._scalar_map("floordiv_s", other)
        lhs, rhs = self._handle_broadcasting(other)                        # dense / -> sparse on our structure
        if div_func == "true_divide":
            if not rhs.dtype.is_floating_point:
                rhs = _k.cast(rhs.reshape(-1), torch.float64).reshape(rhs.shape)
            if not lhs.data.dtype.is_floating_point:
                lhs = lhs.astype(np.float64)
        return lhs._pairwise_dense2sparse(rhs, "floor_divide" if div_func == "floor_divide" else "divide")

    def dot(self, other):                                                  # flat.py:538-568
        ax1 = len(self.shape) - 1
        oshape = tuple(other.shape)
        ax2 = max(0, len(oshape) - 2)
        if self.shape[ax1] != oshape[ax2]:
            raise ValueError("shapes %s and %s not aligned" % (self.shape, oshape))
        nrows = self._size() // self.shape[ax1] if self.shape[ax1] else 0
        if ss is not None and ss.issparse(other):
            other = FlatSparray.from_spmatrix(other)
        if isinstance(other, FlatSparray) and len(oshape) >= 2:
            # sparse . sparse (flat.py:545-561): SpGEMM on the sorted flat layout -- expand / sort / compress
            # instead of the reference's COO -> scipy CSR matmul -> from_spmatrix detour
            out_shape = tuple(self.shape[:-1]) + oshape[:ax2] + oshape[ax2 + 1:]
            if ax2 != 0:                                                   # bring the contracted axis first
                other = other.transpose(*((ax2,) + tuple(range(ax2)) + tuple(range(ax2 + 1, len(oshape)))))
            k = self.shape[ax1]
            ncols = int(np.prod(other.shape[1:], dtype=object))
            vdt = dt.promote(self.data.dtype, other.data.dtype)
            if vdt in (torch.bool, torch.uint8, torch.int8, torch.int16):
                vdt = torch.int32 if vdt != torch.bool else torch.int32
            a_val = self.data if self.data.dtype == vdt else _k.cast(self.data, vdt)
            b_val = other.data if other.data.dtype == vdt else _k.cast(other.data, vdt)
            idx, val = _k.spgemm(self.indices, a_val, nrows, k, other.indices, b_val, ncols)
            want = dt.promote(self.data.dtype, other.data.dtype)
            if val.dtype != want:
                val = _k.cast(val, want)
            return FlatSparray(idx, val, shape=out_shape, is_canonical=True)
        if len(oshape) == 2:
            # dense matrix operand: one SpMV per column (dense result, like flat.py:568 without densifying self)
            dense = _to_device_nd(other)
            cols = dense.t().contiguous()
            vdt = dt.promote(self.data.dtype, cols.dtype)
            out = torch.empty((oshape[1], nrows), dtype=vdt if vdt.is_floating_point else torch.float64,
                              device=self.indices.device)
            for j in range(oshape[1]):
                out[j] = _k.spmv(self.indices, self.data, self.shape[ax1], nrows, cols[j], vdt)
            res = out.t().contiguous().reshape(self.shape[:-1] + (oshape[1],))
            return res.cpu().numpy()
        if len(oshape) != 1:
            raise NotImplementedError("dot with a dense operand of more than two dimensions is out of scope")
        if isinstance(other, FlatSparray):                                 # sparse vector: sparse result
            x = other.todense_device().reshape(-1)
            vdt = dt.promote(self.data.dtype, x.dtype)
            y = _k.spmv(self.indices, self.data, self.shape[ax1], nrows, x, vdt)
            if len(self.shape) == 1:
                return y.cpu().numpy()[0]
            return FlatSparray.from_ndarray(y.reshape(self.shape[:-1]))
        x = _to_device(other)
        vdt = dt.promote(self.data.dtype, x.dtype)
        y = _k.spmv(self.indices, self.data, self.shape[ax1], nrows, x, vdt)
        if not vdt.is_floating_point:
            # numpy's dot of integer operands is an integer (flat.py:566,568); the float64 accumulation is
            # exact below 2^53
            y = _k.cast(y, vdt if vdt != torch.bool else torch.int64)
        if len(self.shape) == 1:
            return y.cpu().numpy()[0]                                      # flat.py:564-566
        return y.reshape(self.shape[:-1]).cpu().numpy()                    # dense rhs -> dense result

    def __pow__(self, exponent):                                           # flat.py:570-578
        if exponent == 0:
            warnings.warn("FlatSparray ** 0 densifies", _EfficiencyWarning)
            return np.ones(self.shape, dtype=self.dtype)
        elif exponent < 0:
            warnings.warn("FlatSparray ** negative exponent densifies", _EfficiencyWarning)
            if self.dtype.kind in "iub" and isinstance(exponent, (int, np.integer)):
                raise ValueError("Integers to negative integer powers are not allowed.")
            return self._dense_scalar_map("pow_s", exponent)                        # self.toarray() ** exponent
        return self._scalar_map("pow_s", exponent)

    def _with_data(self, data):                                            # flat.py:580-581
        return FlatSparray(self.indices.clone(), data, self.shape, is_canonical=True)

    def _map_values(self, data, op, scalar, compare=False):
        """data (op) python/numpy scalar with numpy's result dtype, on any device vector of this array's dtype."""
        if isinstance(scalar, (torch.Tensor, np.ndarray)):
            scalar = scalar.item()
        if isinstance(scalar, complex) or np.iscomplexobj(scalar):
            raise TypeError("complex values are not supported by the sparray_b200 CUDA kernels")
        np_dt = self.dtype
        if compare:
            res = np.result_type(np_dt, scalar) if np_dt != np.bool_ else np.result_type(np.int64, scalar)
        elif op == "mul_s" and np_dt == np.bool_ and isinstance(scalar, (bool, np.bool_)):
            res = np.dtype(np.bool_)
        else:
            res = np.result_type(np_dt, scalar)
            if res == np.bool_:
                res = np.dtype(np.int64)
        tdt = dt.torch_dtype(res)
        if data.dtype != tdt:
            data = _k.cast(data, tdt)
        return _k.unary_map(op, data, scalar)

    def _scalar_map(self, op, scalar, compare=False):
        """data (op) python/numpy scalar with numpy's result dtype."""
        return self._with_data(self._map_values(self.data, op, scalar, compare))

    # -- branches whose result is dense by nature (flat.py:478-481,494-496,520-527,570-577,592-593,604-605):
    #    the reference computes ufunc(self.toarray(), ...) on the host; here the dense array is built and
    #    combined on the device and comes back as the ndarray the reference returns
    def _dense_scalar_map(self, op, scalar, compare=False):
        dense = self.todense_device().reshape(-1)
        return self._map_values(dense, op, scalar, compare).reshape(self.shape).cpu().numpy()

    def _dense_add_scalar(self, scalar):
        if isinstance(scalar, complex) or np.iscomplexobj(scalar):
            raise TypeError("complex values are not supported by the sparray_b200 CUDA kernels")
        res = np.result_type(self.dtype, scalar)
        if res == np.bool_:
            res = np.dtype(np.int64)
        tdt = dt.torch_dtype(res)
        dense = self.todense_device().reshape(-1)
        if dense.dtype != tdt:
            dense = _k.cast(dense, tdt)
        fill = torch.full_like(dense, scalar)                    # a constant operand: plumbing, the add is ours
        return _k.binary_map("add", dense, fill).reshape(self.shape).cpu().numpy()

    def _dense_divide(self, other, div_func, rdivide):
        mine = self.todense_device()
        if isinstance(other, FlatSparray):
            theirs = other.todense_device()
        elif _is_scalar(other):
            theirs = torch.full_like(mine, other, dtype=dt.torch_dtype(np.result_type(self.dtype, other)))
        else:
            theirs = _to_device_nd(other)
        bshape = tuple(int(x) for x in np.broadcast_shapes(tuple(mine.shape), tuple(theirs.shape)))
        a = mine.expand(bshape).contiguous().reshape(-1)
        b = theirs.expand(bshape).contiguous().reshape(-1)
        common = dt.promote(a.dtype, b.dtype)
        if div_func in ("true_divide", "divide") and not common.is_floating_point:
            common = torch.float64
        a = a if a.dtype == common else _k.cast(a, common)
        b = b if b.dtype == common else _k.cast(b, common)
        if rdivide:
            a, b = b, a
        op = "floor_divide" if div_func == "floor_divide" else "divide"
        with np.errstate(all="ignore"):
            return _k.binary_map(op, a, b).reshape(bshape).cpu().numpy()

    def _rdot(self, other):                                                # base.py:79-82: other.dot(self.toarray())
        """dense . sparse -> dense, as (sparse^T . dense^T)^T: SpMV passes over the transposed array."""
        arr = np.asarray(other)
        if arr.ndim == 1:
            return self.transpose().dot(arr)
        lead = arr.shape[:-1]
        res = self.transpose().dot(arr.reshape(-1, arr.shape[-1]).T)       # (n, rows)
        return np.ascontiguousarray(res.T).reshape(lead + tuple(self.shape[1:]))

    def __rmatmul__(self, other):
        return self._rdot(other)

    def _float_map(self, op):
        """np.<op>(data) with the result dtype the installed numpy gives (flat.py:722-739 applies the ufunc to
        `data`): integer data stays integer for sign / floor / ceil / trunc (numpy >= 2.1), everything else is
        computed in floating point."""
        data = self.data
        if data.dtype.is_floating_point:
            return self._with_data(_k.unary_map(op, data))
        res = getattr(np, op)(np.zeros(1, dtype=self.dtype)).dtype
        if res.kind in "iub" and op in ("floor", "ceil", "trunc"):
            out = data.clone()                           # the identity on integers
        elif res.kind in "iu" and op == "sign" and data.dtype in (torch.int32, torch.int64):
            out = _k.unary_map(op, data)
        else:
            out = _k.unary_map(op, _k.cast(data, torch.float64))
        want = dt.torch_dtype(res)
        return self._with_data(out if out.dtype == want else _k.cast(out, want))

    def minimum(self, other):                                              # flat.py:583-593
        if _is_scalar(other) and other >= 0:
            return self._scalar_map("min_s", other)
        if _is_scalar(other):
            return self._dense_scalar_map("min_s", other)                            # np.minimum(self.toarray(), other)
        lhs, rhs = self._handle_broadcasting(other)
        if isinstance(rhs, FlatSparray):
            return lhs._pairwise_sparray(rhs, "minimum")
        return lhs._dense_elementwise(rhs, "minimum")

    def maximum(self, other):                                              # flat.py:595-605
        if _is_scalar(other) and other <= 0:
            return self._scalar_map("max_s", other)
        if _is_scalar(other):
            return self._dense_scalar_map("max_s", other)                            # np.maximum(self.toarray(), other)
        lhs, rhs = self._handle_broadcasting(other)
        if isinstance(rhs, FlatSparray):
            return lhs._pairwise_sparray(rhs, "maximum")
        return lhs._dense_elementwise(rhs, "maximum")

    # -- reductions (flat.py:607-626) --------------------------------------------------------------
    def sum(self, axis=None, dtype=None):
        if dtype is None:
            dtype = self.dtype
        dtype = np.dtype(dtype)
        if dtype.kind == "c":
            raise TypeError("complex values are not supported by the sparray_b200 CUDA kernels")
        if axis is not None:
            axis = int(axis)
            if axis < 0:
                axis += len(self.shape)
        new_shape = None if axis is None else self.shape[:axis] + self.shape[axis + 1:]
        if not new_shape:
            total = _k.sum_all(self.data)
            if dtype == np.bool_:
                return np.bool_(total != 0)
            if dtype.kind in "iu":
                return np.int64(total).astype(dtype)       # data.sum(dtype=self.dtype) wraps like the dtype (flat.py:608-610)
            return dtype.type(total)
        # float64 output whatever the input dtype, like np.bincount(weights=)
        if self._pending is not None and _k.dense_sum_cells(self.shape, axis):
            # the producer's count is still on the device and this path does not need it on the host
            p_idx, p_val, p_cnt = self._pending
            new_idx, new_val, cnt = _k.sum_axis(p_idx, p_val, self.shape, axis, n_dev=p_cnt)
        else:
            new_idx, new_val, cnt = _k.sum_axis(self.indices, self.data, self.shape, axis)
        return FlatSparray._from_kernel(new_idx, new_val, cnt, new_shape)

    # -- unary data ops (flat.py:677-739) -----------------------------------------------------------
    def __abs__(self):
        return self._with_data(_k.unary_map("abs", self.data))

    def __neg__(self):
        if self.data.dtype == torch.bool:
            raise TypeError("The numpy boolean negative, the `-` operator, is not supported")
        return self._with_data(_k.unary_map("neg", self.data))

    @property
    def real(self):
        return self.copy()

    @property
    def imag(self):
        return self._with_data(torch.zeros_like(self.data))

    def __imul__(self, other):
        if _is_scalar(other):
            out = self._scalar_map("mul_s", other)
            if out.data.dtype != self.data.dtype:
                raise TypeError("Cannot cast ufunc 'multiply' output from %s to %s" % (out.dtype, self.dtype))
            self.data = out.data
            return self
        return NotImplemented

    def __itruediv__(self, other):
        if _is_scalar(other):
            recip = 1.0 / other
            out = self._scalar_map("mul_s", recip)
            if out.data.dtype != self.data.dtype:
                raise TypeError("Cannot cast ufunc 'multiply' output from %s to %s" % (out.dtype, self.dtype))
            self.data = out.data
            return self
        return NotImplemented

    def astype(self, t):
        if np.dtype(t).kind == "c":
            from ._complex import ComplexFlatSparray
            part = self.astype(np.float32 if np.dtype(t) == np.complex64 else np.float64)
            return ComplexFlatSparray._from_parts(part, part._with_data(torch.zeros_like(part.data)))
        return self._with_data(_k.cast(self.data, dt.torch_dtype(np.dtype(t))))

    def conj(self):
        return self.copy()

    def copy(self):
        return self._with_data(self.data.clone())

    def min(self):
        return self.dtype.type(_k.min_max(self.data, "min"))

    def max(self):
        return self.dtype.type(_k.min_max(self.data, "max"))


def _make_unary(name):
    def method(self):
        return self._float_map(name)
    method.__doc__ = "Element-wise %s.\n\nSee numpy.%s for more information." % (name, name)
    method.__name__ = name
    return method


for _name in _FIXED_POINT_AT_ZERO:                                        # flat.py:722-739
    setattr(FlatSparray, _name, _make_unary(_name))
