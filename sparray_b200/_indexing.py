"""Host-side index preparation for FlatSparray.__getitem__ -- the mirror of the
reference's ``_prepare_indices`` (flat.py:193-252) and ``_indices_to_ranges``
(flat.py:366-381).  Pure Python / numpy on *index objects* (never on stored
data); importable and testable without CUDA."""
import numbers

import numpy as np

# kinds of multidimensional indexing (flat.py:12-16)
EMPTY_SLICE_INDEX_MASK = 0b1
SLICE_INDEX_MASK = 0b10
INTEGER_INDEX_MASK = 0b100
ARRAY_INDEX_MASK = 0b1000


def _to_host_array(idx):
    if hasattr(idx, "detach"):          # torch tensor used as an index
        idx = idx.detach().cpu().numpy()
    return np.asanyarray(idx)


def prepare_indices(index, shape):
    """Normalise ``index`` against ``shape``.  Returns (tuple of per-axis
    int | slice | 1-D ndarray, kind bitmask).  Raises exactly where the
    reference does: IndexError for too many indices / a second Ellipsis / an
    out-of-range integer, NotImplementedError for N-D index arrays."""
    if not isinstance(index, tuple):
        index = (index,)
    ndim = len(shape)
    ellipsis_at = []
    items = []
    for idx in index:
        if idx is Ellipsis:
            ellipsis_at.append(len(items))
        elif not isinstance(idx, (slice, numbers.Integral)):
            idx = _to_host_array(idx)
            if idx.dtype == np.bool_:
                items.extend(idx.nonzero())
                continue
            if idx.ndim > 1:
                raise NotImplementedError("Multi-dimensional indexing is NYI")
        items.append(idx)
    if len(ellipsis_at) > 1:
        raise IndexError("an index can only have a single ellipsis ('...')")
    missing = ndim - len(items)
    if ellipsis_at:
        pos, = ellipsis_at
        items[pos:pos + 1] = [slice(None)] * (missing + 1)
    elif missing > 0:
        items.extend([slice(None)] * missing)
    if len(items) > ndim:
        raise IndexError("too many indices for FlatSparray")
    assert len(items) == ndim

    kind = 0
    for axis, (idx, dim) in enumerate(zip(items, shape)):
        if isinstance(idx, numbers.Integral):
            if not (-dim <= idx < dim):
                raise IndexError("index %d is out of bounds for axis %d with size %d" % (idx, axis, dim))
            items[axis] = int(idx) + dim if idx < 0 else int(idx)
            kind |= INTEGER_INDEX_MASK
        elif isinstance(idx, slice):
            kind |= EMPTY_SLICE_INDEX_MASK if idx == slice(None) else SLICE_INDEX_MASK
        else:
            arr = np.asarray(idx)
            if arr.size and (arr.min() < -dim or arr.max() >= dim):
                raise IndexError("index out of bounds for axis %d with size %d" % (axis, dim))
            items[axis] = np.where(arr < 0, arr + dim, arr).astype(np.int64)
            kind |= ARRAY_INDEX_MASK
    return tuple(items), kind


def len_range(start, stop, step):
    """_merge.pyx:70-78 in Python ints (host copy of spb_len_range)."""
    if step > 0:
        return (stop - start - 1) // step + 1 if start < stop else 0
    return (start - stop - 1) // (-step) + 1 if stop < start else 0


def indices_to_ranges(items, shape):
    """All-int/slice index -> (ranges int64[ndim,3], new_shape)."""
    ranges = np.zeros((len(items), 3), dtype=np.int64)
    new_shape = []
    for axis, idx in enumerate(items):
        if isinstance(idx, slice):
            row = idx.indices(shape[axis])
            ranges[axis] = row
            new_shape.append(len_range(*row))
        else:
            ranges[axis] = (idx, idx + 1, 1)
    return ranges, new_shape


def c_strides(shape):
    strides = [1] * len(shape)
    for ax in range(len(shape) - 2, -1, -1):
        strides[ax] = strides[ax + 1] * int(shape[ax + 1])
    return strides


def fancy_plan(items, shape):
    """Mixed int / slice / 1-D array index (flat.py:282-324): returns
    (per-output-axis int64 offset arrays, new_shape).  The flat query index of
    output cell (c0, c1, ...) is sum_j offsets[j][c_j]; all integer/array
    indices are zipped into the single output axis at the position of the
    first of them (numpy's rule when they are adjacent, and the reference's
    rule always)."""
    strides = c_strides(shape)
    offsets = []
    new_shape = []
    inner_pos = None
    inner = None
    for axis, idx in enumerate(items):
        if isinstance(idx, slice):
            offs = np.arange(*idx.indices(shape[axis]), dtype=np.int64) * strides[axis]
            offsets.append(offs)
            new_shape.append(len(offs))
        else:
            term = np.asarray(idx, dtype=np.int64) * strides[axis]
            if inner_pos is None:
                inner_pos = len(offsets)
                offsets.append(None)
                new_shape.append(-1)
                inner = term
            else:
                inner = inner + term          # broadcasting zips scalars with arrays
    if inner_pos is not None:
        inner = np.atleast_1d(inner)
        offsets[inner_pos] = inner
        new_shape[inner_pos] = len(inner)
    return offsets, new_shape
