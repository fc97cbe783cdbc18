"""Host-side planning for the re-key + radix-sort kernels (csrc/sort.cu): which
axes can be merged, which strides the kernel needs, and how few key bits have
to be sorted.  Pure Python integers; importable and testable without CUDA."""

import functools

MAX_REKEY_DIMS = 8


def c_strides(shape):
    strides = [1] * len(shape)
    for ax in range(len(shape) - 2, -1, -1):
        strides[ax] = strides[ax + 1] * int(shape[ax + 1])
    return strides


def prod(xs):
    out = 1
    for x in xs:
        out *= int(x)
    return out


def rekey_strides(shape, out_axes):
    """Re-keying ``new = ravel(coords[out_axes], shape[out_axes])`` of C-order flat
    indices into ``shape``.  ``out_axes`` lists old axes in their new order (a
    permutation for transpose; one axis missing for an axis reduction).

    Returns (in_strides, out_strides): per *merged* old axis group, outermost
    first, the C-order stride of the group in the old shape and its stride in the
    new shape (0 when dropped).  Neighbouring old axes that stay neighbours in the
    same order (or are both dropped) are merged, and size-1 axes vanish, so the
    kernel does as few divisions as possible.  The last in_stride is always 1.
    """
    shape = [int(s) for s in shape]
    ndim = len(shape)
    new_shape = [shape[a] for a in out_axes]
    new_strides = c_strides(new_shape)
    out_of = {a: new_strides[j] for j, a in enumerate(out_axes)}
    old_strides = c_strides(shape)
    groups = []          # [size, in_stride, out_stride]
    for ax in range(ndim):
        if shape[ax] == 1:
            continue
        cur = [shape[ax], old_strides[ax], out_of.get(ax, 0)]
        if groups:
            g = groups[-1]
            both_dropped = g[2] == 0 and cur[2] == 0
            contiguous = g[2] != 0 and cur[2] != 0 and g[2] == cur[2] * cur[0]
            if both_dropped or contiguous:
                groups[-1] = [g[0] * cur[0], cur[1], cur[2]]
                continue
        groups.append(cur)
    if not groups:
        return [1], [0]
    if groups[-1][1] != 1:
        # trailing size-1 axes were skipped: the finest group must still have stride 1
        raise AssertionError("internal: finest group stride != 1")
    if len(groups) > MAX_REKEY_DIMS:
        raise NotImplementedError("more than %d non-mergeable axes" % MAX_REKEY_DIMS)
    return [g[1] for g in groups], [g[2] for g in groups]


def transpose_plan(shape, axes):
    return _transpose_plan(tuple(int(s) for s in shape), tuple(int(a) for a in axes))


@functools.lru_cache(maxsize=256)
def _transpose_plan(shape, axes):
    """Plan for FlatSparray.transpose: (in_strides, out_strides, sort_divisor, sort_bits).

    The input is sorted by the old key.  An LSD radix sort is stable, so if the
    trailing new axes axes[m:] are in increasing old order, elements that agree on
    the leading new axes already arrive in their final relative order: only the
    ravel of the leading m new axes -- new_key // prod(new_shape[m:]) -- has to
    be sorted."""
    shape = [int(s) for s in shape]
    axes = [int(a) for a in axes]
    in_s, out_s = rekey_strides(shape, axes)
    live = [a for a in axes if shape[a] != 1]          # size-1 axes never matter for ordering
    m = len(live)
    while m > 0 and (m == len(live) or live[m - 1] < live[m]):
        m -= 1
    # now live[m:] is the longest increasing suffix; live[:m] are the axes to sort by
    lead = prod(shape[a] for a in live[:m])
    tail = prod(shape[a] for a in live[m:])
    bits = (lead - 1).bit_length() if m > 0 else 0
    return in_s, out_s, tail, bits


def axis_sum_plan(shape, axis):
    return _axis_sum_plan(tuple(int(s) for s in shape), int(axis))


@functools.lru_cache(maxsize=256)
def _axis_sum_plan(shape, axis):
    """Plan for sum(axis): ('last', divisor) when the reduced axis is the last one
    (the remaining key stays sorted: no sort at all), else ('sort', in_strides,
    out_strides, sort_bits) -- dropping a middle axis interleaves the later axes,
    so the whole new key is sorted."""
    shape = [int(s) for s in shape]
    ndim = len(shape)
    trailing_ones = all(s == 1 for s in shape[axis + 1:])
    if axis == ndim - 1 or trailing_ones:
        return ("last", shape[axis] * prod(shape[axis + 1:]))
    out_axes = [a for a in range(ndim) if a != axis]
    in_s, out_s = rekey_strides(shape, out_axes)
    total = prod(shape[a] for a in out_axes)
    return ("sort", in_s, out_s, (total - 1).bit_length())
