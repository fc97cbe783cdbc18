"""Thin Python wrappers: allocate outputs as torch tensors (device memory and
streams are torch's job), call the C ABI (include/sparray_b200.h), trim to the
device-side count.  No arithmetic happens in Python or in torch ops here."""
import ctypes
import functools

import numpy as np
import torch

from . import _dtypes as dt
from ._lib import check, lib, ptr, require_cuda, stream, workspace


def _empty(n, dtype, like):
    return torch.empty(int(n), dtype=dtype, device=like.device)


def _trim(t, n):
    """First n elements of a worst-case-sized output.  When most of the buffer is unused the
    slice is copied so the big allocation goes straight back to the caching allocator (the
    reference's `c[:end]` views pin their full buffers)."""
    if n * 2 < t.numel():
        return t[:n].clone()
    return t[:n]


def _count(t):
    """Read a device int64 count (the one host sync of an op with a data-dependent size)."""
    return int(t.item())


def _i64c(t, name):
    """The seam's calling convention: int64, C-contiguous."""
    if t.dtype != torch.int64:
        raise ValueError("Buffer dtype mismatch for %s, expected int64 but got %s" % (name, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s is not C-contiguous" % name)
    return t


# ---- section 1: the compat seam -------------------------------------------------------------

def union1d_sorted(a, b, return_masks=False):
    require_cuda()
    a, b = _i64c(a, "a"), _i64c(b, "b")
    na, nb = a.numel(), b.numel()
    out = _empty(na + nb, torch.int64, a)
    cnt = _empty(1, torch.int64, a)
    lut = a_only = b_only = None
    if return_masks:
        lut = _empty(na + nb, torch.uint8, a)
        a_only = _empty(na, torch.bool, a)
        b_only = _empty(nb, torch.bool, a)
    check(lib.spb_union1d_sorted(ptr(a), na, ptr(b), nb, ptr(out), ptr(lut), ptr(a_only), ptr(b_only),
                                 ptr(cnt), workspace(), stream()))
    n = _count(cnt)
    if not return_masks:
        return out[:n]
    return out[:n], lut[:n], a_only, b_only


def intersect1d_sorted(a, b, return_inds=False):
    require_cuda()
    a, b = _i64c(a, "a"), _i64c(b, "b")
    na, nb = a.numel(), b.numel()
    m = min(na, nb)
    out = _empty(m, torch.int64, a)
    cnt = _empty(1, torch.int64, a)
    ai = bi = None
    if return_inds:
        ai = _empty(m, torch.int64, a)
        bi = _empty(m, torch.int64, a)
    check(lib.spb_intersect1d_sorted(ptr(a), na, ptr(b), nb, ptr(out), ptr(ai), ptr(bi), ptr(cnt),
                                     workspace(), stream()))
    n = _count(cnt)
    if not return_inds:
        return out[:n]
    return out[:n], ai[:n], bi[:n]


def combine_ranges(ranges, shape, result_size, inner=False, device=None):
    require_cuda()
    ranges = np.ascontiguousarray(ranges, dtype=np.int64)
    shape_arr = np.ascontiguousarray(shape, dtype=np.int64)
    if ranges.ndim != 2 or ranges.shape[1] != 3 or ranges.shape[0] != len(shape_arr):
        raise ValueError("ranges must have shape (ndim, 3)")
    out = torch.empty(int(result_size), dtype=torch.int64, device=device or "cuda")
    check(lib.spb_combine_ranges(ranges.ctypes.data_as(ctypes.c_void_p), shape_arr.ctypes.data_as(ctypes.c_void_p),
                                 len(shape_arr), ptr(out), int(result_size), int(bool(inner)), stream()))
    return out


def len_range(start, stop, step):
    return int(lib.spb_len_range(int(start), int(stop), int(step)))


# ---- section 2: fused sparse (op) sparse ----------------------------------------------------

_WIDEN = {torch.bool: torch.int32, torch.uint8: torch.int32, torch.int8: torch.int32, torch.int16: torch.int32}
_WIDEN.update(dt.COMPUTE)          # float16, uint16 / 32 / 64: no kernels of their own (see _dtypes.py)


def _bits(t):
    """same-sized integer view of a conversion-only dtype, for kernels that only move values"""
    return t.view(dt.BITS[t.dtype]) if t is not None and t.dtype in dt.BITS else t


def _wide(t):
    """arithmetic carrier of a conversion-only dtype"""
    return cast(t, dt.COMPUTE[t.dtype]) if t.dtype in dt.COMPUTE else t


def _back(t, orig):
    return cast(t, orig) if orig in dt.COMPUTE and t.dtype != orig else t


def setop_binop(kind, op, a_idx, a_val, b_idx, b_val, lazy=False):
    """kind: 'union' | 'intersect'.  Values must already share a dtype.  The fused kernels
    are instantiated for 4- and 8-byte values; bool / 8- / 16-bit operands are widened to
    int32 for the merge and narrowed back (bool: + is OR, * is AND, as in numpy).
    Returns (indices, values, None), or with lazy=True possibly (worst-case-sized indices,
    values, device count): the count is not read back, so the call does not synchronise."""
    require_cuda()
    assert a_val.dtype == b_val.dtype
    orig = a_val.dtype
    if orig in _WIDEN:
        a_val, b_val = cast(a_val, _WIDEN[orig]), cast(b_val, _WIDEN[orig])
    na, nb = a_idx.numel(), b_idx.numel()
    cap = na + nb if kind == "union" else min(na, nb)
    out_dtype = torch.bool if op in dt.COMPARISONS else a_val.dtype
    out_idx = _empty(cap, torch.int64, a_idx)
    out_val = _empty(cap, out_dtype, a_idx)
    cnt = _empty(1, torch.int64, a_idx)
    fn = lib.spb_union_binop if kind == "union" else lib.spb_intersect_binop
    check(fn(dt.BINOPS[op], dt.spb_dtype(a_val.dtype), ptr(a_idx), ptr(a_val), na, ptr(b_idx), ptr(b_val), nb,
             ptr(out_idx), ptr(out_val), ptr(cnt), workspace(), stream()))
    narrow = orig in _WIDEN and op not in dt.COMPARISONS
    if lazy and not narrow and cap:
        return out_idx, out_val, cnt
    n = _count(cnt)
    out_idx, out_val = _trim(out_idx, n), _trim(out_val, n)
    if narrow:
        out_val = cast(out_val, orig)
    return out_idx, out_val, None


# ---- section 3: maps, conversion, lookup -----------------------------------------------------

def unary_map(op, data, scalar=0):
    require_cuda()
    if data.dtype in dt.COMPUTE:
        orig = data.dtype
        if orig == torch.float16 and op.endswith("_s"):
            scalar = float(np.float16(scalar))     # numpy rounds a Python scalar to the array's float16 first
        out = unary_map(op, _wide(data), scalar)
        return out if out.dtype == torch.bool else _back(out, orig)
    out_dtype = torch.bool if op.endswith("_s") and op[:2] in ("lt", "le", "gt", "ge", "eq", "ne") else data.dtype
    out = torch.empty_like(data, dtype=out_dtype)
    is_float = data.dtype.is_floating_point
    check(lib.spb_unary_map(dt.UNOPS[op], dt.spb_dtype(data.dtype), ptr(data), ptr(out), data.numel(),
                            float(scalar) if is_float else 0.0, 0 if is_float else int(scalar), stream()))
    return out


def cast(data, dtype):
    require_cuda()
    if data.dtype == dtype:
        return data.clone()          # device-to-device memcpy
    out = torch.empty_like(data, dtype=dtype)
    check(lib.spb_cast(dt.spb_cast_dtype(data.dtype), dt.spb_cast_dtype(dtype), ptr(data), ptr(out), data.numel(),
                       stream()))
    return out


def complex_split(c):
    """complex device tensor (any shape, contiguous) -> (re, im) real tensors of the same shape"""
    require_cuda()
    part = dt.COMPLEX_PART[c.dtype]
    re = torch.empty(c.shape, dtype=part, device=c.device)
    im = torch.empty(c.shape, dtype=part, device=c.device)
    check(lib.spb_complex_split(dt.spb_dtype(part), ptr(c), c.numel(), ptr(re), ptr(im), stream()))
    return re, im


def complex_join(re, im):
    require_cuda()
    assert re.dtype == im.dtype and re.shape == im.shape
    ctype = {torch.float32: torch.complex64, torch.float64: torch.complex128}[re.dtype]
    c = torch.empty(re.shape, dtype=ctype, device=re.device)
    check(lib.spb_complex_join(dt.spb_dtype(re.dtype), ptr(re), ptr(im), re.numel(), ptr(c), stream()))
    return c


def scatter_dense(idx, val, size):
    require_cuda()
    dense = torch.zeros(int(size), dtype=val.dtype, device=val.device)   # cudaMemset: plumbing
    val = _bits(val)
    check(lib.spb_scatter_dense(dt.spb_dtype(val.dtype), ptr(idx), ptr(val), idx.numel(), ptr(dense), stream()))
    return dense


def scatter_into(dense, idx, val):
    """dense[idx[i]] = val[i] in place (dense is a flat device tensor of val's dtype)."""
    require_cuda()
    val = _bits(val)
    check(lib.spb_scatter_dense(dt.spb_dtype(val.dtype), ptr(idx), ptr(val), idx.numel(), ptr(dense), stream()))
    return dense


def gather(src, pos):
    require_cuda()
    out = torch.empty(pos.numel(), dtype=src.dtype, device=src.device)
    check(lib.spb_gather(dt.spb_dtype(_bits(src).dtype), ptr(src), ptr(pos), pos.numel(), ptr(out), stream()))
    return out


def lower_bound(sorted_idx, query, want_found=True):
    require_cuda()
    pos = torch.empty(query.numel(), dtype=torch.int64, device=query.device)
    found = torch.empty(query.numel(), dtype=torch.bool, device=query.device) if want_found else None
    check(lib.spb_lower_bound(ptr(sorted_idx), sorted_idx.numel(), ptr(query), query.numel(), ptr(pos), ptr(found),
                              stream()))
    return pos, found


def select_nonzero(dense_flat):
    require_cuda()
    if dense_flat.dtype in dt.COMPUTE:
        idx, val = select_nonzero(_wide(dense_flat))
        return idx, _back(val, dense_flat.dtype)
    n = dense_flat.numel()
    idx = torch.empty(n, dtype=torch.int64, device=dense_flat.device)
    val = torch.empty(n, dtype=dense_flat.dtype, device=dense_flat.device)
    cnt = torch.empty(1, dtype=torch.int64, device=dense_flat.device)
    check(lib.spb_select_nonzero(dt.spb_dtype(dense_flat.dtype), ptr(dense_flat), n, ptr(idx), ptr(val), ptr(cnt),
                                 workspace(), stream()))
    k = _count(cnt)
    return idx[:k].clone(), val[:k].clone()


def sum_all(data):
    """Returns a python float (float dtypes, float64 accumulate) or int."""
    require_cuda()
    data = _wide(data)
    is_float = data.dtype.is_floating_point
    out = torch.empty(1, dtype=torch.float64 if is_float else torch.int64, device=data.device)
    check(lib.spb_sum_all(dt.spb_dtype(data.dtype), ptr(data), data.numel(), ptr(out), workspace(), stream()))
    return out.item()


def select_pairs(idx, val, keep):
    require_cuda()
    if val.dtype in dt.BITS:
        oi, ov = select_pairs(idx, _bits(val), keep)
        return oi, ov.view(val.dtype)
    n = idx.numel()
    out_idx = torch.empty(n, dtype=torch.int64, device=idx.device)
    out_val = torch.empty(n, dtype=val.dtype, device=idx.device)
    cnt = torch.empty(1, dtype=torch.int64, device=idx.device)
    check(lib.spb_select_flagged(dt.spb_dtype(val.dtype), ptr(idx), ptr(val), ptr(keep), n, ptr(out_idx),
                                 ptr(out_val), ptr(cnt), workspace(), stream()))
    k = _count(cnt)
    return _trim(out_idx, k), _trim(out_val, k)


def min_max(data, which):
    require_cuda()
    if data.numel() == 0:
        raise ValueError("zero-size array to reduction operation %s which has no identity" % which)
    if data.dtype in dt.COMPUTE:
        return dt.np_dtype(data.dtype).type(min_max(_wide(data), which))
    out = torch.empty(1, dtype=data.dtype, device=data.device)
    check(lib.spb_min_max(dt.spb_dtype(data.dtype), ptr(data), data.numel(), int(which == "max"), ptr(out),
                          workspace(), stream()))
    return out.cpu().numpy()[0]


# ---- section 4: reductions over sorted keys ---------------------------------------------------

def reduce_by_key(keys, vals, divisor=1):
    """Sorted keys -> (unique keys // divisor, float64 sums)."""
    require_cuda()
    vals = _wide(vals)
    n = keys.numel()
    out_keys = torch.empty(n, dtype=torch.int64, device=keys.device)
    out_sums = torch.empty(n, dtype=torch.float64, device=keys.device)
    cnt = torch.empty(1, dtype=torch.int64, device=keys.device)
    check(lib.spb_reduce_by_key(dt.spb_dtype(vals.dtype), ptr(keys), ptr(vals), n, int(divisor), ptr(out_keys),
                                ptr(out_sums), ptr(cnt), workspace(), stream()))
    k = _count(cnt)
    return _trim(out_keys, k), _trim(out_sums, k)


def spmv(idx, val, ncols, nrows, x, vdtype):
    """Dense y (dtype vdtype) = A @ x for A given by sorted flat indices."""
    require_cuda()
    if vdtype not in (torch.float32, torch.float64):
        vdtype = torch.float64
    if val.dtype != vdtype:
        val = cast(val, vdtype)
    if x.dtype != vdtype:
        x = cast(x, vdtype)
    y = torch.empty(int(nrows), dtype=torch.float64, device=idx.device)
    check(lib.spb_spmv(dt.spb_dtype(vdtype), ptr(idx), ptr(val), idx.numel(), int(ncols), int(nrows), ptr(x), ptr(y),
                       workspace(), stream()))
    return y if vdtype == torch.float64 else cast(y, vdtype)


def spgemm(a_idx, a_val, m, k, b_idx, b_val, n):
    if a_val.dtype in dt.COMPUTE:
        a_val, b_val = _wide(a_val), _wide(b_val)     # (the caller casts the result to the promoted dtype)
    return _spgemm(a_idx, a_val, m, k, b_idx, b_val, n)


def _spgemm(a_idx, a_val, m, k, b_idx, b_val, n):
    """C = A (m x k) . B (k x n) on sorted flat indices -> canonical (indices, values) of C (m x n):
    expand / sort / compress, the O(products) form of flat.py:545-561.  Values share one dtype."""
    require_cuda()
    assert a_val.dtype == b_val.dtype
    dev = a_idx.device
    na = a_idx.numel()
    empty = (torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=a_val.dtype, device=dev))
    if na == 0 or b_idx.numel() == 0:
        return empty
    # rows of B are contiguous runs of its sorted flat indices (the queries j * n are index plumbing)
    queries = torch.arange(int(k) + 1, device=dev, dtype=torch.int64) * int(n)
    row_ptr, _ = lower_bound(b_idx, queries, want_found=False)
    lens = torch.empty(na, dtype=torch.int64, device=dev)
    check(lib.spb_spgemm_rowlen(ptr(a_idx), na, int(k), ptr(row_ptr), ptr(lens), stream()))
    offs = torch.empty(na, dtype=torch.int64, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    check(lib.spb_exclusive_scan_i64(ptr(lens), na, ptr(offs), ptr(total), workspace(), stream()))
    nprod = _count(total)
    if nprod == 0:
        return empty
    p_idx = torch.empty(nprod, dtype=torch.int64, device=dev)
    p_val = torch.empty(nprod, dtype=a_val.dtype, device=dev)
    check(lib.spb_spgemm_expand(dt.spb_dtype(a_val.dtype), ptr(a_idx), ptr(a_val), na, int(k), ptr(offs), ptr(row_ptr),
                                ptr(b_idx), ptr(b_val), int(n), nprod, ptr(p_idx), ptr(p_val), stream()))
    return canonicalize(p_idx, p_val, max_key=int(m) * int(n) - 1)


# ---- section 5: re-key + radix sort ------------------------------------------------------------

def rekey_sort(idx, val, in_strides, out_strides, sort_divisor, sort_bits):
    """(new keys sorted, values permuted alike); inputs untouched."""
    require_cuda()
    if val is not None and val.dtype in dt.BITS:
        oi, ov = rekey_sort(idx, _bits(val), in_strides, out_strides, sort_divisor, sort_bits)
        return oi, ov.view(val.dtype)
    from . import _plan
    n = idx.numel()
    out_idx = torch.empty_like(idx)
    out_val = torch.empty_like(val) if val is not None else None
    need_tmp = sort_bits > 8
    tmp_idx = torch.empty_like(idx) if need_tmp else None
    tmp_val = torch.empty_like(val) if (need_tmp and val is not None) else None
    ndim = len(in_strides)
    ins = (ctypes.c_int64 * max(ndim, 1))(*in_strides)
    outs = (ctypes.c_int64 * max(ndim, 1))(*out_strides)
    check(lib.spb_rekey_sort(dt.spb_dtype(val.dtype) if val is not None else dt.SPB_I64, ptr(idx), ptr(val), n, ndim,
                             ctypes.cast(ins, ctypes.c_void_p), ctypes.cast(outs, ctypes.c_void_p),
                             int(sort_divisor), int(sort_bits), ptr(out_idx), ptr(out_val), ptr(tmp_idx),
                             ptr(tmp_val), workspace(), stream()))
    return out_idx, out_val


def transpose(idx, val, shape, axes):
    from . import _plan
    in_s, out_s, div, bits = _plan.transpose_plan(shape, axes)
    return rekey_sort(idx, val, in_s, out_s, div, bits)


DENSE_SUM_MAX_CELLS = 1 << 26      # 512 MiB of float64 accumulators


@functools.lru_cache(maxsize=8)
def _l2_resident_cells(device_index):
    """Cells of the dense accumulator slab (8-byte sum + 1 flag byte each) that fit in HALF of this device's
    L2: below this the dense path is taken whatever the element count -- every atomic then hits L2 and the
    compaction pass reads the flags from it."""
    l2 = 126 << 20                                  # B200; the planning helpers also run without a device
    if device_index >= 0 and torch.cuda.is_available():
        props = torch.cuda.get_device_properties(device_index)
        l2 = int(getattr(props, "L2_cache_size", 0)) or l2
    return max(l2 // 2 // 9, 1 << 16)


def _cur_dev():
    return torch._C._cuda_getDevice() if torch.cuda.is_available() else -1


def dense_sum_cells(shape, axis):
    """Cell count of sum(axis)'s result if the dense-accumulator path applies whatever nnz is, else 0."""
    from . import _plan
    cells = _plan.prod(shape) // max(int(shape[axis]), 1)
    return cells if 0 < cells <= _l2_resident_cells(_cur_dev()) else 0


def _use_dense(n, cells, last, n_dev):
    if n_dev is not None:
        return True
    if not n or cells > DENSE_SUM_MAX_CELLS:
        return False
    if last:
        # the input is already grouped: beyond an L2-resident slab the one-pass segmented reduce wins
        # (config-2 operand, 16.7 M cells: 198 us against 213 us)
        return cells <= _l2_resident_cells(_cur_dev())
    return cells <= 64 * n + _l2_resident_cells(_cur_dev())


def sum_axis(idx, val, shape, axis, n_dev=None, window=None):
    """(new sorted unique flat indices, float64 sums, device count or None) -- flat.py:617-625.
    Whenever the reduced array is small enough to hold densely (and not much sparser than the input):
    warp-aggregated float64 atomics into a dense accumulator + one compaction pass, no sort, any axis.
    Otherwise: last axis -> segmented reduce over the already sorted keys; other axes -> re-key +
    radix sort + segmented reduce.  The dense path leaves its element count on the device
    (worst-case-sized outputs); n_dev is the device count of a producer that was not read back
    either (only valid when the dense path is certain: dense_sum_cells()).
    window = (cell_lo, cell_hi): the caller knows that every output cell lies in that range of the
    reduced array (a range shard of a sharded array): accumulators are only held for the window."""
    from . import _plan
    plan = _plan.axis_sum_plan(shape, axis)
    n = idx.numel()
    total_cells = _plan.prod(shape) // max(int(shape[axis]), 1)
    lo, hi = (0, total_cells) if window is None else (int(window[0]), int(window[1]))
    cells = hi - lo
    if n_dev is None and window is None and n:
        # one streaming pass over whole slabs of the leading coordinates (csrc/slab_sum.cu) when the library
        # says the shape and fill allow it: neither a dense accumulator array nor a sort
        slab = slab_geometry(shape, axis)
        if slab is not None and _prefer_slab(n, total_cells, slab[2]) and lib.spb_axis_sum_slab_ok(n, *slab):
            return axis_sum_slab(idx, val, slab, min(n, total_cells))
    if plan[0] == "last":
        if _use_dense(n, cells, True, n_dev):
            # the elements of an output cell are contiguous: warp-aggregated atomics into the dense accumulator
            return axis_sum_dense(idx, val, [plan[1], 1], [1, 0], lo, hi, n_dev)
        return reduce_by_key(idx, val, plan[1]) + (None,)
    _, in_s, out_s, bits = plan
    if _use_dense(n, cells, False, n_dev):
        return axis_sum_dense(idx, val, in_s, out_s, lo, hi, n_dev)
    assert n_dev is None
    keys, vals = rekey_sort(idx, val, in_s, out_s, 1, bits)
    return reduce_by_key(keys, vals, 1) + (None,)


def _prefer_slab(n, out_cells, suffix):
    """Measured on the config-2 / config-4 operands: with the reduced axis last the slab kernels win unless the
    runs are long (>= 8 stored elements per output cell: the lane-pre-reducing dense kernel is faster there);
    for other axes they win once the dense accumulators would not stay in L2."""
    if suffix == 1:
        return n < 8 * out_cells
    return out_cells > _l2_resident_cells(_cur_dev())


def slab_geometry(shape, axis):
    """(prod(shape), shape[axis] * suffix, suffix) with suffix = prod(shape[axis+1:]); None for empty shapes."""
    from . import _plan
    total = _plan.prod(shape)
    suffix = _plan.prod(shape[axis + 1:])
    if total < 1 or total >= 1 << 62:
        return None
    return total, int(shape[axis]) * suffix, suffix


def axis_sum_slab(idx, val, slab, cap):
    require_cuda()
    val = _wide(val)
    out_idx = torch.empty(cap, dtype=torch.int64, device=idx.device)
    out_sums = torch.empty(cap, dtype=torch.float64, device=idx.device)
    cnt = torch.empty(1, dtype=torch.int64, device=idx.device)
    check(lib.spb_axis_sum_slab(dt.spb_dtype(val.dtype), ptr(idx), ptr(val), idx.numel(), slab[0], slab[1], slab[2],
                                ptr(out_idx), ptr(out_sums), ptr(cnt), workspace(), stream()))
    return out_idx, out_sums, cnt


@functools.lru_cache(maxsize=512)
def _stride_args(in_strides, out_strides):
    """(ndim, void* in, void* out) of two stride tuples as host int64 arrays; the arrays are kept alive
    by the cache entry (the C side reads them during the call only)."""
    ndim = len(in_strides)
    ins = (ctypes.c_int64 * ndim)(*in_strides)
    outs = (ctypes.c_int64 * ndim)(*out_strides)
    return ndim, ctypes.cast(ins, ctypes.c_void_p), ctypes.cast(outs, ctypes.c_void_p), (ins, outs)


def axis_sum_dense(idx, val, in_strides, out_strides, cell_lo, cell_hi, n_dev=None):
    require_cuda()
    val = _wide(val)
    n = idx.numel()
    cap = min(n, cell_hi - cell_lo)
    out_idx = torch.empty(cap, dtype=torch.int64, device=idx.device)
    out_sums = torch.empty(cap, dtype=torch.float64, device=idx.device)
    cnt = torch.empty(1, dtype=torch.int64, device=idx.device)
    ndim, ins, outs, _keep = _stride_args(tuple(in_strides), tuple(out_strides))
    check(lib.spb_axis_sum_window(dt.spb_dtype(val.dtype), ptr(idx), ptr(val), n, ptr(n_dev), ndim, ins, outs,
                                  int(cell_lo), int(cell_hi), ptr(out_idx), ptr(out_sums), ptr(cnt), workspace(),
                                  stream()))
    return out_idx, out_sums, cnt


def axis_accumulate(idx, val, in_strides, out_strides, cells):
    val = _wide(val)
    """Local half of a sharded dense axis sum: (float64 accumulators [cells], flag bytes [cells padded
    to 64]) holding this shard's contributions; to be all-reduced (SUM / MAX) across ranks."""
    require_cuda()
    cells = int(cells)
    acc = torch.zeros((cells + 15) // 16 * 16, dtype=torch.float64, device=idx.device)   # padded: 16-byte pair loads
    touched = torch.zeros((cells + 63) // 64 * 64, dtype=torch.uint8, device=idx.device)
    ndim, ins, outs, _keep = _stride_args(tuple(in_strides), tuple(out_strides))
    check(lib.spb_axis_accumulate(dt.spb_dtype(val.dtype), ptr(idx), ptr(val), idx.numel(), ndim, ins, outs, cells,
                                  ptr(acc), ptr(touched), stream()))
    return acc, touched


def emit_touched(acc, touched, lo, hi):
    """Touched cells of [lo, hi) as (sorted absolute cell numbers, float64 sums); lo % 64 == 0."""
    require_cuda()
    lo, hi = int(lo), int(hi)
    if hi <= lo:
        return (torch.empty(0, dtype=torch.int64, device=acc.device),
                torch.empty(0, dtype=torch.float64, device=acc.device))
    out_idx = torch.empty(hi - lo, dtype=torch.int64, device=acc.device)
    out_sums = torch.empty(hi - lo, dtype=torch.float64, device=acc.device)
    cnt = torch.empty(1, dtype=torch.int64, device=acc.device)
    check(lib.spb_emit_touched(ptr(acc), ptr(touched), lo, hi, ptr(out_idx), ptr(out_sums), ptr(cnt), workspace(),
                               stream()))
    k = _count(cnt)
    return _trim(out_idx, k), _trim(out_sums, k)


def canonicalize(idx, val, max_key=None):
    """Sort by index and sum duplicates in float64, cast back (flat.py:32-35)."""
    if idx.numel() == 0:
        return idx, val
    if max_key is None:
        max_key = int(min_max(idx, "max"))
        if int(min_max(idx, "min")) < 0:
            raise ValueError("negative flat index")
    bits = max(int(max_key), 1).bit_length()
    keys, vals = rekey_sort(idx, val, [], [], 1, bits)
    out_idx, sums = reduce_by_key(keys, vals, 1)
    return out_idx, cast(sums, val.dtype)


# ---- section 6: __getitem__ ------------------------------------------------------------------------

def getitem_box(idx, val, shape, ranges, new_shape):
    """ints + slices: one pass over the stored elements inside the box's flat-index span."""
    require_cuda()
    if val.dtype in dt.BITS:
        oi, ov = getitem_box(idx, _bits(val), shape, ranges, new_shape)
        return oi, ov.view(val.dtype)
    ranges = np.ascontiguousarray(ranges, dtype=np.int64)
    shape_arr = np.ascontiguousarray(shape, dtype=np.int64)
    strides = np.ones(len(shape), dtype=object)
    for ax in range(len(shape) - 2, -1, -1):
        strides[ax] = strides[ax + 1] * int(shape[ax + 1])
    lens = [len_range(*r) for r in ranges]
    lo_flat = sum(int(r[0]) * int(s) for r, s in zip(ranges, strides))
    hi_flat = sum((int(r[0]) + (l - 1) * int(r[2])) * int(s) for r, l, s in zip(ranges, lens, strides))
    # restrict the scan to the stored elements in [lo_flat, hi_flat]: two binary searches
    q = torch.tensor([lo_flat, hi_flat + 1], dtype=torch.int64, device=idx.device)
    pos, _ = lower_bound(idx, q, want_found=False)
    lo, hi = (int(x) for x in pos.cpu().tolist())
    sub_idx, sub_val = idx[lo:hi], val[lo:hi]
    n = hi - lo
    out_idx = torch.empty(n, dtype=torch.int64, device=idx.device)
    out_val = torch.empty(n, dtype=val.dtype, device=idx.device)
    cnt = torch.empty(1, dtype=torch.int64, device=idx.device)
    check(lib.spb_getitem_box(dt.spb_dtype(val.dtype), ptr(sub_idx), ptr(sub_val), n, len(shape),
                              shape_arr.ctypes.data_as(ctypes.c_void_p), ranges.ctypes.data_as(ctypes.c_void_p),
                              ptr(out_idx), ptr(out_val), ptr(cnt), workspace(), stream()))
    k = _count(cnt)
    return out_idx[:k].clone(), out_val[:k].clone()


def outer_sum(offsets, new_shape, device):
    """Flat query of a mixed slice/array index: outer sum of per-axis offset lists."""
    require_cuda()
    lens = np.asarray([len(o) for o in offsets], dtype=np.int64)
    cat = torch.from_numpy(np.concatenate([np.asarray(o, dtype=np.int64) for o in offsets])).to(device)
    n = int(np.prod(lens, dtype=object))
    out = torch.empty(n, dtype=torch.int64, device=device)
    check(lib.spb_outer_sum(ptr(cat), len(offsets), lens.ctypes.data_as(ctypes.c_void_p), ptr(out), n, stream()))
    return out


def lookup_compact(sorted_idx, val, query):
    require_cuda()
    if val.dtype in dt.BITS:
        oi, ov = lookup_compact(sorted_idx, _bits(val), query)
        return oi, ov.view(val.dtype)
    nq = query.numel()
    out_idx = torch.empty(nq, dtype=torch.int64, device=query.device)
    out_val = torch.empty(nq, dtype=val.dtype, device=query.device)
    cnt = torch.empty(1, dtype=torch.int64, device=query.device)
    check(lib.spb_lookup_compact(dt.spb_dtype(val.dtype), ptr(sorted_idx), ptr(val), sorted_idx.numel(), ptr(query),
                                 nq, ptr(out_idx), ptr(out_val), ptr(cnt), workspace(), stream()))
    k = _count(cnt)
    return out_idx[:k].clone(), out_val[:k].clone()


# ---- section 7: sparse broadcasting, dense operands ---------------------------------------------------

def binary_map(op, a, b):
    """Elementwise a (op) b on two aligned device vectors of one dtype."""
    require_cuda()
    assert a.dtype == b.dtype and a.numel() == b.numel()
    if a.dtype in dt.COMPUTE:
        out = binary_map(op, _wide(a), _wide(b))
        return out if out.dtype == torch.bool else _back(out, a.dtype)
    out = torch.empty_like(a, dtype=torch.bool if op in dt.COMPARISONS else a.dtype)
    check(lib.spb_binary_map(dt.BINOPS[op], dt.spb_dtype(a.dtype), ptr(a), ptr(b), ptr(out), a.numel(), stream()))
    return out


def broadcast_expand(idx, val, in_shape, out_shape):
    """Unsorted (index, value) pairs of the array broadcast from in_shape (left-padded) to out_shape."""
    require_cuda()
    if val.dtype in dt.BITS:
        oi, ov = broadcast_expand(idx, _bits(val), in_shape, out_shape)
        return oi, ov.view(val.dtype)
    ins = np.ascontiguousarray(in_shape, dtype=np.int64)
    outs = np.ascontiguousarray(out_shape, dtype=np.int64)
    reps = int(np.prod([o for i, o in zip(ins, outs) if i == 1 and o > 1], dtype=object))
    n_out = idx.numel() * reps
    out_idx = torch.empty(n_out, dtype=torch.int64, device=idx.device)
    out_val = torch.empty(n_out, dtype=val.dtype, device=idx.device)
    check(lib.spb_broadcast_expand(dt.spb_dtype(val.dtype), ptr(idx), ptr(val), idx.numel(), len(ins),
                                   ins.ctypes.data_as(ctypes.c_void_p), outs.ctypes.data_as(ctypes.c_void_p),
                                   ptr(out_idx), ptr(out_val), stream()))
    return out_idx, out_val
