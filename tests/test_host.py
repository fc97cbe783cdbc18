"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header
declares, and the host-side logic (index preparation, re-key / sort planning, dtype
tables, exact division constants) is right.  No compute kernels are called."""
import itertools
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sparray_b200 import _lib
    with open(os.path.join(REPO, "include", "sparray_b200.h")) as fh:
        header = fh.read()
    declared = set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), "libsparray_b200.so does not export %s" % name
        assert name in _lib.PROTOTYPES, "no ctypes prototype for %s" % name
    assert set(_lib.PROTOTYPES) <= declared
    assert b"sm_100a" in _lib.lib.spb_version()
    assert _lib.lib.spb_error_string(10003).startswith(b"no CUDA device")


def test_no_cpu_fallback_without_gpu():
    import torch
    import sparray_b200 as sp
    from sparray_b200._lib import SparrayB200Error
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(SparrayB200Error):
        sp.FlatSparray([1, 2], [1.0, 2.0], shape=(4,))
    from sparray_b200 import compat
    with pytest.raises(SparrayB200Error):
        compat.union1d_sorted(np.array([1]), np.array([2]))


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "sparray_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "flat_oracle" not in src and "merge_oracle" not in src and "oracle/" not in src, f


def test_len_range_matches_python_range():
    from sparray_b200 import compat, _indexing
    for t in itertools.product([-7, -1, 0, 3, 10], [-5, 0, 4, 11], [-3, -1, 1, 2, 5]):
        assert compat.len_range(*t) == len(range(*t)) == _indexing.len_range(*t)


def test_fastdiv_exact():
    from sparray_b200._lib import lib
    rng = np.random.default_rng(0)
    divisors = [1, 2, 3, 5, 7, 10, 64, 1000, 10 ** 6, 10 ** 6 + 3, 2 ** 31 - 1, 2 ** 32, 2 ** 32 + 1, 10 ** 12,
                2 ** 62 - 57, 2 ** 62] + [int(x) for x in rng.integers(1, 2 ** 62, 40)]
    for d in divisors:
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 63 - 1, 2 ** 63 - d] + \
             [int(x) for x in rng.integers(0, 2 ** 63 - 1, 60)]
        for n in ns:
            if 0 <= n < 2 ** 63:
                assert lib.spb_debug_fastdiv(n, d) == n // d, (n, d)


def _emulate_rekey(idx, in_s, out_s):
    rem, acc = idx.copy(), np.zeros_like(idx)
    for g, (a, b) in enumerate(zip(in_s, out_s)):
        if g < len(in_s) - 1:
            q = rem // a
            rem = rem - q * a
            acc += q * b
        else:
            acc += rem * b
    return acc


@pytest.mark.parametrize("shape", [(7, 5), (3, 4, 5), (2, 3, 4, 5), (6, 1, 3), (16, 8, 4), (4, 4, 4), (1, 5, 1, 6),
                                   (3, 1, 1, 4, 2), (9,)])
def test_transpose_and_axis_sum_plans(shape):
    from sparray_b200 import _plan as pl
    size = int(np.prod(shape))
    idx = np.arange(size)
    multi = np.unravel_index(idx, shape)
    for axes in itertools.permutations(range(len(shape))):
        in_s, out_s, div, bits = pl.transpose_plan(shape, axes)
        new = _emulate_rekey(idx, in_s, out_s)
        want = np.ravel_multi_index(tuple(multi[a] for a in axes), tuple(shape[a] for a in axes))
        assert np.array_equal(new, want)
        assert len(in_s) <= len(shape) and in_s[-1] == 1
        sk = new // div
        assert sk.max() < (1 << bits) or (bits == 0 and sk.max() == 0)
        order = np.argsort(sk & ((1 << bits) - 1), kind="stable")     # what the LSD passes do
        assert np.all(np.diff(new[order]) > 0), (shape, axes)
    for ax in range(len(shape)):
        keep = [a for a in range(len(shape)) if a != ax]
        if not keep:
            continue
        want = np.ravel_multi_index(tuple(multi[a] for a in keep), tuple(shape[a] for a in keep))
        plan = pl.axis_sum_plan(shape, ax)
        if plan[0] == "last":
            assert np.array_equal(idx // plan[1], want)
        else:
            new = _emulate_rekey(idx, plan[1], plan[2])
            assert np.array_equal(new, want) and new.max() < (1 << plan[3])


def test_plans_for_baseline_configs():
    from sparray_b200 import _plan as pl
    # config 3: swapaxes(0,2) of (4096,4096,1024): only the 22 bits of (z, y) are sorted, not 34
    assert pl.transpose_plan((4096, 4096, 1024), (2, 1, 0))[2:] == (4096, 22)
    # config 2 last-axis sum needs no sort; axis 2 sorts the 22-bit output key
    assert pl.axis_sum_plan((256, 256, 256, 64), 3) == ("last", 64)
    assert pl.axis_sum_plan((256, 256, 256, 64), 2)[0] == "sort" and pl.axis_sum_plan((256, 256, 256, 64), 2)[3] == 22
    # 2-D transpose sorts on the column only
    assert pl.transpose_plan((10 ** 6, 10 ** 6), (1, 0))[2:] == (10 ** 6, 20)


def test_prepare_indices_mirrors_reference_errors():
    from sparray_b200 import _indexing as ix
    shape = (6, 5, 7)
    items, kind = ix.prepare_indices(3, shape)
    assert items == (3, slice(None), slice(None)) and kind == ix.INTEGER_INDEX_MASK | ix.EMPTY_SLICE_INDEX_MASK
    items, kind = ix.prepare_indices((Ellipsis, -1), shape)
    assert items == (slice(None), slice(None), 6)
    items, kind = ix.prepare_indices((1, 2, 3), shape)
    assert kind == ix.INTEGER_INDEX_MASK
    items, kind = ix.prepare_indices(slice(None), shape)
    assert kind == ix.EMPTY_SLICE_INDEX_MASK
    items, kind = ix.prepare_indices(([1, 2], slice(1, 3)), shape)
    assert kind == ix.ARRAY_INDEX_MASK | ix.SLICE_INDEX_MASK | ix.EMPTY_SLICE_INDEX_MASK
    items, kind = ix.prepare_indices((np.array([True, False, True, False, False, True]),), shape)
    assert np.array_equal(items[0], [0, 2, 5])
    with pytest.raises(IndexError):
        ix.prepare_indices(6, shape)
    with pytest.raises(IndexError):
        ix.prepare_indices((1, 2, 3, 4), shape)
    with pytest.raises(IndexError):
        ix.prepare_indices((Ellipsis, 1, Ellipsis), shape)
    with pytest.raises(NotImplementedError):
        ix.prepare_indices((np.zeros((2, 2), dtype=int),), shape)


def test_indices_to_ranges_and_fancy_plan_against_numpy():
    from sparray_b200 import _indexing as ix
    shape = (6, 5, 7)
    grid = np.arange(int(np.prod(shape))).reshape(shape)
    for index in [(slice(1, 5), slice(None), slice(2, 6)), (slice(0, 6, 2), 3, slice(None, None, 2)), (2,)]:
        items, _ = ix.prepare_indices(index, shape)
        ranges, new_shape = ix.indices_to_ranges(items, shape)
        want = grid[index]
        assert list(want.shape) == new_shape
        starts = ranges[:, 0]
        assert grid[tuple(starts)] == want.ravel()[0]
    for index in [([1, 3], slice(None), slice(2, 5)), ([4, 0, 3], 2, slice(None)), ([0, 2, 5], [1, 1, 4], [6, 0, 3]),
                  (slice(None), [4, 2], 5)]:
        items, _ = ix.prepare_indices(index, shape)
        offsets, new_shape = ix.fancy_plan(items, shape)
        q = offsets[0]
        for o in offsets[1:]:
            q = np.add.outer(q, o).ravel()
        want = grid[index]
        assert tuple(new_shape) == want.shape, index
        assert np.array_equal(q.reshape(new_shape), want), index


def test_dtype_tables():
    import torch
    from sparray_b200 import _dtypes as dt
    assert dt.promote(torch.float32, torch.float64) == torch.float64
    assert dt.promote(torch.int64, torch.float32) == torch.float64
    assert dt.promote(torch.int32, torch.float32) == torch.float64
    assert dt.np_dtype(torch.bool) == np.bool_
    with pytest.raises(TypeError):
        dt.spb_dtype(torch.complex64)
    with open(os.path.join(REPO, "include", "sparray_b200.h")) as fh:
        header = fh.read()
    enum = re.search(r"enum spb_unop \{(.*?)\};", header, re.S).group(1)
    names = [n.strip().split("=")[0].strip() for n in enum.replace("\n", " ").split(",") if n.strip()]
    names = [n for n in names if n != "SPB_UNOP_COUNT"]
    assert [n[len("SPB_U_"):].lower() for n in names] == sorted(dt.UNOPS, key=dt.UNOPS.get)


def test_dense_sum_eligibility_and_cached_plans():
    """Host-side decisions of sum(axis): which shapes may consume a pending (device-side) count, and
    that the cached plans are the plain ones."""
    from sparray_b200 import _kernels as k, _plan
    assert k.dense_sum_cells((256, 256, 256, 64), 2) == 256 * 256 * 64          # 4.2 M cells: dense whatever nnz is
    assert k.dense_sum_cells((256, 256, 256, 64), 3) == 0                       # 16.7 M cells: needs the real count
    assert k.dense_sum_cells((10 ** 6, 10 ** 6), 1) == 10 ** 6
    assert k.dense_sum_cells((4, 5), 0) == 5
    for shape, axis in (((6, 5, 4), 0), ((6, 5, 4), 1), ((6, 5, 4), 2), ((7, 1, 1), 0)):
        assert _plan.axis_sum_plan(shape, axis) == _plan.axis_sum_plan(list(shape), axis)
        assert _plan.axis_sum_plan(shape, axis) is _plan.axis_sum_plan(shape, axis)      # cached
    assert _plan.axis_sum_plan((6, 5, 4), 2) == ("last", 4)
    assert _plan.axis_sum_plan((7, 1, 1), 0) == ("last", 7)
    ndim, ins, outs, keep = k._stride_args((20, 4, 1), (4, 0, 1))
    assert ndim == 3 and list(keep[0]) == [20, 4, 1] and list(keep[1]) == [4, 0, 1]
    assert k._stride_args((20, 4, 1), (4, 0, 1))[1] is ins                       # cached: the arrays stay alive
