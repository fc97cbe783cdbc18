"""CPU tests: pin the oracle (oracle/) to the reference.

1. the reference's own known-answer vectors (sparray/tests/test_compat.py:10-53);
2. every committed golden case (produced by running the reference itself);
3. when /root/reference is present (build container), live randomised
   comparison against the reference on both of its kernel paths.
"""
import os
import sys

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import golden_runner as gr

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import flat_oracle as fo  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_loader  # noqa: E402


class OracleSeam:
    union1d_sorted = staticmethod(fo.union1d_sorted)
    intersect1d_sorted = staticmethod(fo.intersect1d_sorted)
    combine_ranges = staticmethod(fo.combine_ranges)
    len_range = staticmethod(fo.len_range)
    asinput = staticmethod(lambda x: x)


def test_kat_intersect():            # test_compat.py:10-18
    a, b = np.array([0, 1, 4, 6, 8, 9]), np.array([2, 4, 5, 6, 7])
    assert_array_equal(fo.intersect1d_sorted(a, b), [4, 6])
    x, ai, bi = fo.intersect1d_sorted(a, b, return_inds=True)
    assert_array_equal(x, [4, 6])
    assert_array_equal(ai, [2, 3])
    assert_array_equal(bi, [1, 3])


def test_kat_union():                # test_compat.py:20-29
    a, b = np.array([0, 2, 4, 8]), np.array([1, 4, 6, 8])
    assert_array_equal(fo.union1d_sorted(a, b), [0, 1, 2, 4, 6, 8])
    x, lut, ao, bo = fo.union1d_sorted(a, b, return_masks=True)
    assert_array_equal(x, [0, 1, 2, 4, 6, 8])
    assert_array_equal(lut, [0, 1, 0, 2, 1, 2])
    assert_array_equal(ao, [True, True, False, False])
    assert_array_equal(bo, [True, False, True, False])


def test_kat_combine_ranges():       # test_compat.py:38-46
    ranges = np.array([[0, 2, 1], [1, 2, 1], [1, 4, 2]])
    assert_array_equal(fo.combine_ranges(ranges, (2, 3, 4), 4), [5, 7, 17, 19])
    assert_array_equal(fo.combine_ranges(ranges, (2, 3, 4), 2, inner=True), [5, 19])


def test_kat_len_range():            # test_compat.py:48-53
    for t in [(0, 5, 1), (5, 1, 1), (5, 0, -1), (0, 5, -2)]:
        assert fo.len_range(*t) == len(range(*t))


# (the "prog" cases -- SURVEY.md section 8(f) rows -- are outputs of the reference that the PRODUCT is held to
# directly, tests/test_gpu_ops.py::test_golden_f_rows; the oracle restates the section 8(a) path only)
@pytest.mark.parametrize("case", [c for c in gr.cases() if c["op"] != "prog"], ids=lambda c: c["name"])
def test_golden(case):
    got = gr.run_case(case, fo.OracleSparray, OracleSeam)
    gr.check_case(case, got)


def test_reference_kernel_binary_matches_port():
    """oracle/_ref (the reference's Cython kernel, compiled in place) against
    the C port, on random inputs."""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import build_ref
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference tree / Cython here)")
    rng = np.random.default_rng(7)
    for _ in range(20):
        a = np.unique(rng.integers(0, 5000, rng.integers(0, 3000)))
        b = np.unique(rng.integers(0, 5000, rng.integers(0, 3000)))
        for x, y in zip(ref.union1d_sorted(a, b, return_masks=True), fo.union1d_sorted(a, b, return_masks=True)):
            assert_array_equal(np.asarray(x), y)
        for x, y in zip(ref.intersect1d_sorted(a, b, return_inds=True), fo.intersect1d_sorted(a, b, return_inds=True)):
            assert_array_equal(np.asarray(x), y)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("use_cython", [True, False])
def test_live_against_reference(use_cython):
    ref = ref_loader.load_reference(use_cython=use_cython)
    rng = np.random.default_rng(11)
    for shape in [(50, 40), (9, 8, 7, 3), (300,)]:
        size = int(np.prod(shape))
        for dt in (np.float64, np.float32):
            ops = []
            for seed in (1, 2):
                idx = np.unique(rng.integers(0, size, size // 3))
                val = rng.standard_normal(len(idx)).astype(dt)
                ops.append((idx, val))
            ra, rb = (ref.FlatSparray(i, v, shape=shape, is_canonical=True) for i, v in ops)
            oa, ob = (fo.OracleSparray(i, v, shape=shape, is_canonical=True) for i, v in ops)
            for f in (lambda a, b: a + b, lambda a, b: a - b, lambda a, b: a * b, lambda a, b: a < b,
                      lambda a, b: a.minimum(b), lambda a, b: a.maximum(b), lambda a, b: a != b):
                r, o = f(ra, rb), f(oa, ob)
                assert_array_equal(r.indices, o.indices)
                assert_array_equal(r.data, o.data)
                assert r.data.dtype == o.data.dtype
            for ax in range(len(shape)):
                r, o = ra.sum(axis=ax), oa.sum(axis=ax)
                if hasattr(r, "indices"):
                    assert_array_equal(r.indices, o.indices)
                    assert_array_equal(r.data, o.data)
                else:
                    assert r == o
            if len(shape) > 1 and tuple(reversed(shape)) != shape:
                r, o = ra.T, oa.T
                assert_array_equal(r.indices, o.indices)
                assert_array_equal(r.data, o.data)
            if len(shape) == 2:
                x = rng.standard_normal(shape[1]).astype(dt)
                np.testing.assert_allclose(ra.dot(x), oa.dot(x), rtol=1e-5 if dt == np.float32 else 1e-12)
