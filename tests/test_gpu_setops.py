"""GPU parity: merge-path union / intersection kernels (the compat seam and the
fused sparse(op)sparse path) against the golden fixtures and the CPU oracle.
Bit-exact: indices, lut/masks/positions, and every single elementwise value op."""
import numpy as np
import pytest

import golden_runner as gr
from gpu_util import assert_sparse_same, fo, rand_operand, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import torch
    import sparray_b200
    from sparray_b200 import compat

    class Seam:
        union1d_sorted = staticmethod(compat.union1d_sorted)
        intersect1d_sorted = staticmethod(compat.intersect1d_sorted)
        combine_ranges = staticmethod(compat.combine_ranges)
        len_range = staticmethod(compat.len_range)
        asinput = staticmethod(lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64)).cuda())

    sparray_b200.Seam = Seam
    return sparray_b200


@pytest.mark.parametrize("case", gr.cases("seam_merge") + gr.cases("seam_combine_ranges") + gr.cases("seam_len_range"),
                         ids=lambda c: c["name"])
def test_golden_seam(sp, case):
    gr.check_case(case, gr.run_case(case, sp.FlatSparray, sp.Seam))


@pytest.mark.parametrize("case", gr.cases("binop"), ids=lambda c: c["name"])
def test_golden_binop(sp, case):
    gr.check_case(case, gr.run_case(case, sp.FlatSparray, sp.Seam))


def test_kat_reference_vectors(sp):
    """sparray/tests/test_compat.py:10-46, verbatim answers."""
    from sparray_b200 import compat as sc
    a, b = np.array([0, 1, 4, 6, 8, 9]), np.array([2, 4, 5, 6, 7])
    x, ai, bi = sc.intersect1d_sorted(a, b, return_inds=True)
    assert to_np(x).tolist() == [4, 6] and to_np(ai).tolist() == [2, 3] and to_np(bi).tolist() == [1, 3]
    a, b = np.array([0, 2, 4, 8]), np.array([1, 4, 6, 8])
    x, lut, ao, bo = sc.union1d_sorted(a, b, return_masks=True)
    assert to_np(x).tolist() == [0, 1, 2, 4, 6, 8]
    assert to_np(lut).tolist() == [0, 1, 0, 2, 1, 2]
    assert to_np(ao).tolist() == [True, True, False, False]
    assert to_np(bo).tolist() == [True, False, True, False]
    ranges = np.array([[0, 2, 1], [1, 2, 1], [1, 4, 2]])
    assert to_np(sc.combine_ranges(ranges, (2, 3, 4), 4)).tolist() == [5, 7, 17, 19]
    assert to_np(sc.combine_ranges(ranges, (2, 3, 4), 2, inner=True)).tolist() == [5, 19]


def test_seam_calling_convention(sp):
    """Same refusals as the Cython typed memoryviews."""
    from sparray_b200 import compat as sc
    with pytest.raises(ValueError):
        sc.union1d_sorted(np.array([1, 2], dtype=np.int32), np.array([1], dtype=np.int64))
    with pytest.raises(ValueError):
        sc.intersect1d_sorted(np.arange(10, dtype=np.int64)[::2], np.array([1], dtype=np.int64))


SIZES = [(0, 0), (0, 5), (7, 0), (1, 1), (2047, 1), (2048, 2048), (2049, 2047), (4096, 1), (5000, 7000),
         (100000, 3), (3, 100000), (300000, 310000)]


@pytest.mark.parametrize("na,nb", SIZES)
def test_seam_random_vs_oracle(sp, na, nb):
    rng = np.random.default_rng(na * 7919 + nb)
    hi = max(4 * (na + nb), 10)
    a = np.unique(rng.integers(0, hi, na)).astype(np.int64)
    b = np.unique(rng.integers(0, hi, nb)).astype(np.int64)
    from sparray_b200 import compat as sc
    for got, want in zip(sc.union1d_sorted(a, b, return_masks=True), fo.union1d_sorted(a, b, return_masks=True)):
        np.testing.assert_array_equal(to_np(got).astype(want.dtype), want)
    for got, want in zip(sc.intersect1d_sorted(a, b, return_inds=True), fo.intersect1d_sorted(a, b, return_inds=True)):
        np.testing.assert_array_equal(to_np(got), want)
    np.testing.assert_array_equal(to_np(sc.union1d_sorted(a, b)), fo.union1d_sorted(a, b))
    np.testing.assert_array_equal(to_np(sc.intersect1d_sorted(a, b)), fo.intersect1d_sorted(a, b))


@pytest.mark.parametrize("overlap", ["identical", "disjoint_blocks", "interleaved", "dense_pairs"])
def test_structured_overlaps(sp, overlap):
    n = 20000
    if overlap == "identical":
        a = b = np.arange(0, 3 * n, 3, dtype=np.int64)
    elif overlap == "disjoint_blocks":
        a, b = np.arange(n, dtype=np.int64), np.arange(n, 2 * n, dtype=np.int64) + 2 ** 40
    elif overlap == "interleaved":
        a, b = np.arange(0, 2 * n, 2, dtype=np.int64), np.arange(1, 2 * n, 2, dtype=np.int64)
    else:   # every pair straddles some thread/tile boundary position
        a = np.arange(n, dtype=np.int64)
        b = np.arange(0, n, 1, dtype=np.int64)[::1].copy()
        b = b[b % 3 != 1]
    rng = np.random.default_rng(5)
    av, bv = rng.standard_normal(len(a)), rng.standard_normal(len(b))
    A = sp.FlatSparray(a, av, shape=(2 ** 41,), is_canonical=True)
    B = sp.FlatSparray(b, bv, shape=(2 ** 41,), is_canonical=True)
    oa = fo.OracleSparray(a, av, shape=(2 ** 41,), is_canonical=True)
    ob = fo.OracleSparray(b, bv, shape=(2 ** 41,), is_canonical=True)
    assert_sparse_same(A + B, oa + ob)
    assert_sparse_same(A * B, oa * ob)
    assert_sparse_same(A - B, oa - ob)
    assert_sparse_same(A <= B, oa <= ob)


DTYPES = [np.float64, np.float32, np.int64, np.int32]
OPS = {"add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b,
       "min": lambda a, b: a.minimum(b), "max": lambda a, b: a.maximum(b),
       "lt": lambda a, b: a < b, "le": lambda a, b: a <= b, "gt": lambda a, b: a > b,
       "ge": lambda a, b: a >= b, "eq": lambda a, b: a == b, "ne": lambda a, b: a != b}


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("shape,nnz", [((1000, 1000), 60000), ((64, 64, 64, 16), 200000), ((2 ** 36,), 50000)])
def test_binops_random_vs_oracle(sp, dtype, shape, nnz):
    rng = np.random.default_rng(nnz + np.dtype(dtype).itemsize)
    ai, av = rand_operand(rng, shape, nnz, dtype)
    bi, bv = rand_operand(rng, shape, nnz + 1234, dtype)
    # force some exact matches with special values
    common = ai[::7]
    bi = np.union1d(bi, common)
    bv = rand_operand(rng, (len(bi) * 2,), len(bi) * 2, dtype)[1][:len(bi)]
    if np.dtype(dtype).kind == "f":
        av[:4] = [np.nan, -0.0, 0.0, np.inf]
        bv[:4] = [1.0, 0.0, -0.0, -np.inf]
    A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
    B = sp.FlatSparray(bi, bv, shape=shape, is_canonical=True)
    oa = fo.OracleSparray(ai, av, shape=shape, is_canonical=True)
    ob = fo.OracleSparray(bi, bv, shape=shape, is_canonical=True)
    with np.errstate(all="ignore"):
        for name, f in OPS.items():
            assert_sparse_same(f(A, B), f(oa, ob))


def test_mixed_dtype_promotion(sp):
    rng = np.random.default_rng(3)
    shape = (500, 400)
    for da, db in [(np.float32, np.float64), (np.int64, np.float32), (np.int32, np.int64), (np.int32, np.float32)]:
        ai, av = rand_operand(rng, shape, 30000, da)
        bi, bv = rand_operand(rng, shape, 30000, db)
        A, B = sp.FlatSparray(ai, av, shape=shape, is_canonical=True), sp.FlatSparray(bi, bv, shape=shape, is_canonical=True)
        oa, ob = fo.OracleSparray(ai, av, shape=shape, is_canonical=True), fo.OracleSparray(bi, bv, shape=shape, is_canonical=True)
        for f in (OPS["add"], OPS["mul"], OPS["sub"], OPS["lt"], OPS["max"]):
            assert_sparse_same(f(A, B), f(oa, ob))


def test_unaligned_views(sp):
    """Operands that start at odd element offsets (8-byte / 4-byte aligned only):
    exercises the head/tail path around the 16-byte TMA bulk copies."""
    import torch
    rng = np.random.default_rng(9)
    shape = (3000, 3000)
    for dtype in (np.float32, np.float64):
        ai, av = rand_operand(rng, shape, 50000, dtype)
        bi, bv = rand_operand(rng, shape, 50000, dtype)
        for off_a, off_b in [(1, 0), (0, 1), (1, 3), (3, 2)]:
            def shifted(x, off):
                buf = torch.empty(len(x) + off, dtype=torch.from_numpy(x[:1]).dtype, device="cuda")
                buf[off:] = torch.from_numpy(x).cuda()
                return buf[off:]
            A = sp.FlatSparray(shifted(ai, off_a), shifted(av, off_b), shape=shape, is_canonical=True)
            B = sp.FlatSparray(shifted(bi, off_b), shifted(bv, off_a), shape=shape, is_canonical=True)
            oa = fo.OracleSparray(ai, av, shape=shape, is_canonical=True)
            ob = fo.OracleSparray(bi, bv, shape=shape, is_canonical=True)
            assert_sparse_same(A + B, oa + ob)
            assert_sparse_same(A * B, oa * ob)


def test_config1_full_size(sp):
    """BASELINE config 1: two 10000x10000 float64 operands at 1% density, a+b and a*b, bit-exact."""
    shape = (10000, 10000)
    ai, av = fo.random_operand(shape, 0.01, np.float64, 1)
    bi, bv = fo.random_operand(shape, 0.01, np.float64, 2)
    A, B = sp.FlatSparray(ai, av, shape=shape, is_canonical=True), sp.FlatSparray(bi, bv, shape=shape, is_canonical=True)
    oa, ob = fo.OracleSparray(ai, av, shape=shape, is_canonical=True), fo.OracleSparray(bi, bv, shape=shape, is_canonical=True)
    assert_sparse_same(A + B, oa + ob)
    assert_sparse_same(A * B, oa * ob)


def test_host_buffer_seam(sp):
    """The spb_host_* entry points a shim inside the reference's compat.py would bind."""
    import ctypes
    from sparray_b200._lib import lib, check
    rng = np.random.default_rng(1)
    a = np.unique(rng.integers(0, 10 ** 6, 40000)).astype(np.int64)
    b = np.unique(rng.integers(0, 10 ** 6, 50000)).astype(np.int64)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    out = np.empty(len(a) + len(b), np.int64)
    lut = np.empty(len(a) + len(b), np.uint8)
    ao, bo = np.empty(len(a), np.uint8), np.empty(len(b), np.uint8)
    cnt = ctypes.c_int64(0)
    check(lib.spb_host_union1d_sorted(p(a), len(a), p(b), len(b), p(out), p(lut), p(ao), p(bo), ctypes.byref(cnt)))
    w = fo.union1d_sorted(a, b, return_masks=True)
    np.testing.assert_array_equal(out[:cnt.value], w[0])
    np.testing.assert_array_equal(lut[:cnt.value], w[1])
    np.testing.assert_array_equal(ao.astype(bool), w[2])
    np.testing.assert_array_equal(bo.astype(bool), w[3])
    m = min(len(a), len(b))
    out, ia, ib = np.empty(m, np.int64), np.empty(m, np.int64), np.empty(m, np.int64)
    check(lib.spb_host_intersect1d_sorted(p(a), len(a), p(b), len(b), p(out), p(ia), p(ib), ctypes.byref(cnt)))
    w = fo.intersect1d_sorted(a, b, return_inds=True)
    for g, ww in zip((out, ia, ib), w):
        np.testing.assert_array_equal(g[:cnt.value], ww)


def test_wide_key_tiles(sp):
    """Every tile spans far more than 2^32 (keys are multiples of 2^33): the 64-bit key path of the
    merge, including the unions' direct-to-global emission, against the oracle."""
    rng = np.random.default_rng(33)
    n = 30000
    a = (np.unique(rng.integers(0, 2 ** 29, n)) << 33).astype(np.int64)
    b = (np.unique(rng.integers(0, 2 ** 29, n)) << 33).astype(np.int64)
    b[::7] = a[:len(b[::7])] if len(a) >= len(b[::7]) else b[::7]        # force some matches
    b = np.unique(b)
    av = rng.standard_normal(len(a)).astype(np.float32)
    bv = rng.standard_normal(len(b)).astype(np.float32)
    shape = (2 ** 62,)
    A = sp.FlatSparray(a, av, shape=shape, is_canonical=True)
    B = sp.FlatSparray(b, bv, shape=shape, is_canonical=True)
    for op, uf in (("add", np.add), ("maximum", np.maximum)):
        wi, wv = fo.pairwise_union(a, av, b, bv, uf)
        got = A + B if op == "add" else A.maximum(B)
        assert_sparse_same(got, (wi, wv, shape))
    wi, wv = fo.pairwise_intersect(a, av, b, bv, np.multiply)
    assert len(wi) > 100
    assert_sparse_same(A * B, (wi, wv, shape))
    from sparray_b200 import compat as sc
    for got, want in zip(sc.union1d_sorted(a, b, return_masks=True), fo.union1d_sorted(a, b, return_masks=True)):
        np.testing.assert_array_equal(to_np(got).astype(want.dtype), want)
