"""GPU parity: transpose / swapaxes / reshape, axis sums, dot (SpMV), __getitem__,
value maps and the non-canonical ctor -- against the golden fixtures (produced by
the reference) and against the CPU oracle on seeded random inputs.

Tolerances (north_star): indices and moved values bit-exact; float64 reductions
and dot 1e-12 relative, float32 1e-5 relative (relative to the largest output)."""
import numpy as np
import pytest

import golden_runner as gr
from gpu_util import assert_sparse_same, fo, rand_operand, to_np

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5, np.dtype(np.int64): 1e-12,
        np.dtype(np.int32): 1e-12}


@pytest.fixture(scope="module")
def sp():
    import sparray_b200
    return sparray_b200


@pytest.mark.parametrize("case", [c for op in ("transpose", "reduce", "dot_vec", "dot_1d", "getitem", "unary", "ctor")
                                  for c in gr.cases(op)], ids=lambda c: c["name"])
def test_golden(sp, case):
    gr.check_case(case, gr.run_case(case, sp.FlatSparray))


@pytest.mark.parametrize("case", gr.cases("prog"), ids=lambda c: c["name"])
def test_golden_f_rows(sp, case):
    """SURVEY.md section 8(f) rows against outputs of the reference itself (tests/golden/make_golden_f.py)."""
    gr.check_case(case, gr.run_case(case, sp.FlatSparray))


SHAPES = [((300, 500), 40000), ((37, 41, 43), 30000), ((16, 8, 32, 12), 25000), ((64, 64, 64, 16), 300000),
          ((5, 1, 7, 1, 9), 200), ((2 ** 20, 2 ** 18), 100000)]


@pytest.mark.parametrize("shape,nnz", SHAPES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64], ids=lambda d: np.dtype(d).name)
def test_transpose_random(sp, shape, nnz, dtype):
    rng = np.random.default_rng(len(shape) * 1000 + nnz)
    idx, val = rand_operand(rng, shape, nnz, dtype)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    perms = [tuple(range(len(shape) - 1, -1, -1))]
    for _ in range(4):
        perms.append(tuple(rng.permutation(len(shape)).tolist()))
    for axes in perms:
        got = A.transpose(*axes)
        want = fo.transpose(idx, val, shape, axes)
        assert_sparse_same(got, want)
    assert_sparse_same(A.T, fo.transpose(idx, val, shape, None))
    if len(shape) >= 3:
        got = A.swapaxes(0, 2)
        axes = list(range(len(shape)))
        axes[0], axes[2] = 2, 0
        assert_sparse_same(got, fo.transpose(idx, val, shape, axes))


def test_transpose_shape_preserving_permutation(sp):
    """the reference returns self here; numpy semantics are the arbiter."""
    rng = np.random.default_rng(2)
    for shape, axes in [((40, 40), (1, 0)), ((8, 8, 8), (2, 1, 0)), ((16, 16, 16, 4), (2, 1, 0, 3))]:
        idx, val = rand_operand(rng, shape, 500, np.float64)
        A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
        got = A.transpose(*axes)
        np.testing.assert_array_equal(got.toarray(), np.transpose(fo.toarray(idx, val, shape), axes))
        assert_sparse_same(got, fo.transpose(idx, val, shape, axes))


def test_transpose_round_trip_property(sp):
    """Size-independent property: transposing there and back is the identity; values only move."""
    rng = np.random.default_rng(4)
    shape = (128, 96, 80, 24)
    idx, val = rand_operand(rng, shape, 2_000_000, np.float32)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    axes = (2, 0, 3, 1)
    inv = tuple(np.argsort(axes).tolist())
    B = A.transpose(*axes)
    bi = to_np(B.indices)
    assert np.all(np.diff(bi) > 0)
    assert abs(float(to_np(B.data).astype(np.float64).sum()) - float(val.astype(np.float64).sum())) < 1e-3
    assert_sparse_same(B.transpose(*inv), (idx, val, shape))


def test_reshape_ravel_are_metadata_only(sp):
    rng = np.random.default_rng(6)
    idx, val = rand_operand(rng, (30, 40, 50), 5000, np.float64)
    A = sp.FlatSparray(idx, val, shape=(30, 40, 50), is_canonical=True)
    B = A.reshape((1200, 50))
    assert B.indices.data_ptr() == A.indices.data_ptr() and B.data.data_ptr() == A.data.data_ptr()
    assert A.reshape((-1, 100)).shape == (600, 100)
    assert A.ravel().shape == (60000,)
    np.testing.assert_array_equal(B.toarray(), fo.toarray(idx, val, (30, 40, 50)).reshape(1200, 50))
    # config 3's "reshape then .T": metadata reshape followed by a real re-sort
    got = B.T
    assert_sparse_same(got, fo.transpose(idx, val, (1200, 50), None))


@pytest.mark.parametrize("shape,nnz", SHAPES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32], ids=lambda d: np.dtype(d).name)
def test_sum_axis_random(sp, shape, nnz, dtype):
    rng = np.random.default_rng(len(shape) * 77 + nnz)
    idx, val = rand_operand(rng, shape, nnz, dtype)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    for axis in range(len(shape)):
        got = A.sum(axis=axis)
        want = fo.sum_axis(idx, val, shape, axis)
        assert got.dtype == np.float64
        assert_sparse_same(got, want, exact=False, rtol=RTOL[np.dtype(dtype)])
    total = A.sum()
    np.testing.assert_allclose(float(total), float(val.astype(np.float64).sum()), rtol=1e-5 if dtype == np.float32 else 1e-12)
    got = A.mean(axis=0)
    want = fo.mean_axis(idx, val, shape, 0)
    assert_sparse_same(got, want, exact=False, rtol=RTOL[np.dtype(dtype)])


def test_sum_axis_long_segments(sp):
    """Segments far longer than a tile (carry crosses many tiles) and all-equal keys."""
    rng = np.random.default_rng(8)
    shape = (3, 200000)
    idx, val = rand_operand(rng, shape, 500000, np.float64)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    assert_sparse_same(A.sum(axis=1), fo.sum_axis(idx, val, shape, 1), exact=False, rtol=1e-12)
    assert_sparse_same(A.sum(axis=0), fo.sum_axis(idx, val, shape, 0), exact=False, rtol=1e-12)
    shape = (1, 10 ** 6)
    idx, val = rand_operand(rng, shape, 300000, np.float32)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    assert_sparse_same(A.sum(axis=1), fo.sum_axis(idx, val, shape, 1), exact=False, rtol=1e-5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32, np.int16, np.int8],
                         ids=lambda d: np.dtype(d).name)
def test_last_axis_sum_and_spmv_long_rows_edge_cases(sp, dtype):
    """The long-run kernels (a lane owns four consecutive elements; csrc/axis_sum.cu
    axis_accumulate_lastaxis_kernel, csrc/reduce.cu spmv_runs_kernel): rows that start and end inside a
    lane, lanes that span three rows (1-element rows between long ones), a ragged last chunk, operands
    that are views at an odd element offset (no 128-bit loads), indices past prod(shape) after resize()."""
    import torch
    rng = np.random.default_rng(21)
    ncols = 1000
    # row lengths: long rows with single-element and empty rows sprinkled in
    lens = rng.integers(60, 200, size=400)
    lens[rng.random(400) < 0.15] = 1
    lens[rng.random(400) < 0.05] = 0
    lens[5:9] = [1, 1, 2, 1]                                   # a lane spanning several rows
    rows = np.repeat(np.arange(400), lens)
    cols = np.concatenate([np.sort(rng.choice(ncols, size=k, replace=False)) for k in lens]) if len(rows) else rows
    idx = (rows * ncols + cols).astype(np.int64)
    n = len(idx)
    assert n >= 8 * 400
    val = (rng.integers(-50, 50, size=n) if np.dtype(dtype).kind == "i" else rng.standard_normal(n)).astype(dtype)
    rtol = 1e-5 if np.dtype(dtype) == np.float32 else 1e-12
    shape = (400, ncols)
    x = rng.standard_normal(ncols)
    for cut in (0, 1, 2, 3, 5):                                # views: odd offsets, ragged tails
        i_t = torch.from_numpy(idx).cuda()[cut:]
        v_t = torch.from_numpy(val).cuda()[cut:]
        A = sp.FlatSparray(i_t, v_t, shape=shape, is_canonical=True)
        want = fo.sum_axis(idx[cut:], val[cut:], shape, 1)
        assert_sparse_same(A.sum(axis=1), want, exact=False, rtol=rtol)
        if np.dtype(dtype).kind == "f":
            xx = x.astype(dtype)
            got = A.dot(xx)
            ref = fo.dot_dense_vector(idx[cut:], val[cut:], shape, xx)
            np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * np.abs(ref).max())
    # a 3-D array: the two leading axes keep their place, the last one is summed
    shape3 = (20, 20, ncols)
    A = sp.FlatSparray(torch.from_numpy(idx).cuda(), torch.from_numpy(val).cuda(), shape=shape3, is_canonical=True)
    assert_sparse_same(A.sum(axis=2), fo.sum_axis(idx, val, shape3, 2), exact=False, rtol=rtol)
    # stored indices past prod(shape) (resize() allows them): they must not touch anything
    A = sp.FlatSparray(torch.from_numpy(idx).cuda(), torch.from_numpy(val).cuda(), shape=shape, is_canonical=True)
    A.resize((300, ncols))
    keep = idx < 300 * ncols
    got = A.sum(axis=1)
    want = fo.sum_axis(idx[keep], val[keep], (300, ncols), 1)
    assert_sparse_same(got, want, exact=False, rtol=rtol)


SLAB_CASES = [
    # shape, nnz, axes: multi-tile runs of the one-pass slab kernel (csrc/slab_sum.cu)
    ((64, 48, 40, 16), 400_000, (3, 2, 1)),          # last axis, 16-cell suffix, 640-cell suffix
    ((300, 7, 100, 3), 200_000, (3, 2, 1)),          # nothing a power of two
    ((1000, 1000), 600_000, (1,)),                   # rows of a matrix
    ((50, 20000), 300_000, (1,)),                    # slabs several tiles long
    ((2000, 3, 64), 150_000, (2, 1)),
]


@pytest.mark.parametrize("shape,nnz,axes", SLAB_CASES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int8], ids=lambda d: np.dtype(d).name)
def test_sum_axis_slab_kernel(sp, shape, nnz, axes, dtype):
    """The slab kernels against the oracle: skewed fill (whole regions of empty tiles, tiles far fuller than the
    average, so several batches per tile), explicit zeros, sums that cancel to zero."""
    from sparray_b200 import _kernels as k
    from sparray_b200._lib import lib
    rng = np.random.default_rng(nnz + len(shape))
    size = int(np.prod(shape))
    # two thirds of the elements in the first fifth of the index space, none in the second fifth
    a = rng.integers(0, size // 5, size=2 * nnz // 3, dtype=np.int64)
    b = rng.integers(2 * (size // 5), size, size=nnz // 3, dtype=np.int64)
    idx = np.unique(np.concatenate([a, b]))
    if np.dtype(dtype).kind in "iu":
        val = rng.integers(-100, 100, size=len(idx)).astype(dtype)
    else:
        val = rng.standard_normal(len(idx)).astype(dtype)
    val[::7] = 0                                     # explicit zeros are stored elements like any other
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    for axis in axes:
        slab = k.slab_geometry(shape, axis)
        assert lib.spb_axis_sum_slab_ok(len(idx), *slab) == 1, "case no longer exercises the slab kernel"
        want = fo.sum_axis(idx, val, shape, axis)
        new_shape = want[2]
        cells = int(np.prod(new_shape))
        exact = np.dtype(dtype).kind in "iu"
        # called directly: the host's choice between this and the dense-accumulator path is a measured preference
        oi, ov, cnt = k.axis_sum_slab(A.indices, A.data, slab, min(len(idx), cells))
        m = int(cnt.item())
        got = sp.FlatSparray(oi[:m], ov[:m], shape=new_shape, is_canonical=True)
        assert got.dtype == np.float64
        assert_sparse_same(got, want, exact=exact, rtol=RTOL.get(np.dtype(dtype), 1e-12))
        # ... and whatever path sum() itself takes
        assert_sparse_same(A.sum(axis=axis), want, exact=exact, rtol=RTOL.get(np.dtype(dtype), 1e-12))


def test_sum_axis_slab_ignores_indices_past_the_shape(sp):
    """resize() may leave flat indices >= prod(shape) behind (flat.py:185-187 only asserts prod(shape) >= nnz):
    they must not touch anybody's accumulators."""
    from sparray_b200 import _kernels as k
    import torch
    shape = (100, 50)
    idx = np.arange(0, 6000, 3, dtype=np.int64)      # 2000 elements, 334 of them past the 5000 cells
    val = np.ones(len(idx), dtype=np.float64)
    ti, tv = torch.from_numpy(idx).cuda(), torch.from_numpy(val).cuda()
    oi, ov, cnt = k.axis_sum_slab(ti, tv, k.slab_geometry(shape, 1), 100)
    m = int(cnt.item())
    keep = idx < 5000
    want = np.bincount(idx[keep] // 50, minlength=100).astype(np.float64)
    np.testing.assert_array_equal(oi[:m].cpu().numpy(), np.nonzero(want)[0])
    np.testing.assert_array_equal(ov[:m].cpu().numpy(), want[want > 0])


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_sum_axis_zero_sums_survive(sp, axis):
    """Dense path without flags (accumulators start at -0.0 as the "untouched" mark): stored zeros of either sign,
    sums that cancel exactly, NaN and inf must all come out as stored elements, with +0.0 where numpy's
    bincount gives +0.0 -- and an untouched cell must not."""
    from sparray_b200 import _kernels as k
    rng = np.random.default_rng(5 + axis)
    shape = (12, 16, 1024)               # a long innermost axis: the flag-less variant applies to axes 0 and 1
    cells = 12 * 16 * 1024 // shape[axis]
    idx = np.unique(rng.integers(0, 12 * 16 * 1024, size=2 * cells, dtype=np.int64))   # ~2 elements per output cell
    val = rng.standard_normal(len(idx))
    val[::5] = 0.0
    val[1::5] = -0.0
    val[2::50] = np.nan
    val[3::70] = np.inf
    # pairs that cancel exactly: two elements of the same output cell with opposite values
    coords = np.array(np.unravel_index(idx, shape))
    keep = [a for a in range(3) if a != axis]
    out_key = np.ravel_multi_index(coords[keep], [shape[a] for a in keep])
    order = np.argsort(out_key, kind="stable")
    same = np.nonzero(out_key[order][1:] == out_key[order][:-1])[0][::3]
    val[order[same + 1]] = -val[order[same]]
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    assert len(idx) >= cells // 4 and k.dense_sum_cells(shape, axis)
    got = A.sum(axis=axis)
    wi, wd, ws = fo.sum_axis(idx, val, shape, axis)
    gi, gd = to_np(got.indices), to_np(got.data)
    np.testing.assert_array_equal(gi, wi)
    np.testing.assert_array_equal(np.isnan(gd), np.isnan(wd))
    ok = ~np.isnan(wd)
    np.testing.assert_allclose(gd[ok], wd[ok], rtol=1e-12, atol=1e-12)
    zero = ok & (wd == 0) & (gd == 0)            # (a different order of additions may leave a rounding residue)
    assert zero.sum() > min(100, cells // 8)
    np.testing.assert_array_equal(np.signbit(gd[zero]), np.signbit(wd[zero]))
    # twice: the accumulators were handed back clean
    np.testing.assert_array_equal(to_np(A.sum(axis=axis).indices), wi)


@pytest.mark.parametrize("dtype", [np.float16, np.uint16, np.uint32, np.uint64], ids=lambda d: np.dtype(d).name)
def test_conversion_only_dtypes(sp, dtype):
    """float16 and the unsigned integers have no kernels of their own: arithmetic runs on a wider carrier
    (float32 / int32 / int64) and is narrowed back, value-moving kernels see same-sized integers.  Against numpy on
    the dense arrays, bit for bit (one float16 op through float32 rounds like numpy's own float16 loops)."""
    rng = np.random.default_rng(int(np.dtype(dtype).itemsize) * 7 + ord(np.dtype(dtype).kind))
    shape = (23, 17, 9)
    dt_ = np.dtype(dtype)

    def dense():
        d = np.zeros(shape, dtype=dt_)
        m = rng.random(shape) < 0.3
        if dt_.kind == "f":
            d[m] = rng.standard_normal(int(m.sum())).astype(dt_)
        else:
            d[m] = rng.integers(1, 200, size=int(m.sum())).astype(dt_)
        return d
    da, db = dense(), dense()
    A, B = sp.FlatSparray.from_ndarray(da), sp.FlatSparray.from_ndarray(db)
    assert A.dtype == dt_ and np.array_equal(A.toarray(), da)
    for name, got, want in [
            ("add", A + B, da + db), ("sub", A - B, da - db), ("mul", A * B, da * db),
            ("min", A.minimum(B), np.minimum(da, db)), ("max", A.maximum(B), np.maximum(da, db)),
            ("transpose", A.transpose(2, 0, 1), da.transpose(2, 0, 1)), ("T", A.T, da.T),
            ("row", A[3], da[3]), ("box", A[2:20:3, :, 1:7], da[2:20:3, :, 1:7]),
            ("scalar", A * 3, da * dt_.type(3)), ("astype", A.astype(np.float64), da.astype(np.float64)),
            ("ne", A != B, da != db), ("lt", A < B, da < db)]:
        g = got.toarray() if hasattr(got, "toarray") else np.asarray(got)
        assert g.dtype == want.dtype, (name, g.dtype, want.dtype)
        np.testing.assert_array_equal(g, want, err_msg=name)
    for axis in range(3):
        np.testing.assert_allclose(A.sum(axis=axis).toarray(), da.astype(np.float64).sum(axis=axis), rtol=1e-12, err_msg="sum")
    total = A.sum()                                # data.sum(dtype=self.dtype), flat.py:608-610: wraps for small integers
    assert total.dtype == dt_
    np.testing.assert_allclose(float(total), float(da.sum(dtype=dt_)), rtol=2e-3 if dt_.kind == "f" else 0)
    assert A.max() == da[da != 0].max() and A.min() == da[da != 0].min()        # over the stored elements (flat.py:719-725)
    x = rng.standard_normal(9).astype(np.float64)
    np.testing.assert_allclose(to_np(A.reshape((23 * 17, 9)).dot(x)), da.reshape(23 * 17, 9).astype(np.float64).dot(x), rtol=1e-10, atol=1e-10)
    C = A.copy()
    C[1, 2] = dt_.type(5)
    want = da.copy()
    want[1, 2] = 5
    np.testing.assert_array_equal(C.toarray(), want)
    # non-canonical ctor: duplicates are summed
    idx = np.array([5, 3, 5, 9, 3], dtype=np.int64)
    val = np.array([1, 2, 3, 4, 5]).astype(dt_)
    D = sp.FlatSparray(idx, val, shape=(12,))
    np.testing.assert_array_equal(to_np(D.indices), [3, 5, 9])
    np.testing.assert_array_equal(to_np(D.data), np.array([7, 4, 4]).astype(dt_))


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64], ids=lambda d: np.dtype(d).name)
def test_complex_values(sp, dtype):
    """complex FlatSparrays are composed from the real kernels on the (re, im) parts (sparray_b200/_complex.py):
    against numpy on the dense arrays -- the reference's own complex cases are ``dense2d * 1j`` through the
    element-wise ops (test_math.py:18-20).  Bit-exact where the real formulas are numpy's own (+ - neg conj, products
    with a real operand, data movement, comparisons); a few ulp for complex x complex (numpy fuses the multiply-adds)
    and for abs (numpy: hypot)."""
    rng = np.random.default_rng(11)
    shape = (13, 9, 7)
    ct = np.dtype(dtype)
    ft = np.float32 if ct == np.complex64 else np.float64

    def dense(p=0.3):
        m = rng.random(shape) < p
        d = np.zeros(shape, dtype=ct)
        d[m] = (rng.standard_normal(int(m.sum())) + 1j * rng.standard_normal(int(m.sum()))).astype(ct)
        d[rng.random(shape) < 0.05] = 2.0               # purely real entries
        d[rng.random(shape) < 0.05] = -1.5j             # purely imaginary entries
        return d
    da, db = dense(), dense()
    dr = np.where(rng.random(shape) < 0.3, rng.standard_normal(shape), 0).astype(ft)
    A, B, R = sp.FlatSparray.from_ndarray(da), sp.FlatSparray.from_ndarray(db), sp.FlatSparray.from_ndarray(dr)
    assert A.dtype == ct and sp.is_sparray(A) and A.nnz == np.count_nonzero(da)
    np.testing.assert_array_equal(A.toarray(), da)
    idx, val = A.to_numpy()
    np.testing.assert_array_equal(val, da.ravel()[idx])
    for name, got, want in [
            ("add", A + B, da + db), ("sub", A - B, da - db), ("mul", A * B, da * db),
            ("add_real", A + R, da + dr), ("radd_real", R + A, dr + da), ("mul_real", A * R, da * dr), ("rmul_real", R * A, dr * da),
            ("rsub_real", R - A, dr - da), ("neg", -A, -da), ("conj", A.conj(), da.conj()), ("real", A.real, da.real), ("imag", A.imag, da.imag),
            ("scalar_real", A * 2.5, da * ft(2.5)), ("scalar_cplx", A * (1 - 2j), da * ct.type(1 - 2j)), ("real_times_cplx", R * 2j, dr * ct.type(2j)),
            ("transpose", A.transpose(2, 0, 1), da.transpose(2, 0, 1)), ("T", A.T, da.T),
            ("row", A[3], da[3]), ("box", A[2:11:3, :, 1:6], da[2:11:3, :, 1:6]),
            ("reshape", A.reshape((13 * 9, 7)), da.reshape(13 * 9, 7)), ("astype", A.astype(np.complex128), da.astype(np.complex128)),
            ("ne", A != B, da != db), ("pow2", A ** 2, da * da), ("to_complex", R.astype(ct), dr.astype(ct))]:
        g = got.toarray()
        assert g.dtype == want.dtype, (name, g.dtype, want.dtype)
        if name in ("mul", "scalar_cplx", "pow2"):
            # complex x complex: numpy's SIMD loops fuse the multiply-adds (fmaddsub), which no composition of
            # separately rounded products reproduces bit for bit -- and which depends on the CPU numpy runs on
            np.testing.assert_allclose(g, want, rtol=4 * np.finfo(ft).eps, atol=4 * np.finfo(ft).eps, err_msg=name)
        else:
            np.testing.assert_array_equal(g, want, err_msg=name)
    np.testing.assert_allclose(abs(A).toarray(), np.abs(da), rtol=1e-6 if ct == np.complex64 else 1e-14)
    np.testing.assert_allclose((A / 4).toarray(), da / 4, rtol=1e-6 if ct == np.complex64 else 1e-15)
    np.testing.assert_allclose(A.sum(), da.sum(), rtol=1e-5 if ct == np.complex64 else 1e-12)
    assert A[3, 2, 1] == da[3, 2, 1] and A[0, 0, 0] == da[0, 0, 0]
    # (A == B) on the union of the stored elements
    eq = (A == A.copy())
    assert eq.dtype == np.bool_ and bool(to_np(eq.data).all())
    C = A.copy()
    C[1, 2, 3] = 3 - 4j
    C[:, 1, 0] = 1j
    want = da.copy()
    want[1, 2, 3] = 3 - 4j
    want[:, 1, 0] = 1j
    np.testing.assert_array_equal(C.toarray(), want)
    # non-canonical ctor: duplicates are summed
    D = sp.FlatSparray(np.array([5, 3, 5, 9], dtype=np.int64), np.array([1 + 1j, 2, 3 - 2j, 4j], dtype=ct), shape=(12,))
    np.testing.assert_array_equal(to_np(D.indices), [3, 5, 9])
    np.testing.assert_array_equal(to_np(D.data), np.array([2, 4 - 1j, 4j], dtype=ct))
    # what cannot be composed from the real kernels is refused loudly, as is what the reference refuses
    for bad in (lambda: A.sum(axis=0), lambda: A < B, lambda: A.sin(), lambda: A.dot(np.ones(7)), lambda: A + 1, lambda: A.max()):
        with pytest.raises(TypeError):
            bad()


def test_config2_chain(sp):
    """BASELINE config 2 at full size: c = a*b (intersection) then c.sum(axis=2); also a.sum(axis=2)."""
    shape = (256, 256, 256, 64)
    ai, av = fo.random_operand(shape, 0.01, np.float32, 1)
    bi, bv = fo.random_operand(shape, 0.01, np.float32, 2)
    A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
    B = sp.FlatSparray(bi, bv, shape=shape, is_canonical=True)
    C = A * B
    ci, cv = fo.pairwise_intersect(ai, av, bi, bv, np.multiply)
    assert_sparse_same(C, (ci, cv, shape))
    assert_sparse_same(C.sum(axis=2), fo.sum_axis(ci, cv, shape, 2), exact=False, rtol=1e-5)
    assert_sparse_same(A.sum(axis=2), fo.sum_axis(ai, av, shape, 2), exact=False, rtol=1e-5)
    assert_sparse_same(A.sum(axis=3), fo.sum_axis(ai, av, shape, 3), exact=False, rtol=1e-5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=lambda d: np.dtype(d).name)
def test_spmv_random(sp, dtype):
    rng = np.random.default_rng(12)
    rtol = RTOL[np.dtype(dtype)]
    for shape, nnz in [((2000, 3000), 200000), ((50, 40, 600), 100000), ((1, 5000), 2000), ((10 ** 5, 10 ** 5), 10 ** 6)]:
        idx, val = rand_operand(rng, shape, nnz, dtype)
        x = rng.standard_normal(shape[-1]).astype(dtype)
        A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
        got = A.dot(x)
        want = fo.dot_dense_vector(idx, val, shape, x)
        assert isinstance(got, np.ndarray) and got.dtype == want.dtype and got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * np.abs(want).max())
        # sparse-vector operand -> sparse result (flat.py:545-561)
        xs = x.copy()
        xs[rng.random(len(xs)) < 0.5] = 0
        got = A.dot(sp.FlatSparray.from_ndarray(xs))
        want = fo.dot_dense_vector(idx, val, shape, xs)
        np.testing.assert_allclose(got.toarray(), want, rtol=rtol, atol=rtol * np.abs(want).max())
    with pytest.raises(ValueError):
        A.dot(np.ones(7, dtype=dtype))


def test_getitem_random_vs_dense(sp):
    """The reference's own method: differential test against dense numpy indexing."""
    rng = np.random.default_rng(14)
    shape = (23, 17, 29)
    idx, val = rand_operand(rng, shape, 4000, np.float64)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    D = fo.toarray(idx, val, shape)
    O = fo.OracleSparray(idx, val, shape=shape, is_canonical=True)
    queries = [(3,), (-1,), (5, 6), (slice(None), 4), (Ellipsis, 7), (2, Ellipsis, 3), (slice(2, 20, 3),),
               (slice(1, None), slice(None, 9), slice(4, 25, 5)), (slice(5, 6), 3, slice(None)),
               ([1, 5, 20], [0, 16, 3], [28, 0, 7]), ([20, 5, 1], [0, 16, 3], [28, 0, 7]),
               ([3, 3, 9],), ([1, 3], slice(None), slice(2, 9)), ([4, 0, 7], 2, slice(None)),
               (np.arange(23) % 3 == 0,), (slice(None), [16, 2], 5)]
    for q in queries:
        got = A[q]
        want_dense = D[q]
        np.testing.assert_array_equal(got.toarray(), want_dense, err_msg=str(q))
        gi = to_np(got.indices)
        assert np.all(np.diff(gi) > 0)
        cells = np.arange(D.size).reshape(shape)[q].ravel()
        if len(np.unique(cells)) == len(cells):     # the reference (and so the oracle) is undefined on repeats
            assert_sparse_same(got, O[q])
    for q in [(3, 4, 5), (-1, -1, -1), (0, 0, 0), (22, 16, 28)]:
        assert A[q] == D[q]
    assert A[...] is A and A[:, :, :] is A
    with pytest.raises(IndexError):
        A[23]
    with pytest.raises(IndexError):
        A[1, 2, 3, 4]


def test_getitem_rows_of_large_matrix(sp):
    """asv workload shapes (bench/benchmarks/ops.py:19-38): arr[876], arr[:,273], arr[:5,:5], arr[154,145]."""
    rng = np.random.default_rng(16)
    shape = (3000, 4000)
    idx, val = rand_operand(rng, shape, 1_200_000, np.float64)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    O = fo.OracleSparray(idx, val, shape=shape, is_canonical=True)
    for q in [(876,), (slice(None), 273), (slice(None, 5), slice(None, 5)), (slice(100, 2900, 7), slice(3, 3999, 11))]:
        assert_sparse_same(A[q], O[q])
    assert A[154, 145] == O[154, 145]
    assert_sparse_same(A.diagonal(), O[np.arange(3000), np.arange(3000)])
    assert_sparse_same(A.diagonal(5), O[np.arange(3000), np.arange(3000) + 5])
    assert_sparse_same(A.diagonal(-7), O[np.arange(2993) + 7, np.arange(2993)])


def test_ctor_canonicalises(sp):
    rng = np.random.default_rng(18)
    for dtype in (np.float64, np.float32, np.int64):
        idx = rng.integers(0, 5000, size=40000).astype(np.int64)
        val = rand_operand(rng, (10 ** 6,), 80000, dtype)[1][:40000]
        A = sp.FlatSparray(idx, val, shape=(50, 100))
        wi, wv = fo.canonicalize(idx, val)
        rt = 1e-5 if dtype == np.float32 else 1e-12
        assert_sparse_same(A, (wi, wv, (50, 100)), exact=(dtype == np.int64), rtol=rt)
    A = sp.FlatSparray([4, 1, 3, 0], [2, -1, 1, -2])      # shape inferred (flat.py:36-37)
    assert A.shape == (5,)
    np.testing.assert_array_equal(A.toarray(), [-2, -1, 0, 1, 2])


def test_converters_and_unary(sp):
    rng = np.random.default_rng(20)
    D = rng.standard_normal((40, 50, 6))
    D[rng.random(D.shape) < 0.8] = 0
    A = sp.FlatSparray.from_ndarray(D)
    wi, wv, _ = fo.from_ndarray(D)
    assert_sparse_same(A, (wi, wv, D.shape))
    np.testing.assert_array_equal(A.toarray(), D)
    np.testing.assert_array_equal((-A).toarray(), -D)
    np.testing.assert_array_equal(abs(A).toarray(), abs(D))
    np.testing.assert_array_equal((A * 2.5).toarray(), D * 2.5)
    np.testing.assert_array_equal((3 * A).toarray(), 3 * D)
    np.testing.assert_array_equal((A / 4).toarray(), D * (1. / 4))
    np.testing.assert_array_equal((A ** 2).toarray(), D ** 2)
    np.testing.assert_array_equal((A > 0.5).toarray(), D > 0.5)
    np.testing.assert_array_equal(A.astype(np.float32).toarray(), D.astype(np.float32))
    np.testing.assert_allclose(A.sin().toarray(), np.sin(D), rtol=1e-15, atol=1e-16)
    assert A.min() == wv.min() and A.max() == wv.max()
    assert A.nnz == len(wi) and A.dtype == np.float64 and len(A) == 40
    nz = A.nonzero()
    for g, w in zip(nz, np.nonzero(D)):
        np.testing.assert_array_equal(g, w)
    import scipy.sparse as ss
    M = ss.random(60, 70, density=0.2, format="csr", random_state=3)
    S = sp.FlatSparray.from_spmatrix(M)
    np.testing.assert_array_equal(S.toarray(), M.toarray())
    np.testing.assert_array_equal(S.tocoo().toarray(), M.toarray())
    B = A.copy()
    B *= 2
    np.testing.assert_array_equal(B.toarray(), D * 2)
    assert sp.is_sparray(A) and not sp.is_sparray(D)
    with pytest.raises(NotImplementedError):
        A += A


def test_assignment_and_setdiag(sp):
    """__setitem__ / setdiag (flat.py:327-364,139-160,395-407); the reference's
    own assertions (sparray/tests/test_indexing.py:99-145) against dense numpy / scipy."""
    import scipy.sparse as ss
    dense1d = np.arange(5) - 2
    dense2d = np.array([[0, 0, 0], [4, 5, 7], [6, 2, 0], [1, 3, 8]], dtype=float) / 2.
    sp1d = sp.FlatSparray([0, 1, 3, 4], [-2, -1, 1, 2], shape=dense1d.shape)
    sp2d = sp.FlatSparray([1, 3, 4, 5, 6, 7, 9, 10, 11], [0, 2, 2.5, 3.5, 3, 1, 0.5, 1.5, 4], shape=dense2d.shape)
    for index in [3, 2, slice(None, 3)]:                       # in structure, not in structure, sub-array
        a, d = sp1d.copy(), dense1d.copy()
        a[index] = 99
        d[index] = 99
        np.testing.assert_array_equal(d, a.toarray())
    for index in [(3, 1), (0, 2), (slice(None, 2), slice(None, 2)), (slice(1, None), 2)]:
        a, d = sp2d.copy(), dense2d.copy()
        a[index] = 99
        d[index] = 99
        np.testing.assert_array_equal(d, a.toarray())
        assert np.all(np.diff(a.indices.cpu().numpy()) > 0)
    a, d = sp2d.copy(), dense2d.copy()
    a[1:3, 0:2] = np.array([1., 2., 3., 4.])                   # one value per cell of the box
    d[1:3, 0:2] = np.array([[1., 2.], [3., 4.]])
    np.testing.assert_array_equal(d, a.toarray())
    with pytest.raises(ValueError):
        sp2d.copy()[:] = 1
    with pytest.raises(NotImplementedError):
        sp2d.copy()[[0, 1], 1] = 1
    sparse2d = ss.csr_matrix(dense2d)
    for k in [0, 1, -1, 2, -2]:
        a, m = sp2d.copy(), sparse2d.copy().tolil()
        a.setdiag(99, offset=k)
        m.setdiag(99, k=k)
        np.testing.assert_array_equal(m.toarray(), a.toarray(), err_msg="k=%d" % k)
    # larger, random: assign a strided box into a 3-D array
    rng = np.random.default_rng(22)
    shape = (40, 30, 20)
    idx, val = rand_operand(rng, shape, 5000, np.float32)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    D = fo.toarray(idx, val, shape)
    A[3:35:2, 5, 1:19:3] = -7.5
    D[3:35:2, 5, 1:19:3] = -7.5
    np.testing.assert_array_equal(D, A.toarray())


def test_dense_operands_and_sparse_broadcasting(sp):
    """broadcasting between sparse operands without densifying and the
    dense (op) sparse helpers (flat.py:436-472).  Mirrors the reference's assertions in
    sparray/tests/test_math.py:26-47,57-64,66-94,224-228 against dense numpy."""
    import scipy.sparse as ss
    from numpy.testing import assert_array_equal
    rng = np.random.default_rng(31)
    dense2d = np.array([[0, 0, 0], [4, 5, 7], [6, 2, 0], [1, 3, 8]], dtype=float) / 2.
    sp2d = sp.FlatSparray([1, 3, 4, 5, 6, 7, 9, 10, 11], [0, 2, 2.5, 3.5, 3, 1, 0.5, 1.5, 4], shape=dense2d.shape)
    # dense ndarray operands, same shape and broadcast
    for bshape in [dense2d.shape, (dense2d.shape[0], 1), (1, dense2d.shape[1])]:
        b = rng.random(bshape)
        assert_array_equal(dense2d + b, sp2d + b)
        assert_array_equal(b + dense2d, b + sp2d)
        assert_array_equal(dense2d - b, sp2d - b)
        assert_array_equal(b - dense2d, b - sp2d)
        assert_array_equal(dense2d * b, (sp2d * b).toarray())
        assert_array_equal(b * dense2d, (b * sp2d).toarray())
        assert_array_equal(dense2d < b, sp2d < b)
        assert_array_equal(np.minimum(dense2d, b), sp2d.minimum(b))
        assert_array_equal(np.maximum(dense2d, b), sp2d.maximum(b))
        np.testing.assert_allclose((sp2d / (b + 1)).toarray(), dense2d / (b + 1), rtol=1e-15)
        assert_array_equal((sp2d // (b + 1)).toarray(), dense2d // (b + 1))
    # sparse operands that broadcast: (4,1) and (1,3) against (4,3)
    for shape in [(dense2d.shape[0], 1), (1, dense2d.shape[1])]:
        s = ss.random(*shape, density=0.6, random_state=5)
        b = sp.FlatSparray.from_spmatrix(s)
        sd = s.toarray()
        r = sp2d + b
        assert isinstance(r, sp.FlatSparray) and r.shape == dense2d.shape
        assert_array_equal(dense2d + sd, r.toarray())
        assert_array_equal(sd + dense2d, (b + sp2d).toarray())
        assert_array_equal(dense2d * sd, (sp2d * b).toarray())
        assert_array_equal(dense2d * sd, (b * sp2d).toarray())
        assert_array_equal(np.maximum(dense2d, sd), sp2d.maximum(b).toarray())
        assert np.all(np.diff(r.indices.cpu().numpy()) > 0)
    # a bigger N-D broadcast: (6,1,8,1) with (6,5,8,7)
    big = rng.standard_normal((6, 5, 8, 7)); big[rng.random(big.shape) < 0.7] = 0
    small = rng.standard_normal((6, 1, 8, 1)); small[rng.random(small.shape) < 0.5] = 0
    A, S = sp.FlatSparray.from_ndarray(big), sp.FlatSparray.from_ndarray(small)
    assert_array_equal(big + small, (A + S).toarray())
    assert_array_equal(big * small, (A * S).toarray())
    assert_array_equal(small - big, (S - A).toarray())
    low = sp.FlatSparray.from_ndarray(small[0, 0])             # fewer dims: (8,1) against 4-D
    assert_array_equal(big + small[0, 0], (A + low).toarray())
    wi, wv, _ = fo.from_ndarray(np.broadcast_to(small, big.shape))
    assert_sparse_same(S._broadcast(big.shape), (wi, wv, big.shape))


def test_dot_matrix_operands(sp):
    """a.dot(B) for a dense or sparse 2-D B: one SpMV per column (flat.py:538-568; the reference's
    assertions test_math.py:180-213, _matmul.py:19-28 against numpy)."""
    rng = np.random.default_rng(33)
    D = rng.standard_normal((30, 20)); D[rng.random(D.shape) < 0.7] = 0
    Bm = rng.standard_normal((20, 7))
    A = sp.FlatSparray.from_ndarray(D)
    np.testing.assert_allclose(A.dot(Bm), D.dot(Bm), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(A @ Bm, D @ Bm, rtol=1e-12, atol=1e-12)
    Bs = Bm.copy(); Bs[rng.random(Bs.shape) < 0.6] = 0
    r = A.dot(sp.FlatSparray.from_ndarray(Bs))
    assert isinstance(r, sp.FlatSparray) and r.shape == (30, 7)
    np.testing.assert_allclose(r.toarray(), D.dot(Bs), rtol=1e-12, atol=1e-12)
    T = rng.standard_normal((4, 5, 20)); T[rng.random(T.shape) < 0.5] = 0
    np.testing.assert_allclose(sp.FlatSparray.from_ndarray(T).dot(Bm), T.dot(Bm), rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        A.dot(np.ones((19, 3)))


def _synth_device(torch, n, size, seed):
    """Sorted unique flat indices in [0, size) with random gaps, and small-integer float64 values
    (every partial sum is then exact in float64, so the checks below can be bit-exact)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    gap = max(size // n, 1)
    steps = torch.randint(1, 2 * gap, (n,), device="cuda", generator=g, dtype=torch.int64)
    idx = torch.cumsum(steps, 0) - 1
    idx = idx[idx < size].contiguous()
    val = torch.randint(-8, 9, (idx.numel(),), device="cuda", generator=g, dtype=torch.int64).to(torch.float64)
    return idx, val


def test_config4_full_size_properties(sp):
    """BASELINE config 4 at full size (10^6 x 10^6 float64, ~10^8 stored elements): the CPU oracle
    would need minutes, so parity is checked through size-independent properties --
    transpose round trip and sortedness, checksums of sums along both axes, SpMV against the row
    sums and its linearity, a+a / a*a against element-wise maps of the data."""
    import torch
    n_side = 10 ** 6
    idx, val = _synth_device(torch, 100_000_000, n_side * n_side, 7)
    W = sp.FlatSparray(idx, val, shape=(n_side, n_side), is_canonical=True)
    n = W.nnz
    assert n > 90_000_000
    total = float(val.sum().item())

    # transpose: sorted unique keys, same multiset of values, involution
    T = W.T
    assert T.shape == (n_side, n_side) and T.nnz == n
    assert bool((T.indices[1:] > T.indices[:-1]).all().item())
    assert float(T.data.sum().item()) == total
    # element (r, c) of W sits at (c, r) of T: check through a sample of positions
    pick = torch.arange(0, n, 9973, device="cuda")
    r, c = idx[pick] // n_side, idx[pick] % n_side
    pos = torch.searchsorted(T.indices, c * n_side + r)
    assert bool((T.indices[pos] == c * n_side + r).all().item())
    assert bool((T.data[pos] == val[pick]).all().item())
    TT = T.T
    assert bool((TT.indices == idx).all().item()) and bool((TT.data == val).all().item())
    del TT

    # axis sums: both keep the grand total; sum(axis=1) equals the row sums of the transpose's sum(axis=0)
    s1 = W.sum(axis=1)
    s0 = W.sum(axis=0)
    assert float(s1.data.sum().item()) == total and float(s0.data.sum().item()) == total
    t0 = T.sum(axis=0)
    assert bool((t0.indices == s1.indices).all().item()) and bool((t0.data == s1.data).all().item())
    del T, t0

    # SpMV: W @ ones are the row sums; linearity with exactly representable operands
    ones = np.ones(n_side)
    y1 = W.dot(ones)
    dense_rows = np.zeros(n_side)
    dense_rows[to_np(s1.indices)] = to_np(s1.data)
    np.testing.assert_array_equal(y1, dense_rows)
    rng = np.random.default_rng(3)
    x = rng.integers(-4, 5, n_side).astype(np.float64)
    z = rng.integers(-4, 5, n_side).astype(np.float64)
    np.testing.assert_array_equal(W.dot(x + z), W.dot(x) + W.dot(z))

    # fused binops against element-wise maps: a + a == 2a on the same pattern, a * a == a^2
    D = W + W
    assert D.nnz == n and bool((D.indices == idx).all().item()) and bool((D.data == 2 * val).all().item())
    del D
    Q = W * W
    assert Q.nnz == n and bool((Q.indices == idx).all().item()) and bool((Q.data == val * val).all().item())


def test_config4_downscale_vs_oracle(sp):
    """BASELINE config[3] (2-D float64, dot with a dense vector and sum(axis=1)) at a tenth of its stated size --
    10^7 stored elements, 2-D and row-sparse like the original (316228 x 316228) -- with NON-integer data,
    against the oracle's restatement of flat.py:538-568 / :607-626 (the full size is property-checked only:
    test_config4_full_size_properties).  Tolerance: 1e-12 relative (north_star, float64)."""
    rng = np.random.default_rng(44)
    shape = (316_228, 316_228)
    nnz = 10_000_000
    idx = np.unique(rng.integers(0, shape[0] * shape[1], size=nnz, dtype=np.int64))
    val = rng.standard_normal(len(idx))
    x = rng.standard_normal(shape[1])
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    want = fo.dot_dense_vector(idx, val, shape, x)
    got = to_np(A.dot(x))
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    assert_sparse_same(A.sum(axis=1), fo.sum_axis(idx, val, shape, 1), exact=False, rtol=1e-12)
    assert_sparse_same(A.sum(axis=0), fo.sum_axis(idx, val, shape, 0), exact=False, rtol=1e-12)


def test_axis_accumulate_and_emit_ranges(sp):
    """The two halves of the dense axis sum on caller-owned buffers (what the sharded sum all-reduces
    in between): accumulating two halves of an operand separately and emitting the cells in two
    64-aligned ranges gives the oracle's sum(axis)."""
    import torch
    from sparray_b200 import _kernels as k, _plan
    rng = np.random.default_rng(12)
    shape = (37, 29, 23)
    idx, val = rand_operand(rng, shape, 6000, np.float32)
    for axis in (0, 1):
        _, in_s, out_s, _ = _plan.axis_sum_plan(shape, axis)
        cells = int(np.prod(shape)) // shape[axis]
        half = len(idx) // 2
        ti, tv = torch.from_numpy(idx).cuda(), torch.from_numpy(val).cuda()
        acc1, t1 = k.axis_accumulate(ti[:half].contiguous(), tv[:half].contiguous(), in_s, out_s, cells)
        acc2, t2 = k.axis_accumulate(ti[half:].contiguous(), tv[half:].contiguous(), in_s, out_s, cells)
        acc = acc1 + acc2
        touched = torch.maximum(t1, t2)
        cut = (cells // 2) // 64 * 64
        k1, s1 = k.emit_touched(acc, touched, 0, cut)
        k2, s2 = k.emit_touched(acc, touched, cut, cells)
        wi, wv, _ = fo.sum_axis(idx, val, shape, axis)
        got_i = np.concatenate([to_np(k1), to_np(k2)])
        got_v = np.concatenate([to_np(s1), to_np(s2)])
        np.testing.assert_array_equal(got_i, wi)
        np.testing.assert_allclose(got_v, wv, rtol=1e-5, atol=1e-6 * np.abs(wv).max())
        assert float(acc.abs().sum().item()) == 0.0 and int(touched.sum().item()) == 0    # emitted cells were zeroed


@pytest.mark.parametrize("seed", range(6))
def test_sum_and_dot_small_shapes_fuzz(sp, seed):
    """Tiny shapes, dense-ish arrays: inside one warp the same output cell is hit by many elements that
    are NOT adjacent (the warp-aggregated atomics must treat them as separate runs), rows shorter and
    longer than a warp, every axis, several dtypes -- against the oracle."""
    rng = np.random.default_rng(100 + seed)
    ndim = int(rng.integers(2, 5))
    shape = tuple(int(x) for x in rng.integers(2, 9, ndim))
    size = int(np.prod(shape))
    nnz = int(rng.integers(1, size + 1))
    dtype = [np.float64, np.float32, np.int32, np.int64][seed % 4]
    idx = np.sort(rng.choice(size, nnz, replace=False)).astype(np.int64)
    val = rng.integers(-5, 6, nnz).astype(dtype) if np.dtype(dtype).kind == "i" else rng.standard_normal(nnz).astype(dtype)
    A = sp.FlatSparray(idx, val, shape=shape, is_canonical=True)
    for axis in range(ndim):
        wi, wv, wshape = fo.sum_axis(idx, val, shape, axis)
        got = A.sum(axis=axis)
        assert got.shape == tuple(wshape)
        assert_sparse_same(got, (wi, wv, tuple(wshape)), exact=False, rtol=1e-5 if dtype == np.float32 else 1e-12)
    if np.dtype(dtype).kind == "f":
        x = rng.standard_normal(shape[-1]).astype(dtype)
        want = fo.dot_dense_vector(idx, val, shape, x)
        got = A.dot(x)
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), want, rtol=1e-5 if dtype == np.float32 else 1e-12,
                                   atol=(1e-5 if dtype == np.float32 else 1e-12) * max(np.abs(want).max(), 1.0))


def test_lazily_counted_chain_every_axis(sp):
    """(a * b) and (a + b) are lazily counted (their element counts stay on the device); sum(axis)
    consumes the pending count without resolving it.  Every axis, and the results themselves."""
    rng = np.random.default_rng(21)
    shape = (48, 40, 36, 16)
    ai, av = rand_operand(rng, shape, 150000, np.float32)
    bi, bv = rand_operand(rng, shape, 140000, np.float32)
    A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
    B = sp.FlatSparray(bi, bv, shape=shape, is_canonical=True)
    ci, cv = fo.pairwise_intersect(ai, av, bi, bv, np.multiply)
    ui, uv = fo.pairwise_union(ai, av, bi, bv, np.add)
    for axis in range(4):
        C = A * B
        assert C._pending is not None                       # not resolved by the product itself
        S = C.sum(axis=axis)
        assert C._pending is not None, "sum(axis) must not force the count read-back"
        assert_sparse_same(S, fo.sum_axis(ci, cv, shape, axis), exact=False, rtol=1e-5)
        U = A + B
        assert_sparse_same(U.sum(axis=axis), fo.sum_axis(ui, uv, shape, axis), exact=False, rtol=1e-5)
    C = A * B
    assert_sparse_same(C, (ci, cv, shape))
    assert C._pending is None and C.nnz == len(ci)


def test_cuda_graph_capture_of_a_chain(sp):
    """The whole chain c = a*b; c.sum(axis) -- and a union, whose look-back words carry launch epochs -- captured
    into a CUDA graph and replayed: every replay must give the eager result, also after the operands'
    values changed in place and after other (eager) calls ran in between."""
    import torch
    rng = np.random.default_rng(31)
    shape = (64, 64, 48, 16)
    ai, av = rand_operand(rng, shape, 400_000, np.float32)
    bi, bv = rand_operand(rng, shape, 380_000, np.float32)
    A = sp.FlatSparray(ai, av, shape=shape, is_canonical=True)
    B = sp.FlatSparray(bi, bv, shape=shape, is_canonical=True)

    def chain():
        c = A * B
        return c, c.sum(axis=2), A + B

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):                       # eager warm-up on the capture stream: scratch gets its size
            chain()
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        c, s, u = chain()
    for trial in range(3):
        if trial == 1:                           # new values, same pattern: replays must pick them up
            av = rng.standard_normal(len(ai)).astype(np.float32)
            A.data.copy_(torch.from_numpy(av).cuda())
        if trial == 2:                           # eager calls in between advance the epoch counter
            with torch.cuda.stream(side):
                for _ in range(5):
                    (A + B).nnz
            side.synchronize()
        g.replay()
        torch.cuda.synchronize()
        oa = fo.OracleSparray(ai, av, shape=shape, is_canonical=True)
        ob = fo.OracleSparray(bi, bv, shape=shape, is_canonical=True)
        oc = oa * ob
        # pending results: read the device-side counts the replay produced
        for got, want, exact in ((c, oc, True), (s, oc.sum(axis=2), False), (u, oa + ob, True)):
            idx, val, cnt = got._pending
            n = int(cnt.item())
            np.testing.assert_array_equal(idx[:n].cpu().numpy(), want.indices)
            if exact:
                np.testing.assert_array_equal(val[:n].cpu().numpy(), want.data)
            else:
                np.testing.assert_allclose(val[:n].cpu().numpy(), want.data, rtol=1e-5, atol=1e-5)


def test_spgemm_vs_scipy(sp):
    """sparse . sparse (flat.py:545-561: the reference goes reshape -> COO -> scipy CSR matmul -> from_spmatrix):
    the expand / sort / compress kernels against scipy's product, 2-D and N-D operands, several dtypes,
    empty rows, an all-empty operand, explicit zeros, a scipy operand."""
    import scipy.sparse as ss
    rng = np.random.default_rng(41)
    for (m, k, n, da, db, dtype) in [(300, 200, 150, 0.05, 0.04, np.float64), (1000, 800, 900, 0.01, 0.01, np.float32),
                                     (64, 4096, 32, 0.02, 0.3, np.float64), (50, 40, 30, 0.2, 0.2, np.int64),
                                     (50, 40, 30, 0.2, 0.2, np.int32)]:
        Am = ss.random(m, k, density=da, random_state=int(rng.integers(1 << 30)), format="csr", dtype=np.float64)
        Bm = ss.random(k, n, density=db, random_state=int(rng.integers(1 << 30)), format="csr", dtype=np.float64)
        if np.dtype(dtype).kind == "i":
            Am.data = np.rint(Am.data * 20) + 1
            Bm.data = np.rint(Bm.data * 20) + 1
        Am, Bm = Am.astype(dtype), Bm.astype(dtype)
        A, B = sp.FlatSparray.from_spmatrix(Am), sp.FlatSparray.from_spmatrix(Bm)
        C = A.dot(B)
        want = (Am @ Bm).tocoo()
        want.sum_duplicates()
        assert isinstance(C, sp.FlatSparray) and C.shape == (m, n) and C.dtype == np.dtype(dtype)
        wi = want.row.astype(np.int64) * n + want.col
        order = np.argsort(wi)
        np.testing.assert_array_equal(to_np(C.indices), wi[order])
        if np.dtype(dtype).kind == "i":
            np.testing.assert_array_equal(to_np(C.data), want.data[order])
        else:
            tol = 1e-5 if dtype == np.float32 else 1e-12
            np.testing.assert_allclose(to_np(C.data), want.data[order], rtol=tol, atol=tol)
        np.testing.assert_array_equal(to_np((A @ B).indices), wi[order])
        # a scipy operand on the right (flat.py:546)
        np.testing.assert_array_equal(to_np(A.dot(Bm).indices), wi[order])
    # N-D operands: contraction of self's last axis with other's second-to-last (flat.py:539-556)
    T = rng.standard_normal((4, 5, 20)); T[rng.random(T.shape) < 0.6] = 0
    U = rng.standard_normal((3, 20, 6)); U[rng.random(U.shape) < 0.5] = 0
    r = sp.FlatSparray.from_ndarray(T).dot(sp.FlatSparray.from_ndarray(U))
    assert r.shape == (4, 5, 3, 6)
    np.testing.assert_allclose(r.toarray(), T.dot(U), rtol=1e-12, atol=1e-12)
    # an operand with nothing stored
    Z = sp.FlatSparray([], np.zeros(0), shape=(20, 6), is_canonical=True)
    r = sp.FlatSparray.from_ndarray(T).dot(Z)
    assert r.shape == (4, 5, 6) and r.nnz == 0


def test_densifying_branches_return_dense_results(sp):
    """The branches whose result is dense by nature (flat.py:478-481,494-496,520-527,570-577,592-593,604-605)
    return the ndarray the reference returns, with the reference's SparseEfficiencyWarning."""
    import warnings
    rng = np.random.default_rng(77)
    D = rng.standard_normal((7, 9)); D[rng.random(D.shape) < 0.6] = 0
    E = rng.standard_normal((7, 9)); E[rng.random(E.shape) < 0.6] = 0
    A, B = sp.FlatSparray.from_ndarray(D), sp.FlatSparray.from_ndarray(E)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        np.testing.assert_array_equal(A + 2.5, D + 2.5)
        np.testing.assert_array_equal(A - 1, D - 1)
        np.testing.assert_array_equal(A < 1, D < 1)
        np.testing.assert_array_equal(A >= 0, D >= 0)
        np.testing.assert_array_equal(A == 0, D == 0)
        np.testing.assert_array_equal(A != 3, D != 3)
        np.testing.assert_array_equal(A ** 0, D ** 0)
        with np.errstate(all="ignore"):
            np.testing.assert_allclose(A ** -2, D ** -2.0, rtol=1e-14)       # CUDA pow: within an ulp of numpy's
        assert sum("densifies" in str(x.message) for x in w) == 8, [str(x.message) for x in w]
    np.testing.assert_array_equal(A.minimum(-0.5), np.minimum(D, -0.5))
    np.testing.assert_array_equal(A.maximum(0.25), np.maximum(D, 0.25))
    with np.errstate(all="ignore"):
        np.testing.assert_array_equal(A / B, D / E)                      # nan / inf where the divisor is empty
        np.testing.assert_array_equal(A // B, D // E)
        np.testing.assert_array_equal(2.0 / A, 2.0 / D)
        np.testing.assert_array_equal(E / A, E / D)
    I = sp.FlatSparray.from_ndarray(np.arange(12).reshape(3, 4) % 3)
    np.testing.assert_array_equal(I + 1, np.arange(12).reshape(3, 4) % 3 + 1)
    with pytest.raises(ValueError):
        I ** -1
    # reflected matmul / _rdot (base.py:79-82)
    M = rng.standard_normal((5, 7))
    np.testing.assert_allclose(A._rdot(M), M.dot(D), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(M @ A, M @ D, rtol=1e-12, atol=1e-12)
