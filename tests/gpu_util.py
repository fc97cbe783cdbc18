"""Helpers shared by the -m gpu tests (the parity tests proper)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import flat_oracle as fo  # noqa: E402  (the checker; never the thing under test)


def to_np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def assert_sparse_same(got, want, exact=True, rtol=0.0):
    """got: product FlatSparray; want: OracleSparray or (indices, data, shape)."""
    if isinstance(want, tuple):
        wi, wd, ws = want
    else:
        wi, wd, ws = want.indices, want.data, want.shape
    gi, gd = to_np(got.indices), to_np(got.data)
    assert tuple(got.shape) == tuple(ws), (got.shape, ws)
    assert gi.dtype == np.int64
    np.testing.assert_array_equal(gi, wi)
    assert gd.dtype == wd.dtype, (gd.dtype, wd.dtype)
    if exact:
        np.testing.assert_array_equal(gd, wd)
        if wd.dtype.kind == "f":
            np.testing.assert_array_equal(np.signbit(gd), np.signbit(wd))
    else:
        scale = np.max(np.abs(wd)) if wd.size else 1.0
        np.testing.assert_allclose(gd, wd, rtol=rtol, atol=rtol * scale)


def rand_operand(rng, shape, nnz, dtype):
    size = int(np.prod(shape, dtype=object))
    idx = np.unique(rng.integers(0, size, size=nnz, dtype=np.int64))
    if np.dtype(dtype).kind in "iu":
        val = rng.integers(-1000, 1000, size=len(idx)).astype(dtype)
    elif np.dtype(dtype) == np.bool_:
        val = rng.random(len(idx)) < 0.5
    else:
        val = rng.standard_normal(len(idx)).astype(dtype)
    return idx, val
