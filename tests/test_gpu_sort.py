"""Parity tests of the re-key + radix sort (csrc/sort.cu) through the C ABI, against numpy's stable
argsort: every packed-key mode (32-bit packed keys, 32-bit keys with the sorted digit squeezed out for
33..40-bit keys, 64-bit packed keys), every value width, partial tiles, unaligned value slices (the
TMA staging falls back to plain loads), many digit boundaries inside one tile."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def k():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from sparray_b200 import _kernels
    return _kernels


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _check(k, keys, vals, div, bits):
    """Stable sort by keys // div on bits `bits`; bit-exact keys and values."""
    order = np.argsort(keys // div, kind="stable")
    gi, gv = k.rekey_sort(_dev(keys), None if vals is None else _dev(vals), [], [], div, bits)
    np.testing.assert_array_equal(gi.cpu().numpy(), keys[order])
    if vals is not None:
        np.testing.assert_array_equal(gv.cpu().numpy(), vals[order])


# (sort_divisor, sort_bits): total packed bits = sort_bits + bits(divisor)
MODES = [
    (1, 7), (1, 8), (1, 9), (1, 16), (1, 22), (1, 30), (1, 32),             # 32-bit packed keys
    (1 << 10, 22), (1000, 17), (3, 20),                                     # ... with a low part (pow2 / not)
    (1 << 12, 22), (10 ** 6, 20), (1 << 20, 20), (1 << 30, 10), (7, 33),    # squeeze: 33..40 bits
    (1 << 9, 24), (1 << 16, 24), (1 << 18, 15),
    (3, 40), (1 << 20, 30), (1, 48), (10 ** 6, 40),                         # 64-bit packed keys
]


@pytest.mark.parametrize("div,bits", MODES, ids=lambda x: str(x))
@pytest.mark.parametrize("vdtype", [np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_modes(k, div, bits, vdtype):
    rng = np.random.default_rng(div % 1000 + bits)
    for n in (1, 31, 4096, 4097, 50_000, 300_001):
        s = rng.integers(0, 1 << bits, size=n, dtype=np.int64)
        low = rng.integers(0, div, size=n, dtype=np.int64)
        keys = s * div + low
        vals = rng.standard_normal(n).astype(vdtype)
        _check(k, keys, vals, div, bits)


@pytest.mark.parametrize("vdtype", [None, np.uint8, np.int8, np.bool_, np.int16, np.int32, np.int64],
                         ids=lambda d: "keys_only" if d is None else np.dtype(d).name)
@pytest.mark.parametrize("div,bits", [(1, 22), (1 << 12, 22), (3, 40)], ids=["k32", "squeeze", "wide"])
def test_value_widths(k, vdtype, div, bits):
    rng = np.random.default_rng(bits)
    n = 100_003
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.int64) * div + rng.integers(0, div, size=n, dtype=np.int64)
    vals = None
    if vdtype is not None:
        vals = rng.integers(0, 2 if vdtype == np.bool_ else 100, size=n).astype(vdtype)
    _check(k, keys, vals, div, bits)


def test_unaligned_value_slice(k):
    """Values that do not start on a 16-byte boundary: the TMA bulk copy is replaced by plain loads."""
    import torch
    rng = np.random.default_rng(3)
    n = 70_000
    keys = rng.integers(0, 1 << 20, size=n, dtype=np.int64)
    vals = rng.standard_normal(n + 3).astype(np.float32)
    dv = _dev(vals)[1:n + 1]                      # 4 bytes off
    assert dv.data_ptr() % 16 != 0
    order = np.argsort(keys, kind="stable")
    gi, gv = k.rekey_sort(_dev(keys), dv, [], [], 1, 20)
    np.testing.assert_array_equal(gi.cpu().numpy(), keys[order])
    np.testing.assert_array_equal(gv.cpu().numpy(), vals[1:n + 1][order])
    torch.cuda.synchronize()


def test_few_distinct_digits_and_sorted_input(k):
    """Sorted input (one digit per warp: the histogram's warp-uniform path) and skewed digits."""
    rng = np.random.default_rng(5)
    n = 200_000
    for div, bits in [(1, 24), (1 << 14, 24)]:
        keys = np.sort(rng.integers(0, 1 << bits, size=n, dtype=np.int64)) * div
        _check(k, keys, rng.standard_normal(n).astype(np.float32), div, bits)
        keys = (rng.integers(0, 3, size=n, dtype=np.int64) << (bits - 2)) * div      # three values only
        _check(k, keys, rng.standard_normal(n).astype(np.float32), div, bits)


def test_transposes_large(k):
    """5 M stored elements: 3-D swapaxes(0,2) of config 3's shape family (squeeze mode, 22 of 34 bits) and a
    2-D transpose with non-power-of-two dimensions (config 4's family), against the numpy restatement."""
    from gpu_util import fo
    rng = np.random.default_rng(9)
    for shape, axes in [((4096, 4096, 1024), (2, 1, 0)), ((10 ** 6, 10 ** 6), (1, 0)), ((256, 256, 256, 64), (3, 2, 1, 0))]:
        size = int(np.prod(shape, dtype=object))
        idx = np.unique(rng.integers(0, size, size=5_000_000, dtype=np.int64))
        val = rng.standard_normal(len(idx)).astype(np.float32)
        gi, gv = k.transpose(_dev(idx), _dev(val), shape, axes)
        wi, wv, _ = fo.transpose(idx, val, shape, axes)
        np.testing.assert_array_equal(gi.cpu().numpy(), wi)
        np.testing.assert_array_equal(gv.cpu().numpy(), wv)


def test_reduce_by_key_after_workspace_growth(k):
    """Regression (round-1 review): growing the look-back scratch must not make stale 16-byte entries of
    an earlier launch look live.  Small reduce, then one that forces the workspace to grow, then multi-tile
    reduces at the epochs the old code would have replayed."""
    rng = np.random.default_rng(11)

    def one(n, nkeys):
        keys = np.sort(rng.integers(0, nkeys, size=n, dtype=np.int64))
        vals = rng.integers(-5, 6, size=n).astype(np.float64)
        gk, gs = k.reduce_by_key(_dev(keys), _dev(vals), 1)
        uk, inv = np.unique(keys, return_inverse=True)
        np.testing.assert_array_equal(gk.cpu().numpy(), uk)
        np.testing.assert_array_equal(gs.cpu().numpy(), np.bincount(inv, weights=vals))

    for _ in range(3):
        one(1_000_000, 5000)
    # a sort that needs far more look-back words than anything before: the workspace grows
    big = rng.integers(0, 1 << 30, size=12_000_000, dtype=np.int64)
    gi, _ = k.rekey_sort(_dev(big), None, [], [], 1, 30)
    assert bool((gi[1:] >= gi[:-1]).all().item())
    for _ in range(6):
        one(1_000_000, 5000)
        one(3_000_000, 40)
