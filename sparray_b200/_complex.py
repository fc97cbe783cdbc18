"""complex64 / complex128 FlatSparray (flat.py:19-42 takes any numpy dtype; the reference's test_math.py runs its
element-wise cases on ``dense2d * 1j`` too).

The CUDA kernels are instantiated for real values.  A complex array is held as TWO real FlatSparrays -- the real
and the imaginary parts -- that always share one index array, and every operation is composed from the real
operations on the parts (so every index computation and every flop still runs in the kernels of csrc/):

    a + b     (re_a + re_b) + i (im_a + im_b)                       two union merges, identical index sets
    a * b     (re_a re_b - im_a im_b) + i (re_a im_b + im_a re_b)   four intersection merges, two union merges
    a.T, a[...], a[...] = v, reshape, toarray ...                   once per part

Explicit zeros are kept by every kernel, which is what keeps the two parts aligned.  ``spb_complex_split`` /
``spb_complex_join`` convert between numpy's interleaved storage and the parts.  What numpy computes with its own
complex algorithms and cannot be had bit for bit from real operations is either documented as tolerance-level
(``abs``: sqrt(re^2 + im^2) instead of hypot; division by a scalar) or refused with TypeError (transcendental maps,
ordering comparisons, min / max, dot, dense operands, sum over an axis -- the reference itself raises there:
np.bincount cannot take complex weights).
"""
import numpy as np
import torch

from . import _dtypes as dt
from . import _kernels as _k
from .flat import FlatSparray, _is_scalar, _to_device, _to_device_nd


def _is_complex_like(x):
    if isinstance(x, FlatSparray):
        return x.dtype.kind == "c"
    if isinstance(x, torch.Tensor):
        return x.is_complex()
    return np.iscomplexobj(x)


def _bool_combine(a, b, op):
    """elementwise AND ('minimum') / OR ('maximum') of two bool device vectors, in the int32 kernels"""
    r = _k.binary_map(op, _k.cast(a, torch.int32), _k.cast(b, torch.int32))
    return _k.cast(r, torch.bool)


def _refuse(what):
    raise TypeError("%s is not supported for complex FlatSparrays by the sparray_b200 CUDA kernels" % what)


class ComplexFlatSparray(FlatSparray):
    """See the module docstring.  Constructed by ``FlatSparray(indices, complex_data, ...)``."""

    def __init__(self, indices, data, shape=None, is_canonical=False):
        data = _to_device_nd(data).reshape(-1)
        if not data.is_complex():
            data = _k.complex_join(data, torch.zeros_like(data))
        re, im = _k.complex_split(data.contiguous())
        indices = _to_device(indices, torch.int64)
        # both parts are made canonical the same way (same indices in, same indices out)
        self._re = FlatSparray(indices, re, shape=shape, is_canonical=is_canonical)
        im_sp = FlatSparray(indices, im, shape=shape, is_canonical=is_canonical)
        self._im = FlatSparray(self._re.indices, im_sp.data, shape=self._re.shape, is_canonical=True)
        self.shape = self._re.shape
        self._pending = None
        self._joined = None

    @classmethod
    def _from_parts(cls, re, im):
        """re, im: real FlatSparrays of one float dtype with identical indices"""
        if re.dtype != im.dtype:
            common = np.promote_types(re.dtype, im.dtype)
            re, im = re.astype(common), im.astype(common)
        if re.dtype.kind != "f" or re.dtype.itemsize < 4:
            re, im = re.astype(np.float64), im.astype(np.float64)
        self = cls.__new__(cls)
        self._re, self._im = re, im
        self.shape = re.shape
        self._pending = None
        self._joined = None
        return self

    # -- storage -------------------------------------------------------------------------------------------------
    @property
    def indices(self):
        return self._re.indices

    @indices.setter
    def indices(self, value):
        self._re.indices = value
        self._im.indices = value

    @property
    def data(self):
        if self._joined is None:
            self._joined = _k.complex_join(self._re.data, self._im.data)
        return self._joined

    @data.setter
    def data(self, value):
        re, im = _k.complex_split(_to_device_nd(value).reshape(-1).contiguous())
        self._re.data, self._im.data = re, im
        self._joined = None

    @property
    def dtype(self):
        return np.dtype(np.complex64 if self._re.dtype == np.float32 else np.complex128)

    @property
    def real(self):
        return self._re.copy()

    @property
    def imag(self):
        return self._im.copy()

    def _other_parts(self, other):
        """(re, im) of a FlatSparray operand; the imaginary part of a real one is an aligned array of zeros"""
        if isinstance(other, ComplexFlatSparray):
            return other._re, other._im
        re = other if other.dtype.kind == "f" and other.dtype.itemsize >= 4 else other.astype(np.float64)
        return re, re._with_data(torch.zeros_like(re.data))

    # -- converters ------------------------------------------------------------------------------------------------
    @staticmethod
    def _from_dense(arr):
        shape = tuple(arr.shape)
        dense = _to_device_nd(arr).reshape(-1).contiguous()
        re_d, im_d = _k.complex_split(dense)
        # stored = real part or imaginary part non-zero: union of the two index sets, values gathered at it
        ia, _ = _k.select_nonzero(re_d)
        ib, _ = _k.select_nonzero(im_d)
        idx = _k.union1d_sorted(ia, ib)
        mk = lambda d: FlatSparray(idx, _k.gather(d, idx), shape=shape, is_canonical=True)
        return ComplexFlatSparray._from_parts(mk(re_d), mk(im_d))

    def todense_device(self):
        return _k.complex_join(self._re.todense_device().contiguous(), self._im.todense_device().contiguous())

    def nonzero(self):
        keep_re = _k.unary_map("ne_s", self._re.data, 0)
        keep_im = _k.unary_map("ne_s", self._im.data, 0)
        keep = _bool_combine(keep_re, keep_im, "maximum")
        idx, _ = _k.select_pairs(self.indices, self._re.data, keep)
        return np.unravel_index(idx.cpu().numpy(), self.shape)

    # -- once per part ---------------------------------------------------------------------------------------------
    def _both(self, fn):
        return ComplexFlatSparray._from_parts(fn(self._re), fn(self._im))

    def transpose(self, *axes):
        return self._both(lambda p: p.transpose(*axes))

    def reshape(self, new_shape):
        return self._both(lambda p: p.reshape(new_shape))

    def resize(self, new_shape):
        self._re.resize(new_shape)
        self._im.resize(new_shape)
        self.shape = self._re.shape

    def ravel(self):
        return self._both(lambda p: p.ravel())

    def diagonal(self, offset=0, axis1=0, axis2=1):
        return self._both(lambda p: p.diagonal(offset, axis1, axis2))

    def copy(self):
        return self._both(lambda p: p.copy())

    def __getitem__(self, indices):
        re, im = self._re[indices], self._im[indices]
        if isinstance(re, FlatSparray):
            return ComplexFlatSparray._from_parts(re, im)
        return self.dtype.type(complex(float(re), float(im)))

    def __setitem__(self, indices, val):
        val = np.asarray(val) if not isinstance(val, torch.Tensor) else val.cpu().numpy()
        re_v = np.ascontiguousarray(val.real).astype(self._re.dtype)
        im_v = np.ascontiguousarray(val.imag if np.iscomplexobj(val) else np.zeros_like(val.real)).astype(self._im.dtype)
        if val.ndim == 0:
            re_v, im_v = re_v.item(), im_v.item()
        self._re[indices] = re_v
        self._im[indices] = im_v
        self.shape = self._re.shape
        self._joined = None

    def setdiag(self, values, offset=0):
        values = np.asarray(values)
        self._re.setdiag(np.ascontiguousarray(values.real).astype(self._re.dtype), offset)
        self._im.setdiag(np.ascontiguousarray(values.imag if np.iscomplexobj(values) else np.zeros_like(values.real)).astype(self._im.dtype), offset)
        self._joined = None

    def astype(self, t):
        t = np.dtype(t)
        if t.kind == "c":
            part = np.float32 if t == np.complex64 else np.float64
            return self._both(lambda p: p.astype(part))
        return self._re.astype(t)          # numpy drops the imaginary part (with a ComplexWarning)

    # -- arithmetic --------------------------------------------------------------------------------------------------
    def __neg__(self):
        return self._both(lambda p: -p)

    def conj(self):
        return ComplexFlatSparray._from_parts(self._re.copy(), -self._im)

    def __abs__(self):
        # sqrt(re^2 + im^2): numpy uses hypot (tolerance-level difference, no overflow protection)
        sq = self._re * self._re + self._im * self._im
        return sq.sqrt()

    def __add__(self, other):
        if _is_scalar(other):
            if other == 0:
                return self.copy()
            _refuse("adding a non-zero scalar (a dense result)")
        if not isinstance(other, FlatSparray):
            _refuse("a dense operand")
        ore, oim = self._other_parts(other)
        return ComplexFlatSparray._from_parts(self._re + ore, self._im + oim)

    __radd__ = __add__

    def __sub__(self, other):
        if _is_scalar(other):
            if other == 0:
                return self.copy()
            _refuse("subtracting a non-zero scalar (a dense result)")
        if not isinstance(other, FlatSparray):
            _refuse("a dense operand")
        ore, oim = self._other_parts(other)
        return ComplexFlatSparray._from_parts(self._re - ore, self._im - oim)

    def __rsub__(self, other):
        return (-self).__add__(other)

    def __mul__(self, other):
        if _is_scalar(other):
            s = complex(other)
            if s.imag == 0:
                return self._both(lambda p: p * s.real)
            return ComplexFlatSparray._from_parts(self._re * s.real - self._im * s.imag,
                                                  self._re * s.imag + self._im * s.real)
        if not isinstance(other, FlatSparray):
            _refuse("a dense operand")
        if not isinstance(other, ComplexFlatSparray):
            ore = other if other.dtype.kind == "f" and other.dtype.itemsize >= 4 else other.astype(np.float64)
            return ComplexFlatSparray._from_parts(self._re * ore, self._im * ore)
        return ComplexFlatSparray._from_parts(self._re * other._re - self._im * other._im,
                                              self._re * other._im + self._im * other._re)

    __rmul__ = __mul__

    def multiply(self, other):
        return self.__mul__(other)

    def _divide(self, other, div_func="divide", rdivide=False):
        if rdivide or not _is_scalar(other) or complex(other).imag != 0 or div_func not in ("divide", "true_divide"):
            _refuse("this division")
        s = complex(other).real
        return self._both(lambda p: p / s)     # (numpy multiplies by the reciprocal of the complex scalar: last-ulp differences)

    def __pow__(self, exponent):
        if isinstance(exponent, (int, np.integer)) and exponent > 0:
            out = self
            for _ in range(int(exponent) - 1):
                out = out * self
            return out
        _refuse("this power")

    def _comparison(self, other, method_name, op, op_symbol):
        if op not in ("eq", "ne"):
            raise TypeError("'%s' not supported between complex FlatSparrays" % op_symbol)
        if _is_scalar(other):
            s = complex(other)
            if (op == "eq") == (s == 0):
                _refuse("a comparison with a dense result")
            a = self._re._scalar_map(op + "_s", s.real, compare=True)
            b = self._im._scalar_map(op + "_s", s.imag, compare=True)
        elif isinstance(other, FlatSparray):
            ore, oim = self._other_parts(other)
            a = self._re._pairwise_sparray(ore, op)
            b = self._im._pairwise_sparray(oim, op)
        else:
            _refuse("a dense operand")
        # equal: both parts equal; different: either part different
        return a._with_data(_bool_combine(a.data, b.data, "minimum" if op == "eq" else "maximum"))

    def sum(self, axis=None, dtype=None):
        if axis is not None:
            raise TypeError("Cannot cast array data from dtype('%s') to dtype('float64') according to the rule 'safe'" % self.dtype)
        return self.dtype.type(complex(float(self._re.sum()), float(self._im.sum())))

    def dot(self, other):
        _refuse("dot")

    def minimum(self, other):
        _refuse("minimum")

    def maximum(self, other):
        _refuse("maximum")

    def min(self):
        _refuse("min")

    def max(self):
        _refuse("max")

    def _float_map(self, op):
        _refuse(op)

    def _with_data(self, data):
        out = self.copy()
        out.data = data
        return out

    def to_numpy(self):
        return self.indices.cpu().numpy(), self.data.cpu().numpy()
