"""Operator facade of the sparse array types -- the host-side mirror of the
reference's ``sparray/base.py`` (``_BaseSparray``, base.py:6-117).  Pure
dispatch: every operator forwards to a hook implemented by FlatSparray, which
runs CUDA kernels.  Names, argument meaning and error behaviour follow the
reference so its tests read the same against this class."""
import numpy as np

__all__ = ["is_sparray"]


class _BaseSparray:
    __array_priority__ = 999

    def __len__(self):                                   # base.py:10-12
        return self.shape[0]

    def __bool__(self):                                  # base.py:14-18
        if np.prod(self.shape) <= 1:
            return bool(self.nnz)
        raise ValueError("The truth value of an array with more than one "
                         "element is ambiguous. Use a.any() or a.all().")

    def __iter__(self):                                  # base.py:20-22
        for i in range(self.shape[0]):
            yield self[i]

    # comparisons: base.py:24-40
    def __eq__(self, other):
        return self._comparison(other, "__eq__", "eq", "==")

    def __ne__(self, other):
        return self._comparison(other, "__ne__", "ne", "!=")

    def __lt__(self, other):
        return self._comparison(other, "__lt__", "lt", "<")

    def __le__(self, other):
        return self._comparison(other, "__le__", "le", "<=")

    def __gt__(self, other):
        return self._comparison(other, "__gt__", "gt", ">")

    def __ge__(self, other):
        return self._comparison(other, "__ge__", "ge", ">=")

    __hash__ = None

    def __radd__(self, other):                           # base.py:42-43
        return self.__add__(other)

    def __rsub__(self, other):                           # base.py:48-49
        return (-self).__add__(other)

    def __rmul__(self, other):                           # base.py:51-52
        return self.__mul__(other)

    def multiply(self, other):                           # base.py:54-56
        """Element-wise multiplication. Alias for self * other"""
        return self.__mul__(other)

    # division family: base.py:58-74
    def __div__(self, other):
        return self._divide(other)

    def __rdiv__(self, other):
        return self._divide(other, rdivide=True)

    def __truediv__(self, other):
        return self._divide(other, div_func="true_divide")

    def __rtruediv__(self, other):
        return self._divide(other, div_func="true_divide", rdivide=True)

    def __floordiv__(self, other):
        return self._divide(other, div_func="floor_divide")

    def __rfloordiv__(self, other):
        return self._divide(other, div_func="floor_divide", rdivide=True)

    def __matmul__(self, other):                         # base.py:76-77
        return self.dot(other)

    def mean(self, axis=None, dtype=None):               # base.py:84-100
        if dtype is None:
            dtype = self.dtype
        if np.can_cast(dtype, np.float64):
            dtype = np.float64
        elif np.can_cast(dtype, np.complex128):
            dtype = np.complex128
        s = self.sum(axis=axis, dtype=dtype)
        num_elts = np.prod(self.shape) if axis is None else self.shape[axis]
        if num_elts != 1:
            s /= num_elts
        return s

    def __getattr__(self, attr):                         # base.py:102-111
        if attr == "A":
            return self.toarray()
        if attr == "T":
            return self.transpose()
        if attr == "size":
            return self.getnnz()
        if attr == "ndim":
            return len(self.shape)
        raise AttributeError(attr + " not found")

    def __iadd__(self, other):                           # base.py:113-117
        raise NotImplementedError("in-place add is not supported")

    def __isub__(self, other):
        raise NotImplementedError("in-place subtract is not supported")

    def conjugate(self):
        return self.conj()


def is_sparray(x):                                       # base.py:123
    return isinstance(x, _BaseSparray)
