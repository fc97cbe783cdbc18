"""Host-side dtype tables: numpy <-> torch <-> spb_dtype enum (include/sparray_b200.h).
Pure Python; importable without CUDA."""
import numpy as np
import torch

# enum spb_dtype
SPB_F32, SPB_F64, SPB_I32, SPB_I64, SPB_U8, SPB_BOOL, SPB_I8, SPB_I16, SPB_F16, SPB_U16, SPB_U32, SPB_U64 = range(12)

_TORCH_TO_SPB = {
    torch.float32: SPB_F32, torch.float64: SPB_F64, torch.int32: SPB_I32, torch.int64: SPB_I64,
    torch.uint8: SPB_U8, torch.bool: SPB_BOOL, torch.int8: SPB_I8, torch.int16: SPB_I16,
}
# Conversion-only dtypes (include/sparray_b200.h): the kernels have no instantiations for them.  Arithmetic runs
# on COMPUTE[dtype] (spb_cast there and back: float16 through float32 is exact for one op, numpy itself computes
# float16 in float32; unsigned integers wrap the same modulo 2^16 / 2^32 after narrowing; uint64 shares int64's
# bits for + - * == != and holds for ordering / sums below 2^63); kernels that only MOVE values see them as
# BITS[dtype], a same-sized integer view (no copy).
_CAST_ONLY = {torch.float16: SPB_F16, torch.uint16: SPB_U16, torch.uint32: SPB_U32, torch.uint64: SPB_U64}
COMPUTE = {torch.float16: torch.float32, torch.uint16: torch.int32, torch.uint32: torch.int64, torch.uint64: torch.int64}
BITS = {torch.float16: torch.int16, torch.uint16: torch.int16, torch.uint32: torch.int32, torch.uint64: torch.int64}
_TORCH_TO_NP = {
    torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64,
    torch.uint8: np.uint8, torch.bool: np.bool_, torch.int8: np.int8, torch.int16: np.int16,
    torch.float16: np.float16, torch.uint16: np.uint16, torch.uint32: np.uint32, torch.uint64: np.uint64,
    # containers only (sparray_b200/_complex.py): no kernel takes them, spb_dtype() refuses them
    torch.complex64: np.complex64, torch.complex128: np.complex128,
}
COMPLEX_PART = {torch.complex64: torch.float32, torch.complex128: torch.float64}
_NP_TO_TORCH = {np.dtype(v): k for k, v in _TORCH_TO_NP.items()}

# enum spb_binop
BINOPS = {"add": 0, "sub": 1, "mul": 2, "minimum": 3, "maximum": 4,
          "lt": 5, "le": 6, "gt": 7, "ge": 8, "eq": 9, "ne": 10, "take_a": 11, "take_b": 12, "override": 13,
          "divide": 14, "floor_divide": 15}
COMPARISONS = ("lt", "le", "gt", "ge", "eq", "ne")

# enum spb_unop (same order as the header)
UNOPS = {name: i for i, name in enumerate(
    "copy neg abs mul_s div_s floordiv_s pow_s min_s max_s lt_s le_s gt_s ge_s eq_s ne_s "
    "sin tan arcsin arctan sinh tanh arcsinh arctanh rint sign expm1 log1p deg2rad rad2deg "
    "floor ceil trunc sqrt rdiv_s".split())}


def spb_dtype(torch_dtype):
    try:
        return _TORCH_TO_SPB[torch_dtype]
    except KeyError:
        raise TypeError("dtype %s is not supported by the sparray_b200 CUDA kernels" % (torch_dtype,))


def spb_cast_dtype(torch_dtype):
    """enum value for spb_cast, which also converts the conversion-only dtypes"""
    if torch_dtype in _CAST_ONLY:
        return _CAST_ONLY[torch_dtype]
    return spb_dtype(torch_dtype)


def np_dtype(torch_dtype):
    return np.dtype(_TORCH_TO_NP[torch_dtype])


def torch_dtype(np_dt):
    try:
        return _NP_TO_TORCH[np.dtype(np_dt)]
    except KeyError:
        raise TypeError("dtype %s has no device representation in sparray_b200" % (np_dt,))


def promote(dt_a, dt_b):
    """np.promote_types on torch dtypes (flat.py:415)."""
    return torch_dtype(np.promote_types(np_dtype(dt_a), np_dtype(dt_b)))
