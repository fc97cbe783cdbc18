// sparray_b200/csrc/rekey.cuh -- index re-keying shared by the radix sort (sort.cu) and the dense
// axis sums (axis_sum.cu): unravel a C-order flat index by the strides of (merged) axis groups of the
// old shape, permute / drop groups, re-ravel with the strides of the new shape.  This is the
// np.unravel_index + np.ravel_multi_index pair of FlatSparray.transpose (flat.py:109-111) and
// FlatSparray.sum (flat.py:617-619) without the ndim int64 temporaries.
#pragma once
#include "common.cuh"

namespace spb {

constexpr int kMaxRekeyDims = 8;

struct Rekey {
    int ndim;                             // 0: identity
    int pow2;                             // every in_div and every non-zero out_mul is a power of two
    FastDiv in_div[kMaxRekeyDims];        // C-order stride of (merged) axis group ax in the old shape
    int64_t out_mul[kMaxRekeyDims];       // its stride in the new shape; 0 = group dropped
    // pow2: the groups that survive, as bit fields that move: new |= ((old >> f_in) & f_mask) << f_out
    int nfields;
    int f_in[kMaxRekeyDims];
    int f_out[kMaxRekeyDims];
    uint64_t f_mask[kMaxRekeyDims];       // (size of the group) - 1; all ones for the outermost group
    int out32;                            // pow2 and no inner field reaches bit 32: if the largest new key is below
                                          // 2^32 (the caller's knowledge) the fields move in 32-bit registers
};

template <int NF>
__device__ __forceinline__ int64_t move_fields(uint64_t idx, const Rekey& r) {
    uint64_t acc = 0;
#pragma unroll
    for (int f = 0; f < NF; ++f) acc |= ((idx >> r.f_in[f]) & r.f_mask[f]) << r.f_out[f];
    return (int64_t)acc;
}

// the same with the new key in one 32-bit register (Rekey::out32): half the shifts and masks
template <int NF>
__device__ __forceinline__ uint32_t move_fields32(uint64_t idx, const Rekey& r) {
    uint32_t acc = 0;
#pragma unroll
    for (int f = 0; f < NF; ++f) acc |= ((uint32_t)(idx >> r.f_in[f]) & (uint32_t)r.f_mask[f]) << r.f_out[f];
    return acc;
}

__device__ __forceinline__ uint32_t apply_rekey32(int64_t idx, const Rekey& r) {
    switch (r.nfields) {
        case 0: return 0;
        case 1: return move_fields32<1>((uint64_t)idx, r);
        case 2: return move_fields32<2>((uint64_t)idx, r);
        case 3: return move_fields32<3>((uint64_t)idx, r);
        case 4: return move_fields32<4>((uint64_t)idx, r);
        default: break;
    }
    uint32_t acc = 0;
    for (int f = 0; f < r.nfields; ++f) acc |= ((uint32_t)((uint64_t)idx >> r.f_in[f]) & (uint32_t)r.f_mask[f]) << r.f_out[f];
    return acc;
}

// NDIM known at compile time: straight-line code, no per-axis predicates
template <int NDIM>
__device__ __forceinline__ int64_t apply_rekey_n(uint64_t idx, const Rekey& r) {
    uint64_t rem = idx;
    int64_t acc = 0;
#pragma unroll
    for (int ax = 0; ax < NDIM - 1; ++ax) {
        const uint64_t q = fastdiv(rem, r.in_div[ax]);
        rem -= q * r.in_div[ax].d;
        acc += (int64_t)q * r.out_mul[ax];
    }
    return acc + (int64_t)rem * r.out_mul[NDIM - 1];
}

__device__ __forceinline__ int64_t apply_rekey(int64_t idx, const Rekey& r) {
    if (r.ndim == 0) return idx;
    if (r.pow2) {
        // power-of-two shapes (BASELINE configs 2, 3, 5): every surviving group is a bit field that moves;
        // dropped groups cost nothing
        switch (r.nfields) {
            case 0: return 0;
            case 1: return move_fields<1>((uint64_t)idx, r);
            case 2: return move_fields<2>((uint64_t)idx, r);
            case 3: return move_fields<3>((uint64_t)idx, r);
            case 4: return move_fields<4>((uint64_t)idx, r);
            default: break;
        }
        uint64_t acc = 0;
        for (int f = 0; f < r.nfields; ++f) acc |= (((uint64_t)idx >> r.f_in[f]) & r.f_mask[f]) << r.f_out[f];
        return (int64_t)acc;
    }
    switch (r.ndim) {
        case 1: return apply_rekey_n<1>((uint64_t)idx, r);
        case 2: return apply_rekey_n<2>((uint64_t)idx, r);
        case 3: return apply_rekey_n<3>((uint64_t)idx, r);
        case 4: return apply_rekey_n<4>((uint64_t)idx, r);
        default: break;
    }
    uint64_t rem = (uint64_t)idx;
    int64_t acc = 0;
    for (int ax = 0; ax < r.ndim - 1; ++ax) {
        const uint64_t q = fastdiv(rem, r.in_div[ax]);
        rem -= q * r.in_div[ax].d;
        acc += (int64_t)q * r.out_mul[ax];
    }
    return acc + (int64_t)rem * r.out_mul[r.ndim - 1];
}

// ndim == 0 gives the identity.  Returns SPB_OK or SPB_ERR_BAD_ARG.
inline int make_rekey(Rekey& rk, int ndim, const int64_t* in_strides, const int64_t* out_strides) {
    rk = Rekey{};
    if (ndim < 0 || ndim > kMaxRekeyDims) return SPB_ERR_BAD_ARG;
    if (ndim && (!in_strides || !out_strides)) return SPB_ERR_BAD_ARG;
    rk.ndim = ndim;
    rk.pow2 = 1;
    for (int ax = 0; ax < ndim; ++ax) {
        if (in_strides[ax] < 1 || out_strides[ax] < 0) return SPB_ERR_BAD_ARG;
        rk.in_div[ax] = make_fastdiv((uint64_t)in_strides[ax]);
        rk.out_mul[ax] = out_strides[ax];
        const uint64_t m = (uint64_t)out_strides[ax];
        if (rk.in_div[ax].magic != 0) rk.pow2 = 0;
        if (m && (m & (m - 1))) rk.pow2 = 0;
    }
    if (ndim && in_strides[ndim - 1] != 1) return SPB_ERR_BAD_ARG;
    if (rk.pow2) {
        rk.out32 = ndim > 0;
        for (int ax = 0; ax < ndim; ++ax) {
            const uint64_t m = (uint64_t)out_strides[ax];
            if (!m) continue;                     // dropped group
            int s = 0;
            while ((1ull << s) < m) ++s;
            const int f = rk.nfields++;
            rk.f_in[f] = (int)rk.in_div[ax].shift;
            rk.f_out[f] = s;
            rk.f_mask[f] = ax == 0 ? ~0ull : (uint64_t)(in_strides[ax - 1] / in_strides[ax]) - 1;
            // (the outermost group has no size here -- its mask is all ones: the caller must know that the
            // largest new key is below 2^32 before it moves the fields in 32 bits)
            if (s >= 32 || (ax != 0 && (rk.f_mask[f] << s) >= (1ull << 32))) rk.out32 = 0;
        }
    }
    return SPB_OK;
}

}  // namespace spb
