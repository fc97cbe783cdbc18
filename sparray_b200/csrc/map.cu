// sparray_b200/csrc/map.cu -- bandwidth-bound elementwise kernels on the value array
// (the `_with_data` family, flat.py:580-581,677-739), casts, dense <-> sparse
// conversion (flat.py:48-54,74-77), gathers and batched binary search
// (flat.py:262-267).  128-bit vectorised where the pointers allow it.
#include "common.cuh"

#include <cuda_fp16.h>

#include <type_traits>

#include <math.h>

namespace spb {

constexpr int kMapThreads = 256;

inline unsigned grid_for(int64_t work_items, int per_block) {
    int64_t b = ceil_div(work_items, per_block);
    const int64_t cap = (int64_t)device_sm_count() * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// numpy floor_divide: for floats npy_divmod (fmod-based, so it stays exact where
// floor(x / y) would round across an integer); for ints floor toward -inf, x // 0 = 0.
template <typename T>
__device__ __forceinline__ T floor_div(T x, T y) {
    if constexpr (T(0.5) != T(0)) {
        if (y == T(0)) return x / y;
        T mod = fmod(x, y);
        T div = (x - mod) / y;
        if (mod != T(0) && ((y < T(0)) != (mod < T(0)))) div -= T(1);
        if (div != T(0)) {
            T fl = floor(div);
            if (div - fl > T(0.5)) fl += T(1);
            return fl;
        }
        return copysign(T(0), x / y);
    } else {
        if (y == 0) return 0;
        T q = x / y;
        if ((x % y != 0) && ((x < 0) != (y < 0))) --q;
        return q;
    }
}

template <typename T>
__device__ __forceinline__ T sign_of(T x) {
    if constexpr (T(0.5) != T(0)) {
        return x != x ? x : (T)((x > 0) - (x < 0));
    } else {
        return (T)((x > 0) - (x < 0));
    }
}

template <typename T>
struct UnaryCtx {
    int op;
    T s;          // scalar operand in the array dtype
};

// float-only transcendental / rounding maps
template <typename T>
__device__ __forceinline__ T float_map(int op, T x) {
    switch (op) {
        case SPB_U_SIN: return sin(x);
        case SPB_U_TAN: return tan(x);
        case SPB_U_ARCSIN: return asin(x);
        case SPB_U_ARCTAN: return atan(x);
        case SPB_U_SINH: return sinh(x);
        case SPB_U_TANH: return tanh(x);
        case SPB_U_ARCSINH: return asinh(x);
        case SPB_U_ARCTANH: return atanh(x);
        case SPB_U_RINT: return rint(x);
        case SPB_U_EXPM1: return expm1(x);
        case SPB_U_LOG1P: return log1p(x);
        case SPB_U_DEG2RAD: return x * (T)(M_PI / 180.0);
        case SPB_U_RAD2DEG: return x * (T)(180.0 / M_PI);
        case SPB_U_FLOOR: return floor(x);
        case SPB_U_CEIL: return ceil(x);
        case SPB_U_TRUNC: return trunc(x);
        case SPB_U_SQRT: return sqrt(x);
        default: return x;
    }
}

template <typename T>
__device__ __forceinline__ T int_pow(T x, T e) {
    T r = 1;
    while (e > 0) {
        if (e & 1) r *= x;
        x *= x;
        e >>= 1;
    }
    return r;
}

template <typename T>
__device__ __forceinline__ T arith_map(const UnaryCtx<T>& c, T x) {
    constexpr bool is_float = T(0.5) != T(0);
    switch (c.op) {
        case SPB_U_COPY: return x;
        case SPB_U_NEG: return (T)(-x);
        case SPB_U_ABS:
            if constexpr (is_float) return fabs(x);
            else return x < 0 ? (T)(-x) : x;
        case SPB_U_MUL_S: return (T)(x * c.s);
        case SPB_U_DIV_S:
            if constexpr (is_float) return x / c.s;
            else return floor_div<T>(x, c.s);
        case SPB_U_RDIV_S:
            if constexpr (is_float) return c.s / x;
            else return floor_div<T>(c.s, x);
        case SPB_U_FLOORDIV_S: return floor_div<T>(x, c.s);
        case SPB_U_POW_S:
            if constexpr (is_float) return c.s == T(2) ? x * x : (T)pow(x, c.s);
            else return int_pow<T>(x, c.s);
        case SPB_U_MIN_S:
            if constexpr (is_float) return x != x ? x : (x < c.s ? x : c.s);
            else return x < c.s ? x : c.s;
        case SPB_U_MAX_S:
            if constexpr (is_float) return x != x ? x : (x > c.s ? x : c.s);
            else return x > c.s ? x : c.s;
        case SPB_U_SIGN: return sign_of<T>(x);
        default:
            if constexpr (is_float) return float_map<T>(c.op, x);
            else return x;
    }
}

template <typename T>
__device__ __forceinline__ uint8_t compare_map(const UnaryCtx<T>& c, T x) {
    switch (c.op) {
        case SPB_U_LT_S: return x < c.s;
        case SPB_U_LE_S: return x <= c.s;
        case SPB_U_GT_S: return x > c.s;
        case SPB_U_GE_S: return x >= c.s;
        case SPB_U_EQ_S: return x == c.s;
        default: return x != c.s;
    }
}

template <typename T, typename TO, bool CMP>
__global__ void __launch_bounds__(kMapThreads) unary_kernel(const T* __restrict__ in, TO* __restrict__ out, int64_t n,
                                                            UnaryCtx<T> c, int vec_ok) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t done = 0;
    if (vec_ok) {
        const int64_t nvec = n / VEC;
        const uint4* in4 = reinterpret_cast<const uint4*>(in);
        for (int64_t v = tid; v < nvec; v += nthreads) {
            uint4 raw = __ldcs(in4 + v);
            T x[VEC];
            memcpy(x, &raw, 16);
            TO y[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if constexpr (CMP) y[k] = (TO)compare_map<T>(c, x[k]);
                else y[k] = (TO)arith_map<T>(c, x[k]);
            }
            if constexpr (sizeof(TO) == sizeof(T)) {
                uint4 o;
                memcpy(&o, y, 16);
                __stcs(reinterpret_cast<uint4*>(out) + v, o);
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) out[v * VEC + k] = y[k];
            }
        }
        done = nvec * VEC;
    }
    for (int64_t i = done + tid; i < n; i += nthreads) {
        if constexpr (CMP) out[i] = (TO)compare_map<T>(c, in[i]);
        else out[i] = (TO)arith_map<T>(c, in[i]);
    }
}

// out[i] = a[i] (op) b[i] on two aligned value arrays: the dense (op) sparse helpers of
// flat.py:436-449 after the dense operand has been gathered at the stored indices
template <typename T, typename TO, bool CMP>
__global__ void __launch_bounds__(kMapThreads) binary_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                             TO* __restrict__ out, int64_t n, int op) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T x = a[i], y = b[i];
        if constexpr (CMP) {
            uint8_t r;
            switch (op) {
                case SPB_LT: r = x < y; break;
                case SPB_LE: r = x <= y; break;
                case SPB_GT: r = x > y; break;
                case SPB_GE: r = x >= y; break;
                case SPB_EQ: r = x == y; break;
                default: r = x != y; break;
            }
            out[i] = (TO)r;
        } else {
            T r;
            switch (op) {
                case SPB_ADD: r = x + y; break;
                case SPB_SUB: r = x - y; break;
                case SPB_MUL: r = x * y; break;
                case SPB_MINIMUM:
                    if constexpr (T(0.5) != T(0)) r = x != x ? x : (x < y ? x : y);
                    else r = x < y ? x : y;
                    break;
                case SPB_MAXIMUM:
                    if constexpr (T(0.5) != T(0)) r = x != x ? x : (x > y ? x : y);
                    else r = x > y ? x : y;
                    break;
                case SPB_DIVIDE:
                    if constexpr (T(0.5) != T(0)) r = x / y;
                    else r = floor_div<T>(x, y);
                    break;
                case SPB_FLOOR_DIVIDE: r = floor_div<T>(x, y); break;
                case SPB_TAKE_A: r = x; break;
                default: r = y; break;
            }
            out[i] = (TO)r;
        }
    }
}

// value conversion like numpy's astype: one rounding (double -> half goes through __double2half, not through float)
template <typename TI, typename TO>
__device__ __forceinline__ TO convert_value(TI x) {
    if constexpr (std::is_same<TI, __half>::value && std::is_same<TO, __half>::value) {
        return x;
    } else if constexpr (std::is_same<TI, __half>::value) {
        return (TO)__half2float(x);
    } else if constexpr (std::is_same<TO, __half>::value) {
        if constexpr (std::is_same<TI, double>::value) return __double2half(x);
        else return __float2half_rn((float)x);        // integers beyond 2^24 are far past the largest half anyway
    } else {
        return (TO)x;
    }
}
template <typename TI>
__device__ __forceinline__ bool is_nonzero(TI x) {
    if constexpr (std::is_same<TI, __half>::value) return __half2float(x) != 0.0f;
    else return x != TI(0);
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kMapThreads) cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n,
                                                           int to_bool) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const TI x = in[i];
        out[i] = to_bool ? (TO)(is_nonzero<TI>(x) ? 1 : 0) : convert_value<TI, TO>(x);
    }
}

// complex storage (re, im interleaved) <-> two real arrays: the only complex-aware kernels -- every complex
// operation is composed from the real kernels on the two parts (sparray_b200/_complex.py)
template <typename T>
__global__ void __launch_bounds__(kMapThreads) complex_split_kernel(const T* __restrict__ c, int64_t n, T* __restrict__ re,
                                                                    T* __restrict__ im) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        re[i] = c[2 * i];
        im[i] = c[2 * i + 1];
    }
}
template <typename T>
__global__ void __launch_bounds__(kMapThreads) complex_join_kernel(const T* __restrict__ re, const T* __restrict__ im,
                                                                   int64_t n, T* __restrict__ c) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        c[2 * i] = re[i];
        c[2 * i + 1] = im[i];
    }
}

template <typename T>
__global__ void __launch_bounds__(kMapThreads) scatter_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                              int64_t n, T* __restrict__ dense) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dense[idx[i]] = val[i];
}

template <typename T>
__global__ void __launch_bounds__(kMapThreads) gather_kernel(const T* __restrict__ src, const int64_t* __restrict__ pos,
                                                             int64_t n, T* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = src[pos[i]];
}

// out_pos[q] = first i with sorted[i] >= query[q]; found[q] = sorted[i] == query[q]
__global__ void __launch_bounds__(kMapThreads) lower_bound_kernel(const int64_t* __restrict__ sorted, int64_t n,
                                                                  const int64_t* __restrict__ query, int64_t nq,
                                                                  int64_t* __restrict__ out_pos,
                                                                  uint8_t* __restrict__ found) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t key = query[q];
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (__ldg(sorted + mid) < key) lo = mid + 1;
            else hi = mid;
        }
        out_pos[q] = lo;
        if (found) found[q] = (lo < n && __ldg(sorted + lo) == key) ? 1 : 0;
    }
}

// ---- dense -> sparse: keep elements != 0 (flat.py:48-54), single pass with look-back ----------
constexpr int kSelVT = 8;
constexpr int kSelTile = kMapThreads * kSelVT;

template <typename T>
__global__ void __launch_bounds__(kMapThreads) select_nonzero_kernel(const T* __restrict__ dense, int64_t n,
                                                                     int64_t* __restrict__ out_idx, T* __restrict__ out_val,
                                                                     int64_t* __restrict__ out_count, uint64_t* state,
                                                                     uint32_t epoch) {
    __shared__ int warp_sums[kMapThreads / 32];
    __shared__ int64_t tile_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t first = tile * kSelTile + (int64_t)tid * kSelVT;
    T v[kSelVT];
    unsigned keep = 0;
#pragma unroll
    for (int k = 0; k < kSelVT; ++k) {
        int64_t i = first + k;
        if (i < n) {
            v[k] = dense[i];
            if (v[k] != T(0)) keep |= 1u << k;
        }
    }
    const int cnt = __popc(keep);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kMapThreads / 32 ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < kMapThreads / 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kMapThreads / 32) warp_sums[lane] = wi - w;
        const int tile_total = __shfl_sync(0xffffffffu, wi, kMapThreads / 32 - 1);
        uint64_t base = 0;
        if (tile == 0) {
            if (lane == 0) st_state(state, pack_state(kStPrefix, epoch, (uint64_t)tile_total));
        } else {
            if (lane == 0) st_state(state + tile, pack_state(kStAgg, epoch, (uint64_t)tile_total));
            base = lookback_exclusive(state, tile, epoch);
            if (lane == 0) st_state(state + tile, pack_state(kStPrefix, epoch, base + (uint64_t)tile_total));
        }
        if (lane == 0) {
            tile_base_s = (int64_t)base;
            if (tile == gridDim.x - 1) *out_count = (int64_t)base + tile_total;
        }
    }
    __syncthreads();
    int64_t off = tile_base_s + warp_sums[warp] + incl - cnt;
#pragma unroll
    for (int k = 0; k < kSelVT; ++k) {
        if (keep & (1u << k)) {
            out_idx[off] = first + k;
            out_val[off] = v[k];
            ++off;
        }
    }
}

// ---- full reduction of the value array (flat.py:610: data.sum(dtype)) ---------------------------
template <typename T, typename ACC>
__global__ void __launch_bounds__(kMapThreads) sum_all_kernel(const T* __restrict__ in, int64_t n, ACC* __restrict__ partial) {
    ACC acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += (ACC)in[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ ACC ws[kMapThreads / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        ACC t = 0;
        for (int w = 0; w < kMapThreads / 32; ++w) t += ws[w];
        partial[blockIdx.x] = t;
    }
}

template <typename ACC>
__global__ void sum_partials_kernel(const ACC* __restrict__ partial, int nblocks, ACC* __restrict__ out) {
    ACC acc = 0;
    for (int i = threadIdx.x; i < nblocks; i += 32) acc += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) *out = acc;
}

template <typename T, typename TO, bool CMP>
int launch_unary(int op, const void* in, void* out, int64_t n, double sf, int64_t si, cudaStream_t stream) {
    UnaryCtx<T> c;
    c.op = op;
    if constexpr (T(0.5) != T(0)) c.s = (T)sf;
    else c.s = (T)si;
    const int vec_ok = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    constexpr int VEC = 16 / (int)sizeof(T);
    unary_kernel<T, TO, CMP><<<grid_for(n, kMapThreads * VEC * 2), kMapThreads, 0, stream>>>(
        (const T*)in, (TO*)out, n, c, vec_ok);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

}  // namespace spb

using namespace spb;

#define SPB_DISPATCH_DTYPE(dtype, MACRO)          \
    switch (dtype) {                              \
        case SPB_F32: MACRO(float);               \
        case SPB_F64: MACRO(double);              \
        case SPB_I32: MACRO(int32_t);             \
        case SPB_I64: MACRO(int64_t);             \
        case SPB_U8: MACRO(uint8_t);              \
        case SPB_BOOL: MACRO(uint8_t);            \
        case SPB_I8: MACRO(int8_t);               \
        case SPB_I16: MACRO(int16_t);             \
        default: return SPB_ERR_UNSUPPORTED;      \
    }

extern "C" {

int spb_unary_map(int op, int dtype, const void* in, void* out, int64_t n, double scalar_f, int64_t scalar_i,
                  spb_stream_t stream) {
    if (n < 0 || op < 0 || op >= SPB_UNOP_COUNT) return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!in || !out) return SPB_ERR_BAD_ARG;
    const bool cmp = op >= SPB_U_LT_S && op <= SPB_U_NE_S;
    const bool float_only = (op >= SPB_U_SIN && op <= SPB_U_RAD2DEG && op != SPB_U_SIGN) || op == SPB_U_SQRT ||
                            op == SPB_U_FLOOR || op == SPB_U_CEIL || op == SPB_U_TRUNC;
    if (float_only && dtype != SPB_F32 && dtype != SPB_F64) return SPB_ERR_UNSUPPORTED;
#define SPB_UNARY(T)                                                                                       \
    return cmp ? launch_unary<T, uint8_t, true>(op, in, out, n, scalar_f, scalar_i, stream)                \
               : launch_unary<T, T, false>(op, in, out, n, scalar_f, scalar_i, stream)
    switch (dtype) {
        case SPB_F32: SPB_UNARY(float);
        case SPB_F64: SPB_UNARY(double);
        case SPB_I32: SPB_UNARY(int32_t);
        case SPB_I64: SPB_UNARY(int64_t);
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_UNARY
}

int spb_binary_map(int op, int dtype, const void* a, const void* b, void* out, int64_t n, spb_stream_t stream) {
    if (n < 0) return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!a || !b || !out) return SPB_ERR_BAD_ARG;
    const bool cmp = op >= SPB_LT && op <= SPB_NE;
    const unsigned grid = grid_for(n, kMapThreads * 4);
#define SPB_BIN(T)                                                                                          \
    if (cmp) binary_kernel<T, uint8_t, true><<<grid, kMapThreads, 0, stream>>>((const T*)a, (const T*)b,    \
                                                                               (uint8_t*)out, n, op);       \
    else binary_kernel<T, T, false><<<grid, kMapThreads, 0, stream>>>((const T*)a, (const T*)b, (T*)out, n, op); \
    SPB_LAUNCHED(1);                                                                                         \
    return (int)cudaGetLastError()
    switch (dtype) {
        case SPB_F32: SPB_BIN(float);
        case SPB_F64: SPB_BIN(double);
        case SPB_I32: SPB_BIN(int32_t);
        case SPB_I64: SPB_BIN(int64_t);
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_BIN
}

int spb_cast(int dtype_in, int dtype_out, const void* in, void* out, int64_t n, spb_stream_t stream) {
    if (n < 0) return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!in || !out) return SPB_ERR_BAD_ARG;
    const unsigned grid = grid_for(n, kMapThreads * 4);
    const int to_bool = dtype_out == SPB_BOOL;
#define SPB_CAST_OUT(TO) \
    cast_kernel<TI, TO><<<grid, kMapThreads, 0, stream>>>((const TI*)in, (TO*)out, n, to_bool); \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
#define SPB_CAST_IN(TI_)                                  \
    {                                                     \
        using TI = TI_;                                   \
        switch (dtype_out) {                              \
            case SPB_F32: SPB_CAST_OUT(float);            \
            case SPB_F64: SPB_CAST_OUT(double);           \
            case SPB_I32: SPB_CAST_OUT(int32_t);          \
            case SPB_I64: SPB_CAST_OUT(int64_t);          \
            case SPB_U8: SPB_CAST_OUT(uint8_t);           \
            case SPB_BOOL: SPB_CAST_OUT(uint8_t);         \
            case SPB_I8: SPB_CAST_OUT(int8_t);            \
            case SPB_I16: SPB_CAST_OUT(int16_t);          \
            case SPB_F16: SPB_CAST_OUT(__half);           \
            case SPB_U16: SPB_CAST_OUT(uint16_t);         \
            case SPB_U32: SPB_CAST_OUT(uint32_t);         \
            case SPB_U64: SPB_CAST_OUT(uint64_t);         \
            default: return SPB_ERR_UNSUPPORTED;          \
        }                                                 \
    }
    switch (dtype_in) {                                   /* the conversion-only dtypes first */
        case SPB_F16: SPB_CAST_IN(__half);
        case SPB_U16: SPB_CAST_IN(uint16_t);
        case SPB_U32: SPB_CAST_IN(uint32_t);
        case SPB_U64: SPB_CAST_IN(uint64_t);
        default: break;
    }
    SPB_DISPATCH_DTYPE(dtype_in, SPB_CAST_IN)
#undef SPB_CAST_IN
#undef SPB_CAST_OUT
}

int spb_complex_split(int real_dtype, const void* c, int64_t n, void* re, void* im, spb_stream_t stream) {
    if (n < 0) return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!c || !re || !im) return SPB_ERR_BAD_ARG;
    const unsigned grid = grid_for(n, kMapThreads * 4);
    if (real_dtype == SPB_F32) complex_split_kernel<float><<<grid, kMapThreads, 0, stream>>>((const float*)c, n, (float*)re, (float*)im);
    else if (real_dtype == SPB_F64) complex_split_kernel<double><<<grid, kMapThreads, 0, stream>>>((const double*)c, n, (double*)re, (double*)im);
    else return SPB_ERR_UNSUPPORTED;
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

int spb_complex_join(int real_dtype, const void* re, const void* im, int64_t n, void* c, spb_stream_t stream) {
    if (n < 0) return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!c || !re || !im) return SPB_ERR_BAD_ARG;
    const unsigned grid = grid_for(n, kMapThreads * 4);
    if (real_dtype == SPB_F32) complex_join_kernel<float><<<grid, kMapThreads, 0, stream>>>((const float*)re, (const float*)im, n, (float*)c);
    else if (real_dtype == SPB_F64) complex_join_kernel<double><<<grid, kMapThreads, 0, stream>>>((const double*)re, (const double*)im, n, (double*)c);
    else return SPB_ERR_UNSUPPORTED;
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

int spb_scatter_dense(int dtype, const int64_t* idx, const void* val, int64_t n, void* dense, spb_stream_t stream) {
    if (n < 0) return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!idx || !val || !dense) return SPB_ERR_BAD_ARG;
    const unsigned grid = grid_for(n, kMapThreads * 4);
#define SPB_SCATTER(T) \
    scatter_kernel<T><<<grid, kMapThreads, 0, stream>>>(idx, (const T*)val, n, (T*)dense); \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
    SPB_DISPATCH_DTYPE(dtype, SPB_SCATTER)
#undef SPB_SCATTER
}

int spb_gather(int dtype, const void* src, const int64_t* pos, int64_t n, void* out, spb_stream_t stream) {
    if (n < 0) return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!src || !pos || !out) return SPB_ERR_BAD_ARG;
    const unsigned grid = grid_for(n, kMapThreads * 4);
#define SPB_GATHER(T) \
    gather_kernel<T><<<grid, kMapThreads, 0, stream>>>((const T*)src, pos, n, (T*)out); \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
    SPB_DISPATCH_DTYPE(dtype, SPB_GATHER)
#undef SPB_GATHER
}

int spb_lower_bound(const int64_t* sorted, int64_t n, const int64_t* query, int64_t nq, int64_t* out_pos,
                    uint8_t* found, spb_stream_t stream) {
    if (n < 0 || nq < 0) return SPB_ERR_BAD_ARG;
    if (nq == 0) return SPB_OK;
    if ((n && !sorted) || !query || !out_pos) return SPB_ERR_BAD_ARG;
    lower_bound_kernel<<<grid_for(nq, kMapThreads), kMapThreads, 0, stream>>>(sorted, n, query, nq, out_pos, found);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

int spb_select_nonzero(int dtype, const void* dense, int64_t n, int64_t* out_idx, void* out_val, int64_t* out_count,
                       spb_workspace* ws, spb_stream_t stream) {
    if (n < 0 || !out_count) return SPB_ERR_BAD_ARG;
    if (n == 0) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!dense || !out_idx || !out_val) return SPB_ERR_BAD_ARG;
    const int64_t ntiles = ceil_div(n, kSelTile);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)ntiles * 8, stream);
    if (rc) return rc;
    uint32_t epoch;
    rc = workspace_next_epoch(ws, stream, &epoch);
    if (rc) return rc;
#define SPB_SELECT(T)                                                                                          \
    select_nonzero_kernel<T><<<(unsigned)ntiles, kMapThreads, 0, stream>>>((const T*)dense, n, out_idx, (T*)out_val, \
                                                                           out_count, ws_state(ws), epoch);  \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
    SPB_DISPATCH_DTYPE(dtype, SPB_SELECT)
#undef SPB_SELECT
}

// out: device scalar of type double (float dtypes and acc_f64 != 0) or int64 (integer dtypes)
int spb_sum_all(int dtype, const void* val, int64_t n, void* out, spb_workspace* ws, spb_stream_t stream) {
    if (n < 0 || !out) return SPB_ERR_BAD_ARG;
    if (n == 0) return (int)cudaMemsetAsync(out, 0, 8, stream);
    const int nblocks = (int)grid_for(n, kMapThreads * 8);
    int rc = workspace_reserve(ws, 0, stream);
    if (rc) return rc;
    char* part = ws_misc(ws);
#define SPB_SUM(T, ACC)                                                                              \
    sum_all_kernel<T, ACC><<<nblocks, kMapThreads, 0, stream>>>((const T*)val, n, (ACC*)part);      \
    SPB_LAUNCHED(1); \
    sum_partials_kernel<ACC><<<1, 32, 0, stream>>>((const ACC*)part, nblocks, (ACC*)out);           \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
    switch (dtype) {
        case SPB_F32: SPB_SUM(float, double);
        case SPB_F64: SPB_SUM(double, double);
        case SPB_I32: SPB_SUM(int32_t, int64_t);
        case SPB_I64: SPB_SUM(int64_t, int64_t);
        case SPB_U8: SPB_SUM(uint8_t, int64_t);
        case SPB_BOOL: SPB_SUM(uint8_t, int64_t);
        case SPB_I8: SPB_SUM(int8_t, int64_t);
        case SPB_I16: SPB_SUM(int16_t, int64_t);
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_SUM
}

}  // extern "C"
