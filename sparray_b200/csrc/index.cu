// sparray_b200/csrc/index.cu -- __getitem__ on sorted flat indices (flat.py:254-325,
// 366-393) without ever materialising the index box the reference builds with
// combine_ranges (flat.py:272):
//   * spb_getitem_box    : ints + slices.  One streaming pass over the stored
//     elements (restricted by the host to the flat-index span of the box):
//     unravel, test start/stop/step per axis, write the position inside the box.
//   * spb_outer_sum      : flat query indices of a mixed slice / array index as
//     the outer sum of per-axis offset lists (flat.py:317-323).
//   * spb_lookup_compact : _getitem_flatidx (flat.py:383-393) as one binary search
//     per query cell + order-preserving compaction: O(q log nnz), no sort of the
//     query, so unsorted and repeated fancy indices behave like numpy.
#include "common.cuh"

namespace spb {

constexpr int kBoxDims = 8;
constexpr int kIVT = 8;
constexpr int kITile = 256 * kIVT;

struct BoxParams {
    int ndim;
    FastDiv stride[kBoxDims];     // C-order stride of the axis in the array shape
    int64_t start[kBoxDims], stop[kBoxDims];
    FastDiv step[kBoxDims];
    int64_t out_stride[kBoxDims]; // stride of the axis inside the box (0 for integer-indexed axes)
};

template <typename T>
__global__ void __launch_bounds__(256) getitem_box_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                          int64_t n, const BoxParams b, int64_t* __restrict__ out_idx,
                                                          T* __restrict__ out_val, int64_t* out_count, uint64_t* state,
                                                          uint32_t epoch) {
    const int64_t tile = blockIdx.x;
    const int64_t first = tile * kITile + (int64_t)threadIdx.x * kIVT;
    int64_t pos[kIVT];
    unsigned keep = 0;
#pragma unroll
    for (int k = 0; k < kIVT; ++k) {
        const int64_t i = first + k;
        if (i < n) {
            uint64_t rem = (uint64_t)idx[i];
            int64_t acc = 0;
            bool ok = true;
            for (int ax = 0; ax < b.ndim; ++ax) {
                const uint64_t c = fastdiv(rem, b.stride[ax]);
                rem -= c * b.stride[ax].d;
                const int64_t rel = (int64_t)c - b.start[ax];
                if ((int64_t)c < b.start[ax] || (int64_t)c >= b.stop[ax]) {
                    ok = false;
                    break;
                }
                const uint64_t q = fastdiv((uint64_t)rel, b.step[ax]);
                if ((int64_t)(q * b.step[ax].d) != rel) {
                    ok = false;
                    break;
                }
                acc += (int64_t)q * b.out_stride[ax];
            }
            if (ok) {
                keep |= 1u << k;
                pos[k] = acc;
            }
        }
    }
    int64_t off = compact_offset_256(__popc(keep), state, epoch, tile, tile == gridDim.x - 1, out_count);
#pragma unroll
    for (int k = 0; k < kIVT; ++k) {
        if (keep & (1u << k)) {
            out_idx[off] = pos[k];
            out_val[off] = val[first + k];
            ++off;
        }
    }
}

struct OuterParams {
    int ndim;
    FastDiv stride[kBoxDims];     // C-order stride of output axis j in the result shape
    int64_t base[kBoxDims];       // start of axis j's offset list inside `offsets`
};

__global__ void __launch_bounds__(256) outer_sum_kernel(const int64_t* __restrict__ offsets, const OuterParams o,
                                                        int64_t* __restrict__ out, int64_t n) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        uint64_t rem = (uint64_t)c;
        int64_t acc = 0;
        for (int j = 0; j < o.ndim; ++j) {
            const uint64_t q = fastdiv(rem, o.stride[j]);
            rem -= q * o.stride[j].d;
            acc += __ldg(offsets + o.base[j] + (int64_t)q);
        }
        out[c] = acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) lookup_compact_kernel(const int64_t* __restrict__ sorted, const T* __restrict__ val,
                                                             int64_t n, const int64_t* __restrict__ query, int64_t nq,
                                                             int64_t* __restrict__ out_idx, T* __restrict__ out_val,
                                                             int64_t* out_count, uint64_t* state, uint32_t epoch) {
    const int64_t tile = blockIdx.x;
    const int64_t first = tile * kITile + (int64_t)threadIdx.x * kIVT;
    int64_t hit[kIVT];
    unsigned keep = 0;
#pragma unroll
    for (int k = 0; k < kIVT; ++k) {
        const int64_t j = first + k;
        if (j < nq) {
            const int64_t key = query[j];
            int64_t lo = 0, hi = n;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(sorted + mid) < key) lo = mid + 1;
                else hi = mid;
            }
            if (lo < n && __ldg(sorted + lo) == key) {
                keep |= 1u << k;
                hit[k] = lo;
            }
        }
    }
    int64_t off = compact_offset_256(__popc(keep), state, epoch, tile, tile == gridDim.x - 1, out_count);
#pragma unroll
    for (int k = 0; k < kIVT; ++k) {
        if (keep & (1u << k)) {
            out_idx[off] = first + k;
            out_val[off] = val[hit[k]];
            ++off;
        }
    }
}

static int prep_compact(int64_t n, spb_workspace* ws, cudaStream_t stream, int64_t* ntiles, uint32_t* epoch) {
    *ntiles = ceil_div(n, kITile);
    if (*ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)(*ntiles) * 8, stream);
    if (rc) return rc;
    return workspace_next_epoch(ws, stream, epoch);
}

}  // namespace spb

using namespace spb;

#define SPB_BY_SIZE(dtype, MACRO)                                                   \
    switch (dtype) {                                                                \
        case SPB_F32: case SPB_I32: MACRO(uint32_t);                                \
        case SPB_F64: case SPB_I64: MACRO(uint64_t);                                \
        case SPB_U8: case SPB_BOOL: case SPB_I8: MACRO(uint8_t);                    \
        case SPB_I16: MACRO(uint16_t);                                              \
        default: return SPB_ERR_UNSUPPORTED;                                        \
    }

extern "C" {

int spb_getitem_box(int dtype, const int64_t* idx, const void* val, int64_t n, int ndim, const int64_t* shape,
                    const int64_t* ranges, int64_t* out_idx, void* out_val, int64_t* out_count, spb_workspace* ws,
                    spb_stream_t stream) {
    if (n < 0 || ndim < 1 || ndim > kBoxDims || !shape || !ranges || !out_count) return SPB_ERR_BAD_ARG;
    if (n == 0) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!idx || !val || !out_idx || !out_val) return SPB_ERR_BAD_ARG;
    BoxParams b{};
    b.ndim = ndim;
    int64_t stride = 1, ostride = 1;
    for (int ax = ndim - 1; ax >= 0; --ax) {
        const int64_t start = ranges[3 * ax], stop = ranges[3 * ax + 1], step = ranges[3 * ax + 2];
        if (step < 1) return SPB_ERR_UNSUPPORTED;
        b.stride[ax] = make_fastdiv((uint64_t)stride);
        b.start[ax] = start;
        b.stop[ax] = stop;
        b.step[ax] = make_fastdiv((uint64_t)step);
        b.out_stride[ax] = ostride;
        const int64_t len = spb_len_range(start, stop, step);
        ostride *= len > 0 ? len : 1;
        stride *= shape[ax];
    }
    int64_t ntiles;
    uint32_t epoch;
    int rc = prep_compact(n, ws, stream, &ntiles, &epoch);
    if (rc) return rc;
#define SPB_BOX(T)                                                                                              \
    getitem_box_kernel<T><<<(unsigned)ntiles, 256, 0, stream>>>(idx, (const T*)val, n, b, out_idx, (T*)out_val, \
                                                                out_count, ws_state(ws), epoch);                \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
    SPB_BY_SIZE(dtype, SPB_BOX)
#undef SPB_BOX
}

int spb_outer_sum(const int64_t* offsets, int ndim, const int64_t* lens, int64_t* out, int64_t n,
                  spb_stream_t stream) {
    if (ndim < 1 || ndim > kBoxDims || n < 0 || !lens) return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!offsets || !out) return SPB_ERR_BAD_ARG;
    OuterParams o{};
    o.ndim = ndim;
    int64_t stride = 1, base = 0;
    for (int j = 0; j < ndim; ++j) {
        o.base[j] = base;
        base += lens[j];
    }
    for (int j = ndim - 1; j >= 0; --j) {
        o.stride[j] = make_fastdiv((uint64_t)stride);
        stride *= lens[j];
    }
    int64_t blocks = ceil_div(n, 256);
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    outer_sum_kernel<<<(unsigned)blocks, 256, 0, stream>>>(offsets, o, out, n);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

int spb_lookup_compact(int dtype, const int64_t* sorted, const void* val, int64_t n, const int64_t* query, int64_t nq,
                       int64_t* out_idx, void* out_val, int64_t* out_count, spb_workspace* ws, spb_stream_t stream) {
    if (n < 0 || nq < 0 || !out_count) return SPB_ERR_BAD_ARG;
    if (nq == 0 || n == 0) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!sorted || !val || !query || !out_idx || !out_val) return SPB_ERR_BAD_ARG;
    int64_t ntiles;
    uint32_t epoch;
    int rc = prep_compact(nq, ws, stream, &ntiles, &epoch);
    if (rc) return rc;
#define SPB_LOOK(T)                                                                                                \
    lookup_compact_kernel<T><<<(unsigned)ntiles, 256, 0, stream>>>(sorted, (const T*)val, n, query, nq, out_idx,  \
                                                                   (T*)out_val, out_count, ws_state(ws), epoch);  \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
    SPB_BY_SIZE(dtype, SPB_LOOK)
#undef SPB_LOOK
}

}  // extern "C"

// ---- sparse broadcasting without densifying (flat.py:451-472; the reference densifies there) --------
namespace spb {

struct ExpandParams {
    int ndim;                          // merged axes of the (left-padded) input shape
    FastDiv in_div[kBoxDims];          // C-order stride of the axis in the input shape
    int64_t out_mul[kBoxDims];         // stride of the axis in the output shape (0 for size-1 / broadcast axes)
    int nb;                            // broadcast axes
    FastDiv b_div[kBoxDims];           // C-order stride of broadcast axis j inside the replica box
    int64_t b_mul[kBoxDims];           // its stride in the output shape
    FastDiv reps;                      // replicas per stored element
};

template <typename T>
__global__ void __launch_bounds__(256) broadcast_expand_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                               int64_t n_out, const ExpandParams e,
                                                               int64_t* __restrict__ out_idx, T* __restrict__ out_val) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_out; o += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t i = fastdiv((uint64_t)o, e.reps);
        uint64_t c = (uint64_t)o - i * e.reps.d;
        uint64_t rem = (uint64_t)idx[i];
        int64_t key = 0;
        for (int ax = 0; ax < e.ndim; ++ax) {
            const uint64_t q = fastdiv(rem, e.in_div[ax]);
            rem -= q * e.in_div[ax].d;
            key += (int64_t)q * e.out_mul[ax];
        }
        for (int j = 0; j < e.nb; ++j) {
            const uint64_t q = fastdiv(c, e.b_div[j]);
            c -= q * e.b_div[j].d;
            key += (int64_t)q * e.b_mul[j];
        }
        out_idx[o] = key;
        out_val[o] = val[i];
    }
}

}  // namespace spb

extern "C" int spb_broadcast_expand(int dtype, const int64_t* idx, const void* val, int64_t n, int ndim,
                                    const int64_t* in_shape, const int64_t* out_shape, int64_t* out_idx, void* out_val,
                                    spb_stream_t stream) {
    using namespace spb;
    if (n < 0 || ndim < 1 || ndim > kBoxDims || !in_shape || !out_shape) return SPB_ERR_BAD_ARG;
    ExpandParams e{};
    e.ndim = ndim;
    int64_t in_stride = 1, out_stride = 1, reps = 1;
    int64_t bdim[kBoxDims], bmul[kBoxDims];
    int nb = 0;
    for (int ax = ndim - 1; ax >= 0; --ax) {
        if (in_shape[ax] != out_shape[ax] && in_shape[ax] != 1) return SPB_ERR_BAD_ARG;
        e.in_div[ax] = make_fastdiv((uint64_t)in_stride);
        e.out_mul[ax] = in_shape[ax] == 1 ? 0 : out_stride;
        if (in_shape[ax] == 1 && out_shape[ax] > 1) {
            bdim[nb] = out_shape[ax];
            bmul[nb] = out_stride;
            ++nb;
        }
        in_stride *= in_shape[ax];
        out_stride *= out_shape[ax];
    }
    // broadcast axes were collected innermost first: store them outermost first with C-order strides
    e.nb = nb;
    for (int j = 0; j < nb; ++j) {
        const int src = nb - 1 - j;
        e.b_mul[j] = bmul[src];
    }
    int64_t s = 1;
    for (int j = nb - 1; j >= 0; --j) {
        e.b_div[j] = make_fastdiv((uint64_t)s);
        s *= bdim[nb - 1 - j];
    }
    reps = s;
    e.reps = make_fastdiv((uint64_t)reps);
    const int64_t n_out = n * reps;
    if (n_out == 0) return SPB_OK;
    if (!idx || !val || !out_idx || !out_val) return SPB_ERR_BAD_ARG;
    int64_t blocks = ceil_div(n_out, 256);
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
#define SPB_EXP(T)                                                                                              \
    broadcast_expand_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(idx, (const T*)val, n_out, e, out_idx,    \
                                                                     (T*)out_val);                              \
    SPB_LAUNCHED(1);                                                                                            \
    return (int)cudaGetLastError()
    SPB_BY_SIZE(dtype, SPB_EXP)
#undef SPB_EXP
}
