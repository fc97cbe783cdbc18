// sparray_b200/csrc/spgemm.cu -- sparse x sparse matrix product on the sorted-flat-index layout.
//
// Replaces the reshape -> COO -> scipy CSR matmul -> from_spmatrix detour of FlatSparray.dot with a
// sparse operand (flat.py:545-561).  C = A (m x k) . B (k x n), both given as sorted C-order flat
// indices + values.  Expand - sort - compress:
//   1. the rows of B are contiguous runs of its sorted flat indices: row_ptr by a batched lower bound
//      (spb_lower_bound with the queries j * n);
//   2. spb_spgemm_rowlen : every stored A(i, j) meets len = row_ptr[j+1] - row_ptr[j] elements of B;
//   3. spb_exclusive_scan_i64 : where the products of each A element start (block scan + look-back);
//   4. spb_spgemm_expand : one thread per product finds its A element (binary search of the offsets) and
//      writes (i * n + c, a * b);
//   5. the products are made canonical like any unsorted constructor input: radix sort + segmented
//      reduce (spb_rekey_sort + spb_reduce_by_key), float64 accumulation like flat.py:34-35.
#include "common.cuh"

namespace spb {

__global__ void __launch_bounds__(256) spgemm_rowlen_kernel(const int64_t* __restrict__ a_idx, int64_t na, const FastDiv kdiv,
                                                            const int64_t* __restrict__ row_ptr, int64_t* __restrict__ len) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < na; e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = (uint64_t)a_idx[e];
        const uint64_t j = key - fastdiv(key, kdiv) * kdiv.d;
        len[e] = __ldg(row_ptr + j + 1) - __ldg(row_ptr + j);
    }
}

// exclusive scan of non-negative int64 values: tiles of 2048, block scan, decoupled look-back
constexpr int kScanVT = 8;

__global__ void __launch_bounds__(256) scan_i64_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out,
                                                       int64_t* __restrict__ total, uint64_t* state, uint32_t epoch) {
    __shared__ uint64_t s_warp[8];
    __shared__ uint64_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t first = tile * (256 * kScanVT) + (int64_t)tid * kScanVT;
    uint64_t v[kScanVT];
    uint64_t mine = 0;
#pragma unroll
    for (int k = 0; k < kScanVT; ++k) {
        v[k] = first + k < n ? (uint64_t)in[first + k] : 0;
        mine += v[k];
    }
    uint64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t wofs = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < warp) wofs += s_warp[w];
        tile_total += s_warp[w];
    }
    if (warp == 0) {
        uint64_t base = 0;
        if (tile == 0) {
            if (lane == 0) st_state(state, pack_state(kStPrefix, epoch, tile_total));
        } else {
            if (lane == 0) st_state(state + tile, pack_state(kStAgg, epoch, tile_total));
            base = lookback_exclusive(state, tile, epoch);
            if (lane == 0) st_state(state + tile, pack_state(kStPrefix, epoch, base + tile_total));
        }
        if (lane == 0) {
            s_base = base;
            if (tile == (int64_t)gridDim.x - 1) *total = (int64_t)(base + tile_total);
        }
    }
    __syncthreads();
    uint64_t run = s_base + wofs + incl - mine;
#pragma unroll
    for (int k = 0; k < kScanVT; ++k) {
        if (first + k < n) out[first + k] = (int64_t)run;
        run += v[k];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) spgemm_expand_kernel(const int64_t* __restrict__ a_idx, const T* __restrict__ a_val,
                                                            int64_t na, const FastDiv kdiv, const int64_t* __restrict__ offs,
                                                            const int64_t* __restrict__ row_ptr,
                                                            const int64_t* __restrict__ b_idx, const T* __restrict__ b_val,
                                                            const FastDiv ndiv, int64_t nprod, int64_t* __restrict__ out_idx,
                                                            T* __restrict__ out_val) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nprod; p += (int64_t)gridDim.x * blockDim.x) {
        // the A element whose run of products holds p: last e with offs[e] <= p
        int64_t lo = 0, hi = na;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(offs + mid) <= p) lo = mid + 1;
            else hi = mid;
        }
        const int64_t e = lo - 1;
        const uint64_t akey = (uint64_t)__ldg(a_idx + e);
        const uint64_t i = fastdiv(akey, kdiv);
        const uint64_t j = akey - i * kdiv.d;
        const int64_t bp = __ldg(row_ptr + j) + (p - __ldg(offs + e));
        const uint64_t bkey = (uint64_t)__ldg(b_idx + bp);
        const uint64_t c = bkey - fastdiv(bkey, ndiv) * ndiv.d;
        out_idx[p] = (int64_t)(i * ndiv.d + c);
        out_val[p] = (T)(__ldg(a_val + e) * __ldg(b_val + bp));
    }
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_spgemm_rowlen(const int64_t* a_idx, int64_t na, int64_t k, const int64_t* b_row_ptr, int64_t* len,
                      spb_stream_t stream) {
    if (na < 0 || k < 1) return SPB_ERR_BAD_ARG;
    if (na == 0) return SPB_OK;
    if (!a_idx || !b_row_ptr || !len) return SPB_ERR_BAD_ARG;
    int64_t blocks = ceil_div(na, 256);
    const int64_t cap = (int64_t)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    spgemm_rowlen_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a_idx, na, make_fastdiv((uint64_t)k), b_row_ptr, len);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

int spb_exclusive_scan_i64(const int64_t* in, int64_t n, int64_t* out, int64_t* total, spb_workspace* ws,
                           spb_stream_t stream) {
    if (n < 0 || !total) return SPB_ERR_BAD_ARG;
    if (n == 0) return (int)cudaMemsetAsync(total, 0, 8, stream);
    if (!in || !out || !ws) return SPB_ERR_BAD_ARG;
    const int64_t ntiles = ceil_div(n, 256 * kScanVT);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)ntiles * 8, stream);
    if (rc) return rc;
    uint32_t epoch;
    rc = workspace_next_epoch(ws, stream, &epoch);
    if (rc) return rc;
    scan_i64_kernel<<<(unsigned)ntiles, 256, 0, stream>>>(in, n, out, total, ws_state(ws), epoch);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

int spb_spgemm_expand(int dtype, const int64_t* a_idx, const void* a_val, int64_t na, int64_t k, const int64_t* offsets,
                      const int64_t* b_row_ptr, const int64_t* b_idx, const void* b_val, int64_t n, int64_t nprod,
                      int64_t* out_idx, void* out_val, spb_stream_t stream) {
    if (na < 0 || k < 1 || n < 1 || nprod < 0) return SPB_ERR_BAD_ARG;
    if (nprod == 0) return SPB_OK;
    if (!a_idx || !a_val || !offsets || !b_row_ptr || !b_idx || !b_val || !out_idx || !out_val) return SPB_ERR_BAD_ARG;
    int64_t blocks = ceil_div(nprod, 256);
    const int64_t cap = (int64_t)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const FastDiv kd = make_fastdiv((uint64_t)k), nd = make_fastdiv((uint64_t)n);
#define SPB_EXPAND(T)                                                                                              \
    spgemm_expand_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(a_idx, (const T*)a_val, na, kd, offsets, b_row_ptr, \
                                                                  b_idx, (const T*)b_val, nd, nprod, out_idx,      \
                                                                  (T*)out_val);                                    \
    SPB_LAUNCHED(1);                                                                                               \
    return (int)cudaGetLastError()
    switch (dtype) {
        case SPB_F32: SPB_EXPAND(float);
        case SPB_F64: SPB_EXPAND(double);
        case SPB_I32: SPB_EXPAND(int32_t);
        case SPB_I64: SPB_EXPAND(int64_t);
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_EXPAND
}

}  // extern "C"
