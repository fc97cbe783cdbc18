// sparray_b200/csrc/reduce.cu -- segmented reduction over sorted keys.
//
//   * spb_reduce_by_key : group-by-key float64 sum over an already sorted key
//     stream -- the np.unique + np.bincount of FlatSparray.sum(axis) (flat.py:620-625)
//     and of the non-canonical ctor (flat.py:34-35).  The segment key is
//     idx / divisor, so sum(axis=last) needs no re-keyed copy of the indices.
//     One pass, one launch: coalesced loads -> padded shared-memory transpose ->
//     thread-sequential segmented scan -> block scan -> decoupled look-back that
//     carries (number of segment heads, open-segment partial sum) across tiles.
//   * spb_spmv          : y[row] = sum val * x[col] with row = idx / ncols -- the
//     O(nnz) form of a.dot(x) (flat.py:538-568): warp-aggregated float64 atomics,
//     no shared memory, no look-back.
//   * spb_min_max, spb_select_flagged.
#include "common.cuh"

#include <stdlib.h>

namespace spb {

constexpr int kRNT = 256;
constexpr int kRVT = 8;
constexpr int kRTile = kRNT * kRVT;

// look-back word: [63:62] status, [61:48] epoch, [47] tile contains a head, [46:0] head count
constexpr uint64_t kHeadFlag = 1ull << 47;
constexpr uint64_t kHeadCountMask = (1ull << 47) - 1;

__device__ __forceinline__ void st_release(uint64_t* p, uint64_t v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_cg_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// {tagged look-back word, partial sum} travel together in one aligned 16-byte access
struct alignas(16) Entry16 {
    uint64_t w;
    double v;
};
__device__ __forceinline__ void st_entry(Entry16* p, uint64_t w, double v) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w), "l"(__double_as_longlong(v)) : "memory");
}
__device__ __forceinline__ void ld_entry(const Entry16* p, uint64_t& w, double& v) {
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    w = a;
    v = __longlong_as_double((long long)b);
}

struct Carry {       // (heads, open partial sum, any head) -- the scan operand
    int64_t heads;
    double val;
    bool flag;
};
// a then b (a older): heads add; the open sum restarts at b's last head
__device__ __forceinline__ Carry combine(const Carry& a, const Carry& b) {
    Carry r;
    r.heads = a.heads + b.heads;
    r.val = b.flag ? b.val : a.val + b.val;
    r.flag = a.flag || b.flag;
    return r;
}

struct ReduceParams {
    const int64_t* keys;
    const void* vals;
    int64_t n;
    FastDiv div;            // segment key = keys[i] / div.d
    int64_t* out_keys;      // compact output
    double* out_sums;       // compact (or dense: y[seg])
    int64_t* out_count;
    Entry16* entries;       // [ntiles] look-back entries
    uint32_t epoch;
};

// padded position: one spare slot per 8 elements keeps the blocked reads conflict-free
__device__ __forceinline__ int pad8(int e) { return e + (e >> 3); }

template <typename T>
__global__ void __launch_bounds__(kRNT, 4) reduce_by_key_kernel(const ReduceParams p) {
    __shared__ int64_t s_key[kRTile + kRTile / 8];
    __shared__ T s_val[kRTile + kRTile / 8];
    __shared__ int64_t s_first_seg[kRNT + 1];     // first segment key of each thread (+ next tile's)
    __shared__ int64_t s_last_seg[kRNT];
    __shared__ int s_wheads[kRNT / 32];
    __shared__ double s_wval[kRNT / 32];
    __shared__ int s_wflag[kRNT / 32];
    __shared__ int64_t s_tile_heads;
    __shared__ double s_tile_carry;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t tile_start = tile * kRTile;
    const int tile_items = (int)((p.n - tile_start) < kRTile ? (p.n - tile_start) : kRTile);
    const T* vals = reinterpret_cast<const T*>(p.vals);

    // coalesced striped loads -> padded smem
#pragma unroll
    for (int k = 0; k < kRVT; ++k) {
        const int e = k * kRNT + tid;
        if (e < tile_items) {
            s_key[pad8(e)] = __ldcs(p.keys + tile_start + e);
            s_val[pad8(e)] = __ldcs(vals + tile_start + e);
        }
    }
    // segment key just before / after the tile (heads and ends at the tile borders)
    int64_t prev_seg = -1, next_seg = -1;
    if (tid == 0 && tile_start > 0) prev_seg = (int64_t)fastdiv((uint64_t)p.keys[tile_start - 1], p.div);
    if (tid == kRNT - 1 && tile_start + tile_items < p.n)
        next_seg = (int64_t)fastdiv((uint64_t)p.keys[tile_start + tile_items], p.div);
    __syncthreads();

    // blocked: thread t owns items [t*kRVT, t*kRVT + kRVT)
    const int first = tid * kRVT;
    int cnt = tile_items - first;
    cnt = cnt < 0 ? 0 : (cnt > kRVT ? kRVT : cnt);
    int64_t seg[kRVT];
    double v[kRVT];
#pragma unroll
    for (int k = 0; k < kRVT; ++k) {
        seg[k] = -2;
        v[k] = 0.0;
        if (k < cnt) {
            const int64_t key = s_key[pad8(first + k)];
            seg[k] = (int64_t)fastdiv((uint64_t)key, p.div);
            double val = (double)s_val[pad8(first + k)] + 0.0;     // (a stored -0.0 counts as +0.0, as in np.bincount)
            v[k] = val;
        }
    }
    s_first_seg[tid] = cnt > 0 ? seg[0] : -3;
    s_last_seg[tid] = cnt > 0 ? seg[cnt - 1] : -3;
    if (tid == kRNT - 1) s_first_seg[kRNT] = next_seg;
    __syncthreads();
    // key preceding my first item: the last key of the nearest earlier thread that owns items
    // (all earlier threads are full unless I own nothing), or the previous tile's last key
    const int64_t before = tid == 0 ? prev_seg : s_last_seg[tid - 1];
    // key following my last item (-1: nothing follows, the array ends)
    const int64_t after = (first + cnt == tile_items) ? s_first_seg[kRNT] : s_first_seg[tid + 1];

    // thread-sequential segmented scan
    unsigned head_bits = 0;
    int nheads = 0;
    double run = 0.0;
    double incl[kRVT];
#pragma unroll
    for (int k = 0; k < kRVT; ++k) {
        if (k < cnt) {
            const bool head = seg[k] != (k == 0 ? before : seg[k - 1]);
            if (head) {
                head_bits |= 1u << k;
                ++nheads;
                run = v[k];
            } else {
                run += v[k];
            }
            incl[k] = run;
        }
    }
    // block exclusive scan of Carry(nheads, run, nheads > 0)
    Carry mine{nheads, run, nheads > 0};
    Carry inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Carry up;
        up.heads = __shfl_up_sync(0xffffffffu, inc.heads, o);
        up.val = __shfl_up_sync(0xffffffffu, inc.val, o);
        up.flag = __shfl_up_sync(0xffffffffu, (int)inc.flag, o) != 0;
        if (lane >= o) inc = combine(up, inc);
    }
    if (lane == 31) {
        s_wheads[warp] = (int)inc.heads;
        s_wval[warp] = inc.val;
        s_wflag[warp] = inc.flag;
    }
    __syncthreads();
    // exclusive prefix over earlier warps (8 warps: serial is fine), and the tile aggregate
    Carry wpre{0, 0.0, false};
    Carry tile_agg{0, 0.0, false};
#pragma unroll
    for (int w = 0; w < kRNT / 32; ++w) {
        Carry c{s_wheads[w], s_wval[w], s_wflag[w] != 0};
        if (w < warp) wpre = combine(wpre, c);
        tile_agg = combine(tile_agg, c);
    }
    Carry lane_excl;   // exclusive within the warp
    lane_excl.heads = __shfl_up_sync(0xffffffffu, inc.heads, 1);
    lane_excl.val = __shfl_up_sync(0xffffffffu, inc.val, 1);
    lane_excl.flag = __shfl_up_sync(0xffffffffu, (int)inc.flag, 1) != 0;
    if (lane == 0) lane_excl = Carry{0, 0.0, false};
    Carry excl = combine(wpre, lane_excl);      // everything in this tile before my first item

    // ---- decoupled look-back (warp 0) ----------------------------------------------------------
    if (warp == 0) {
        int64_t heads_before = 0;
        double carry = 0.0;
        if (tile > 0) {
            if (lane == 0)
                st_entry(p.entries + tile,
                         pack_state(kStAgg, p.epoch, (uint64_t)tile_agg.heads | (tile_agg.flag ? kHeadFlag : 0)),
                         tile_agg.val);
            bool val_open = true;       // still extending the open-segment sum backwards
            int64_t base = tile - 1;
            while (base >= 0) {
                const int64_t t = base - lane;
                uint64_t w = pack_state(kStPrefix, p.epoch, 0);
                double pv = 0.0;
                if (t >= 0) {
                    for (;;) {
                        ld_entry(p.entries + t, w, pv);
                        if ((w >> 62) != 0 && (uint32_t)((w >> 48) & ((1u << kEpochBits) - 1)) == p.epoch) break;
                        __nanosleep(20);
                    }
                }
                const unsigned has_prefix = __ballot_sync(0xffffffffu, (w >> 62) == kStPrefix);
                const int stop = has_prefix ? (__ffs(has_prefix) - 1) : 31;
                const unsigned flagged = __ballot_sync(0xffffffffu, (w & kHeadFlag) != 0 && lane <= stop);
                const int vstop = flagged ? (__ffs(flagged) - 1) : stop;   // nearest tile holding a head
                int64_t h = lane <= stop ? (int64_t)(w & kHeadCountMask) : 0;
                double vv = (val_open && lane <= vstop) ? pv : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    h += __shfl_xor_sync(0xffffffffu, h, o);
                    vv += __shfl_xor_sync(0xffffffffu, vv, o);
                }
                heads_before += h;
                carry += vv;
                if (flagged) val_open = false;
                if (has_prefix) break;
                base -= 32;
            }
        }
        if (lane == 0) {
            const Carry incl_prefix = combine(Carry{heads_before, carry, heads_before > 0}, tile_agg);
            st_entry(p.entries + tile,
                     pack_state(kStPrefix, p.epoch, (uint64_t)incl_prefix.heads | (incl_prefix.flag ? kHeadFlag : 0)),
                     incl_prefix.val);
            s_tile_heads = heads_before;
            s_tile_carry = carry;
            if (tile_start + tile_items == p.n) *p.out_count = incl_prefix.heads;
        }
    }
    __syncthreads();
    const int64_t heads_before_tile = s_tile_heads;
    const double tile_carry = s_tile_carry;
    // carry reaching my first item = (tile carry) (+) (in-tile exclusive)
    const double carry_in = excl.flag ? excl.val : tile_carry + excl.val;
    int64_t pos = heads_before_tile + excl.heads - 1;     // output slot of the segment open at my first item
    bool seen_head = false;
#pragma unroll
    for (int k = 0; k < kRVT; ++k) {
        if (k < cnt) {
            if (head_bits & (1u << k)) {
                ++pos;
                seen_head = true;
                p.out_keys[pos] = seg[k];
            }
            const int64_t nxt = (k + 1 < cnt) ? seg[k + 1] : after;
            if (nxt != seg[k]) {   // segment ends here
                const double total = seen_head ? incl[k] : carry_in + incl[k];
                p.out_sums[pos] = total;
            }
        }
    }
}

template <typename T>
__device__ __forceinline__ T mm_op(T a, T b, int is_max) {
    if constexpr (T(0.5) != T(0)) {
        if (a != a) return a;
        if (b != b) return b;
    }
    return is_max ? (a > b ? a : b) : (a < b ? a : b);
}

template <typename T>
__global__ void __launch_bounds__(256) min_max_kernel(const T* __restrict__ in, int64_t n, T* __restrict__ partial,
                                                      int is_max) {
    const int64_t start = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    T acc = in[start < n ? start : 0];
    for (int64_t i = start; i < n; i += (int64_t)gridDim.x * blockDim.x) acc = mm_op<T>(acc, in[i], is_max);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = mm_op<T>(acc, __shfl_xor_sync(0xffffffffu, acc, o), is_max);
    __shared__ T ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = ws[0];
        for (int w = 1; w < 8; ++w) t = mm_op<T>(t, ws[w], is_max);
        partial[blockIdx.x] = t;
    }
}

// compaction of (idx, val) pairs by a byte mask, order preserving
template <typename T>
__global__ void __launch_bounds__(256) select_flagged_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                             const uint8_t* __restrict__ keep, int64_t n,
                                                             int64_t* __restrict__ out_idx, T* __restrict__ out_val,
                                                             int64_t* __restrict__ out_count, uint64_t* state,
                                                             uint32_t epoch) {
    constexpr int VT = 8;
    __shared__ int warp_sums[8];
    __shared__ int64_t tile_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t first = tile * (256 * VT) + (int64_t)tid * VT;
    unsigned bits = 0;
#pragma unroll
    for (int k = 0; k < VT; ++k)
        if (first + k < n && keep[first + k]) bits |= 1u << k;
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < 8 ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < 8) warp_sums[lane] = wi - w;
        const int tile_total = __shfl_sync(0xffffffffu, wi, 7);
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
    for (int k = 0; k < VT; ++k) {
        if (bits & (1u << k)) {
            out_idx[off] = idx[first + k];
            out_val[off] = val[first + k];
            ++off;
        }
    }
}

template <typename T>
int launch_reduce(ReduceParams& p, spb_workspace* ws, cudaStream_t stream) {
    const int64_t ntiles = ceil_div(p.n, kRTile);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, 0, stream);
    if (rc) return rc;
    rc = workspace_reserve_state16(ws, (size_t)ntiles * sizeof(Entry16), stream);
    if (rc) return rc;
    rc = workspace_next_epoch(ws, stream, &p.epoch);
    if (rc) return rc;
    p.entries = reinterpret_cast<Entry16*>(ws->state16);
    {
        static PerDevice configured;
        if (configured.first_use(current_device())) {   // let the SM give this kernel all the shared memory it can use
            SPB_CUDA_TRY(cudaFuncSetAttribute(reduce_by_key_kernel<T>,
                                              cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        }
    }
    reduce_by_key_kernel<T><<<(unsigned)ntiles, kRNT, 0, stream>>>(p);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_reduce_by_key(int dtype, const int64_t* keys, const void* vals, int64_t n, int64_t key_divisor,
                      int64_t* out_keys, double* out_sums, int64_t* out_count, spb_workspace* ws,
                      spb_stream_t stream) {
    if (n < 0 || key_divisor < 1 || !out_count) return SPB_ERR_BAD_ARG;
    if (n == 0) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!keys || !vals || !out_keys || !out_sums) return SPB_ERR_BAD_ARG;
    ReduceParams p{};
    p.keys = keys; p.vals = vals; p.n = n; p.div = make_fastdiv((uint64_t)key_divisor);
    p.out_keys = out_keys; p.out_sums = out_sums; p.out_count = out_count;
#define SPB_RBK(T) return launch_reduce<T>(p, ws, stream)
    switch (dtype) {
        case SPB_F32: SPB_RBK(float);
        case SPB_F64: SPB_RBK(double);
        case SPB_I32: SPB_RBK(int32_t);
        case SPB_I64: SPB_RBK(int64_t);
        case SPB_U8: SPB_RBK(uint8_t);
        case SPB_BOOL: SPB_RBK(uint8_t);
        case SPB_I8: SPB_RBK(int8_t);
        case SPB_I16: SPB_RBK(int16_t);
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_RBK
}

}  // extern "C" (reopened below)

namespace spb {

// ---- SpMV: y[row] += val * x[col], one atomic per run of equal rows inside a warp ---------------
// The stored elements are sorted by flat index, so the elements of a row are contiguous: a warp
// reads 32 consecutive (index, value) pairs coalesced, multiplies by the gathered x, combines runs of
// equal rows with a segmented shuffle scan and lets the last lane of each run add the run's sum into
// y (zeroed by the launch wrapper) with ONE float64 atomic.  No shared memory, no barriers, no
// look-back: ~25 instructions per element row instead of the ~130 of the general segmented reduce.
constexpr int kSpmvUnroll = 4;

template <typename T>
__global__ void __launch_bounds__(256) spmv_rows_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                        int64_t n, const FastDiv div, const T* __restrict__ x,
                                                        double* __restrict__ y, const int64_t nrows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t chunk = 32 * kSpmvUnroll;
    for (int64_t base = warp_global * chunk; base < n; base += nwarps * chunk) {
        int64_t key[kSpmvUnroll];
        T v[kSpmvUnroll];
#pragma unroll
        for (int u = 0; u < kSpmvUnroll; ++u) {
            const int64_t i = base + u * 32 + lane;
            key[u] = i < n ? __ldcs(idx + i) : -1;
            v[u] = i < n ? __ldcs(val + i) : T(0);
        }
        int64_t row[kSpmvUnroll];
        double prod[kSpmvUnroll];
#pragma unroll
        for (int u = 0; u < kSpmvUnroll; ++u) {
            row[u] = -1;
            prod[u] = 0.0;
            if (key[u] >= 0) {
                row[u] = (int64_t)fastdiv((uint64_t)key[u], div);
                const int64_t col = key[u] - row[u] * (int64_t)div.d;
                prod[u] = (double)v[u] * (double)__ldg(x + col);
            }
        }
#pragma unroll
        for (int u = 0; u < kSpmvUnroll; ++u) {
            // segmented inclusive scan over lanes: add the neighbour o lanes below while it is in my row
            const int64_t rprev = __shfl_up_sync(0xffffffffu, row[u], 1);
            const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || rprev != row[u]);
            const int start = 31 - __clz(heads & ((2u << lane) - 1u));       // first lane of my run
            double s = prod[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, s, o);
                if (lane - o >= start) s += up;
            }
            const bool last = lane == 31 || ((heads >> (lane + 1)) & 1u);
            // rows beyond the shape (resize() allows stored indices past prod(shape)) must not touch y
            if (row[u] >= 0 && row[u] < nrows && last) atomicAdd(y + row[u], s);
        }
    }
}

// Variant for LONG rows (>= 8 stored elements per row on average): a lane owns kSpmvRun CONSECUTIVE elements
// (one sector of indices, one of float64 values), multiplies them by their gathered x (kSpmvRun independent
// gathers in flight per lane) and adds up those of its first row; a lane whose elements all lie in one row --
// the usual case -- hands (row, sum) to ONE segmented shuffle scan per 32 * kSpmvRun elements, a lane that
// straddles rows adds its pieces atomically on its own.
#ifndef SPB_SPMV_RUN
#define SPB_SPMV_RUN 4
#endif
constexpr int kSpmvRun = SPB_SPMV_RUN;

template <typename T>
__global__ void __launch_bounds__(256) spmv_runs_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                        int64_t n, const FastDiv div, const T* __restrict__ x,
                                                        double* __restrict__ y, const int64_t nrows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t chunk = 32 * kSpmvRun;
    auto load_chunk = [&](int64_t base, int64_t (&key)[kSpmvRun], T (&v)[kSpmvRun]) {
        const int64_t i0 = base + (int64_t)lane * kSpmvRun;
#pragma unroll
        for (int j = 0; j < kSpmvRun; ++j) {
            key[j] = i0 + j < n ? __ldcs(idx + i0 + j) : -1;
            v[j] = i0 + j < n ? __ldcs(val + i0 + j) : T(0);
        }
    };
    // the next chunk's (index, value) loads are in flight while this chunk's x gathers are
    int64_t key[kSpmvRun], nkey[kSpmvRun];
    T v[kSpmvRun], nv[kSpmvRun];
#ifndef SPB_SPMV_PIPE
#define SPB_SPMV_PIPE 1
#endif
    int64_t base = warp_global * chunk;
    if (SPB_SPMV_PIPE && base < n) load_chunk(base, key, v);
    for (; base < n; base += nwarps * chunk) {
        const int64_t nbase = base + nwarps * chunk;
        if (SPB_SPMV_PIPE) {
            if (nbase < n) load_chunk(nbase, nkey, nv);
        } else {
            load_chunk(base, key, v);
        }
        int64_t row[kSpmvRun];
        double prod[kSpmvRun];
#pragma unroll
        for (int j = 0; j < kSpmvRun; ++j) {
            row[j] = -1;
            prod[j] = 0.0;
            if (key[j] >= 0) {
                const int64_t r = (int64_t)fastdiv((uint64_t)key[j], div);
                const int64_t col = key[j] - r * (int64_t)div.d;
                prod[j] = (double)v[j] * (double)__ldg(x + col);
                row[j] = r < nrows ? r : -1;        // rows beyond the shape (resize()) must not touch y
            }
        }
        bool single = row[0] >= 0;
#pragma unroll
        for (int j = 1; j < kSpmvRun; ++j) single = single && row[j] == row[0];
        int64_t cell = -2 - lane;                   // negative and distinct: takes no part in the warp's runs
        double s = 0.0;
        if (single) {
            cell = row[0];
#pragma unroll
            for (int h = 1; h < kSpmvRun; h <<= 1)            // pairwise tree: (p0 + p1) + (p2 + p3) ...
#pragma unroll
                for (int j = 0; j + h < kSpmvRun; j += 2 * h) prod[j] += prod[j + h];
            s = prod[0];
        } else {
            int64_t run = -1;
            double rs = 0.0;
#pragma unroll
            for (int j = 0; j < kSpmvRun; ++j) {
                if (row[j] >= 0) {
                    if (row[j] != run) {
                        if (run >= 0) atomicAdd(y + run, rs);
                        run = row[j];
                        rs = prod[j];
                    } else {
                        rs += prod[j];
                    }
                }
            }
            if (run >= 0) atomicAdd(y + run, rs);
        }
        const int64_t cprev = __shfl_up_sync(0xffffffffu, cell, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || cprev != cell || cell < 0);
        const int start = 31 - __clz(heads & ((2u << lane) - 1u));       // first lane of my run
        if (heads != 0xffffffffu) {                                       // some run is longer than one lane
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, s, o);
                if (lane - o >= start) s += up;
            }
        }
        const bool last = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (cell >= 0 && last) atomicAdd(y + cell, s);
        if (SPB_SPMV_PIPE) {
#pragma unroll
            for (int j = 0; j < kSpmvRun; ++j) {
                key[j] = nkey[j];
                v[j] = nv[j];
            }
        }
    }
}
static_assert((kSpmvRun & (kSpmvRun - 1)) == 0, "the single-run sum is a pairwise tree");

template <typename T>
int launch_spmv_rows(const int64_t* idx, const void* vals, int64_t n, int64_t ncols, int64_t nrows, const void* x,
                     double* y, cudaStream_t stream) {
    int64_t blocks = ceil_div(n, 256 * kSpmvUnroll);
    const int64_t cap = (int64_t)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (n >= 8 * nrows) {                           // long rows: one segmented scan per 128 elements
        spmv_runs_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(idx, (const T*)vals, n, make_fastdiv((uint64_t)ncols),
                                                                 (const T*)x, y, nrows);
        SPB_LAUNCHED(1);
        return (int)cudaGetLastError();
    }
    spmv_rows_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(idx, (const T*)vals, n, make_fastdiv((uint64_t)ncols),
                                                             (const T*)x, y, nrows);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

}  // namespace spb

extern "C" {

int spb_spmv(int dtype, const int64_t* idx, const void* vals, int64_t n, int64_t ncols, int64_t nrows,
             const void* x, double* y, spb_workspace* ws, spb_stream_t stream) {
    if (n < 0 || ncols < 1 || nrows < 0) return SPB_ERR_BAD_ARG;
    if (nrows == 0) return SPB_OK;
    if (!y) return SPB_ERR_BAD_ARG;
    SPB_CUDA_TRY(cudaMemsetAsync(y, 0, (size_t)nrows * sizeof(double), stream));
    if (n == 0) return SPB_OK;
    if (!idx || !vals || !x) return SPB_ERR_BAD_ARG;
    (void)ws;
    switch (dtype) {
        case SPB_F32: return launch_spmv_rows<float>(idx, vals, n, ncols, nrows, x, y, stream);
        case SPB_F64: return launch_spmv_rows<double>(idx, vals, n, ncols, nrows, x, y, stream);
        default: return SPB_ERR_UNSUPPORTED;
    }
}

int spb_min_max(int dtype, const void* vals, int64_t n, int is_max, void* out, spb_workspace* ws,
                spb_stream_t stream) {
    if (n <= 0 || !vals || !out) return SPB_ERR_BAD_ARG;
    int64_t b = ceil_div(n, 256 * 8);
    const int nblocks = (int)(b > 1024 ? 1024 : b);
    int rc = workspace_reserve(ws, 0, stream);
    if (rc) return rc;
    char* part = ws_misc(ws);
#define SPB_MM(T)                                                                                  \
    min_max_kernel<T><<<nblocks, 256, 0, stream>>>((const T*)vals, n, (T*)part, is_max);           \
    SPB_LAUNCHED(1); \
    min_max_kernel<T><<<1, 256, 0, stream>>>((const T*)part, nblocks, (T*)out, is_max);            \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
    switch (dtype) {
        case SPB_F32: SPB_MM(float);
        case SPB_F64: SPB_MM(double);
        case SPB_I32: SPB_MM(int32_t);
        case SPB_I64: SPB_MM(int64_t);
        case SPB_U8: SPB_MM(uint8_t);
        case SPB_BOOL: SPB_MM(uint8_t);
        case SPB_I8: SPB_MM(int8_t);
        case SPB_I16: SPB_MM(int16_t);
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_MM
}

int spb_select_flagged(int dtype, const int64_t* idx, const void* val, const uint8_t* keep, int64_t n,
                       int64_t* out_idx, void* out_val, int64_t* out_count, spb_workspace* ws,
                       spb_stream_t stream) {
    if (n < 0 || !out_count) return SPB_ERR_BAD_ARG;
    if (n == 0) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!idx || !val || !keep || !out_idx || !out_val) return SPB_ERR_BAD_ARG;
    const int64_t ntiles = ceil_div(n, 256 * 8);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)ntiles * 8, stream);
    if (rc) return rc;
    uint32_t epoch;
    rc = workspace_next_epoch(ws, stream, &epoch);
    if (rc) return rc;
#define SPB_SF(T)                                                                                             \
    select_flagged_kernel<T><<<(unsigned)ntiles, 256, 0, stream>>>(idx, (const T*)val, keep, n, out_idx,     \
                                                                   (T*)out_val, out_count, ws_state(ws), epoch); \
    SPB_LAUNCHED(1); \
    return (int)cudaGetLastError()
    switch (dtype) {
        case SPB_F32: SPB_SF(float);
        case SPB_F64: SPB_SF(double);
        case SPB_I32: SPB_SF(int32_t);
        case SPB_I64: SPB_SF(int64_t);
        case SPB_U8: SPB_SF(uint8_t);
        case SPB_BOOL: SPB_SF(uint8_t);
        case SPB_I8: SPB_SF(int8_t);
        case SPB_I16: SPB_SF(int16_t);
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_SF
}

}  // extern "C"
