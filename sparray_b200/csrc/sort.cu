// sparray_b200/csrc/sort.cu -- index re-keying (unravel / permute or drop axes /
// re-ravel) fused into a one-sweep LSD radix sort of (int64 key, value) pairs.
//
// Replaces the np.unravel_index / np.ravel_multi_index temporaries and the
// argsort-based np.unique of FlatSparray.transpose (flat.py:109-112 + ctor :34),
// of sum(axis != last) (flat.py:617-620) and of the non-canonical ctor.
//
//   histogram kernel : one read of the keys, all digit histograms at once
//   one kernel per 8-bit digit: tile-local ranking by warp ballots (stable),
//     per-digit decoupled look-back across tiles, exchange through shared memory,
//     run-coalesced scatter.  The re-key arithmetic (shifts or mul-hi magic
//     division, no 64-bit div instruction) is applied to the keys as they are
//     loaded by the histogram and by the first pass, so the re-keyed index array
//     is never materialised on its own.
//   Only the digits that matter are sorted: the host passes a divisor that strips
//   the trailing axes which already arrive in order (LSD sort is stable).
#include "common.cuh"

namespace spb {

constexpr int kMaxRekeyDims = 8;
constexpr int kSNT = 256;                 // threads (= radix, one digit per thread in the scan phases)
constexpr int kSIPT = 16;                 // items per thread
constexpr int kSTile = kSNT * kSIPT;      // 4096
constexpr int kSWarps = kSNT / 32;
constexpr int kMaxPasses = 8;

struct Rekey {
    int ndim;                             // 0: identity
    FastDiv in_div[kMaxRekeyDims];        // C-order stride of (merged) axis ax in the old shape
    int64_t out_mul[kMaxRekeyDims];       // its stride in the new shape; 0 = axis dropped
};

__device__ __forceinline__ int64_t apply_rekey(int64_t idx, const Rekey& r) {
    if (r.ndim == 0) return idx;
    uint64_t rem = (uint64_t)idx;
    int64_t acc = 0;
#pragma unroll
    for (int ax = 0; ax < kMaxRekeyDims - 1; ++ax) {
        if (ax < r.ndim - 1) {
            const uint64_t q = fastdiv(rem, r.in_div[ax]);
            rem -= q * r.in_div[ax].d;
            acc += (int64_t)q * r.out_mul[ax];
        }
    }
    return acc + (int64_t)rem * r.out_mul[r.ndim - 1];
}

struct SortParams {
    const int64_t* keys_in;
    const void* vals_in;
    int64_t* keys_out;
    void* vals_out;
    int64_t n;
    Rekey rk;
    int do_rekey;
    FastDiv sort_div;       // sort key = key / sort_div.d
    int shift;              // this pass sorts bits [shift, shift + 8) of the sort key
    int npasses;
    uint64_t* hist;         // [kMaxPasses][256] counts, then exclusive offsets
    uint64_t* state;        // [ntiles][256]
    uint32_t epoch;
};

__device__ __forceinline__ uint64_t sort_key(int64_t key, const SortParams& p) {
    return fastdiv((uint64_t)key, p.sort_div);
}

// ---- all digit histograms in one read --------------------------------------------------------
__global__ void __launch_bounds__(kSNT) radix_hist_kernel(const SortParams p) {
    __shared__ unsigned h[kMaxPasses * 256];
    for (int i = threadIdx.x; i < p.npasses * 256; i += kSNT) h[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nthreads = (int64_t)gridDim.x * kSNT;
    // whole warps stay converged so the vote below is legal: iterate on a warp-aligned bound
    const int64_t n_round = (p.n + 31) & ~(int64_t)31;
    for (int64_t i = blockIdx.x * (int64_t)kSNT + threadIdx.x; i < n_round; i += nthreads) {
        const bool valid = i < p.n;
        uint64_t sk = 0;
        if (valid) {
            int64_t key = __ldcs(p.keys_in + i);
            if (p.do_rekey) key = apply_rekey(key, p.rk);
            sk = sort_key(key, p);
        }
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        for (int ps = 0; ps < p.npasses; ++ps) {
            const unsigned d = (unsigned)(sk >> (8 * ps)) & 255u;
            const unsigned d0 = __shfl_sync(0xffffffffu, d, __ffs(vmask) - 1);
            const bool uniform = __all_sync(0xffffffffu, !valid || d == d0);
            if (uniform) {   // runs of equal digits (sorted input): one add per warp
                if (lane == __ffs(vmask) - 1) atomicAdd(&h[ps * 256 + d], (unsigned)__popc(vmask));
            } else if (valid) {
                atomicAdd(&h[ps * 256 + d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.npasses * 256; i += kSNT)
        if (h[i]) atomicAdd(reinterpret_cast<unsigned long long*>(p.hist) + i, (unsigned long long)h[i]);
}

// counts -> exclusive offsets, one block per pass
__global__ void __launch_bounds__(256) radix_scan_kernel(uint64_t* hist) {
    __shared__ uint64_t s[256];
    uint64_t* h = hist + blockIdx.x * 256;
    const int t = threadIdx.x;
    const uint64_t mine = h[t];
    s[t] = mine;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        uint64_t add = t >= o ? s[t - o] : 0;
        __syncthreads();
        s[t] += add;
        __syncthreads();
    }
    h[t] = s[t] - mine;
}

// ---- one digit pass -----------------------------------------------------------------------------
template <typename V, bool HAS_VAL>
__global__ void __launch_bounds__(kSNT, 3) radix_pass_kernel(const SortParams p) {
    __shared__ int64_t s_keys[kSTile];                 // keys, then values, in digit-sorted tile order
    __shared__ uint8_t s_dig[kSTile];
    __shared__ unsigned s_whist[kSWarps][256];
    __shared__ int s_dstart[256];
    __shared__ int64_t s_goff[256];
    __shared__ int s_scan[kSWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * kSTile;
    const int items = (int)((p.n - base) < kSTile ? (p.n - base) : kSTile);
    const int wbase = warp * 32 * kSIPT;

    for (int i = tid; i < kSWarps * 256; i += kSNT) (&s_whist[0][0])[i] = 0;

    int64_t keys[kSIPT];
    uint8_t dig[kSIPT];
    uint16_t rank[kSIPT];
#pragma unroll
    for (int k = 0; k < kSIPT; ++k) {
        const int e = wbase + k * 32 + lane;
        keys[k] = 0;
        if (e < items) {
            int64_t key = __ldcs(p.keys_in + base + e);
            if (p.do_rekey) key = apply_rekey(key, p.rk);
            keys[k] = key;
        }
    }
    __syncthreads();

    // tile-local stable ranking: per warp, item rows in order, lanes in order inside a row
#pragma unroll
    for (int k = 0; k < kSIPT; ++k) {
        const int e = wbase + k * 32 + lane;
        const bool valid = e < items;
        const unsigned d = (unsigned)(sort_key(keys[k], p) >> p.shift) & 255u;
        dig[k] = (uint8_t)d;
        // lanes holding the same digit (invalid lanes get a private value): one MATCH instead of 8 votes
        unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (0x100u | (unsigned)lane));
        if (!valid) peers = 0;
        const int leader = peers ? (__ffs(peers) - 1) : lane;
        unsigned old = 0;
        if (valid && lane == leader) {
            old = s_whist[warp][d];
            s_whist[warp][d] = old + (unsigned)__popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[k] = (uint16_t)(old + (unsigned)__popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    // per digit (thread d): exclusive offsets across warps, tile count, look-back
    {
        const int d = tid;
        unsigned sum = 0;
#pragma unroll
        for (int w = 0; w < kSWarps; ++w) {
            const unsigned t = s_whist[w][d];
            s_whist[w][d] = sum;
            sum += t;
        }
        uint64_t* st = p.state + tile * 256 + d;
        if (tile == 0) st_state(st, pack_state(kStPrefix, p.epoch, sum));
        else st_state(st, pack_state(kStAgg, p.epoch, sum));

        // block exclusive scan of `sum` over the 256 digits
        int incl = (int)sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int wofs = 0;
#pragma unroll
        for (int w = 0; w < kSWarps; ++w)
            if (w < warp) wofs += s_scan[w];
        const int dstart = wofs + incl - (int)sum;

        uint64_t excl = 0;
        if (tile > 0) {
            // per-digit look-back; the words of 8 predecessor tiles are requested together so the walk
            // costs one round trip per 8 tiles instead of one per tile
            int64_t t = tile - 1;
            bool done = false;
            while (!done && t >= 0) {
                uint64_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int64_t tt = t - j;
                    w[j] = tt >= 0 ? ld_state(p.state + tt * 256 + d) : pack_state(kStPrefix, p.epoch, 0);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!done) {
                        const int64_t tt = t - j;
                        while ((w[j] >> 62) == 0 || (uint32_t)((w[j] >> 48) & ((1u << kEpochBits) - 1)) != p.epoch) {
                            __nanosleep(20);
                            w[j] = ld_state(p.state + tt * 256 + d);
                        }
                        excl += w[j] & kCountMask;
                        done = (w[j] >> 62) == kStPrefix;
                    }
                }
                t -= 8;
            }
            st_state(st, pack_state(kStPrefix, p.epoch, excl + sum));
        }
        s_dstart[d] = dstart;
        s_goff[d] = (int64_t)(p.hist[(p.shift >> 3) * 256 + d] + excl) - dstart;
    }
    __syncthreads();

    // exchange keys through shared memory into digit order
#pragma unroll
    for (int k = 0; k < kSIPT; ++k) {
        const int e = wbase + k * 32 + lane;
        if (e < items) {
            const int r = s_dstart[dig[k]] + (int)s_whist[warp][dig[k]] + (int)rank[k];
            rank[k] = (uint16_t)r;
            s_keys[r] = keys[k];
            s_dig[r] = dig[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSIPT; ++j) {
        const int q = j * kSNT + tid;
        if (q < items) p.keys_out[s_goff[s_dig[q]] + q] = s_keys[q];
    }
    if constexpr (HAS_VAL) {
        const V* vin = reinterpret_cast<const V*>(p.vals_in);
        V* vout = reinterpret_cast<V*>(p.vals_out);
        V vals[kSIPT];
#pragma unroll
        for (int k = 0; k < kSIPT; ++k) {
            const int e = wbase + k * 32 + lane;
            if (e < items) vals[k] = __ldcs(vin + base + e);
        }
        __syncthreads();                               // all key reads of s_keys are done
        V* s_vals = reinterpret_cast<V*>(s_keys);
#pragma unroll
        for (int k = 0; k < kSIPT; ++k) {
            const int e = wbase + k * 32 + lane;
            if (e < items) s_vals[rank[k]] = vals[k];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kSIPT; ++j) {
            const int q = j * kSNT + tid;
            if (q < items) vout[s_goff[s_dig[q]] + q] = s_vals[q];
        }
    }
}

// no digits to sort: just re-key (and copy the values)
template <typename V>
__global__ void __launch_bounds__(256) rekey_copy_kernel(const SortParams p) {
    const V* vin = reinterpret_cast<const V*>(p.vals_in);
    V* vout = reinterpret_cast<V*>(p.vals_out);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t key = p.keys_in[i];
        if (p.do_rekey) key = apply_rekey(key, p.rk);
        p.keys_out[i] = key;
        if (vin) vout[i] = vin[i];
    }
}

template <typename V>
int run_sort(SortParams p, int sort_bits, int64_t* tmp_keys, void* tmp_vals, spb_workspace* ws, cudaStream_t stream) {
    const int npasses = (sort_bits + 7) / 8;
    int64_t* const final_keys = p.keys_out;
    void* const final_vals = p.vals_out;
    if (npasses == 0) {
        int64_t blocks = ceil_div(p.n, 256 * 4);
        if (blocks > 148 * 16) blocks = 148 * 16;
        rekey_copy_kernel<V><<<(unsigned)blocks, 256, 0, stream>>>(p);
        SPB_LAUNCHED(1);
        return (int)cudaGetLastError();
    }
    if (npasses > kMaxPasses) return SPB_ERR_BAD_ARG;
    if (npasses > 1 && (!tmp_keys || (p.vals_in && !tmp_vals))) return SPB_ERR_BAD_ARG;
    const int64_t ntiles = ceil_div(p.n, kSTile);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)ntiles * 256 * 8, stream);
    if (rc) return rc;
    p.npasses = npasses;
    p.hist = reinterpret_cast<uint64_t*>(ws_misc(ws) + (512 << 10));     // upper half of the misc slab
    p.state = ws_state(ws);
    SPB_CUDA_TRY(cudaMemsetAsync(p.hist, 0, sizeof(uint64_t) * kMaxPasses * 256, stream));
    int64_t hblocks = ceil_div(p.n, kSNT * 16);
    if (hblocks > 148 * 8) hblocks = 148 * 8;
    radix_hist_kernel<<<(unsigned)hblocks, kSNT, 0, stream>>>(p);
    SPB_LAUNCHED(1);
    radix_scan_kernel<<<npasses, 256, 0, stream>>>(p.hist);
    SPB_LAUNCHED(1);
    const bool has_val = p.vals_in != nullptr;
    {
        static bool configured = false;
        if (!configured) {   // 47 KB static shared memory per CTA: ask for the full carve-out
            SPB_CUDA_TRY(cudaFuncSetAttribute(radix_pass_kernel<V, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            SPB_CUDA_TRY(cudaFuncSetAttribute(radix_pass_kernel<V, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            configured = true;
        }
    }
    for (int ps = 0; ps < npasses; ++ps) {
        const bool to_final = ((npasses - 1 - ps) % 2) == 0;
        p.keys_out = to_final ? final_keys : tmp_keys;
        p.vals_out = to_final ? final_vals : tmp_vals;
        p.shift = 8 * ps;
        rc = workspace_next_epoch(ws, stream, &p.epoch);
        if (rc) return rc;
        if (has_val) {
            radix_pass_kernel<V, true><<<(unsigned)ntiles, kSNT, 0, stream>>>(p);
        } else {
            radix_pass_kernel<V, false><<<(unsigned)ntiles, kSNT, 0, stream>>>(p);
        }
        SPB_LAUNCHED(1);
        SPB_CUDA_TRY(cudaGetLastError());
        p.keys_in = p.keys_out;
        p.vals_in = p.vals_out;
        p.do_rekey = 0;
    }
    return SPB_OK;
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_rekey_sort(int dtype, const int64_t* idx, const void* val, int64_t n, int ndim,
                   const int64_t* in_strides, const int64_t* out_strides, int64_t sort_divisor, int sort_bits,
                   int64_t* out_idx, void* out_val, int64_t* tmp_idx, void* tmp_val, spb_workspace* ws,
                   spb_stream_t stream) {
    if (n < 0 || ndim < 0 || ndim > kMaxRekeyDims || sort_divisor < 1 || sort_bits < 0 || sort_bits > 63)
        return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!idx || !out_idx || (val && !out_val) || (ndim && (!in_strides || !out_strides))) return SPB_ERR_BAD_ARG;
    SortParams p{};
    p.keys_in = idx; p.vals_in = val; p.keys_out = out_idx; p.vals_out = out_val; p.n = n;
    p.rk.ndim = ndim;
    p.do_rekey = ndim > 0;
    for (int ax = 0; ax < ndim; ++ax) {
        if (in_strides[ax] < 1 || out_strides[ax] < 0) return SPB_ERR_BAD_ARG;
        p.rk.in_div[ax] = make_fastdiv((uint64_t)in_strides[ax]);
        p.rk.out_mul[ax] = out_strides[ax];
    }
    if (ndim && in_strides[ndim - 1] != 1) return SPB_ERR_BAD_ARG;
    p.sort_div = make_fastdiv((uint64_t)sort_divisor);
    switch (dtype) {
        case SPB_F32: case SPB_I32: return run_sort<uint32_t>(p, sort_bits, tmp_idx, tmp_val, ws, stream);
        case SPB_F64: case SPB_I64: return run_sort<uint64_t>(p, sort_bits, tmp_idx, tmp_val, ws, stream);
        case SPB_U8: case SPB_BOOL: case SPB_I8: return run_sort<uint8_t>(p, sort_bits, tmp_idx, tmp_val, ws, stream);
        case SPB_I16: return run_sort<uint16_t>(p, sort_bits, tmp_idx, tmp_val, ws, stream);
        default: return SPB_ERR_UNSUPPORTED;
    }
}

}  // extern "C"

// ---- axis sums into a dense accumulator (no sort) ---------------------------------------------------
// When the reduced array prod(new_shape) is small enough to hold densely, sum(axis != last)
// (flat.py:617-625) does not need the re-key + radix sort at all: every stored element adds its
// value into acc[new_key] with a float64 atomic (the accumulator is L2-resident for the BASELINE
// configs), marks the cell as touched (explicit zeros must survive, so "sum == 0" cannot be the
// test), and one compaction pass over the cells emits the touched ones in key order.
namespace spb {

// A warp reads 32 consecutive elements (coalesced), re-keys them, combines runs of equal new keys with a
// segmented shuffle scan (when the reduced axis is the last one the elements of an output cell are
// contiguous; otherwise runs are rare and the scan is a few wasted instructions) and the last lane of
// each run adds the run's sum into its cell with one float64 atomic.
constexpr int kAccUnroll = 4;

template <typename T>
__global__ void __launch_bounds__(256) axis_accumulate_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                              int64_t n, const int64_t* __restrict__ n_dev,
                                                              const Rekey rk, double* __restrict__ acc,
                                                              uint8_t* __restrict__ touched) {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();   // operands (and n_dev) may come from the previous launch
    if (n_dev) {                       // the producer's count never left the device; n is its upper bound
        const int64_t m = *n_dev;
        n = m < n ? m : n;
    }
    const int lane = threadIdx.x & 31;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    if (n <= nthreads) {
        // few elements (typically a lazily counted product whose real count is far below its upper bound):
        // one element per thread, no aggregation -- the shortest critical path
        if (gtid < n) {
            const int64_t key = apply_rekey(__ldcs(idx + gtid), rk);
            atomicAdd(acc + key, (double)__ldcs(val + gtid));
            touched[key] = 1;
        }
        return;
    }
    const int64_t warp_global = gtid >> 5;
    const int64_t nwarps = nthreads >> 5;
    const int64_t chunk = 32 * kAccUnroll;
    for (int64_t base = warp_global * chunk; base < n; base += nwarps * chunk) {
        int64_t key[kAccUnroll];
        double v[kAccUnroll];
#pragma unroll
        for (int u = 0; u < kAccUnroll; ++u) {
            const int64_t i = base + u * 32 + lane;
            key[u] = i < n ? __ldcs(idx + i) : -1;
            v[u] = i < n ? (double)__ldcs(val + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kAccUnroll; ++u)
            if (key[u] >= 0) key[u] = apply_rekey(key[u], rk);
#pragma unroll
        for (int u = 0; u < kAccUnroll; ++u) {
            // runs = maximal stretches of ADJACENT lanes with the same key (equal keys further apart are
            // separate runs and simply cost one more atomic)
            const int64_t kprev = __shfl_up_sync(0xffffffffu, key[u], 1);
            const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || kprev != key[u]);
            const int start = 31 - __clz(heads & ((2u << lane) - 1u));       // first lane of my run
            double s = v[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, s, o);
                if (lane - o >= start) s += up;
            }
            const bool last = lane == 31 || ((heads >> (lane + 1)) & 1u);
            if (key[u] >= 0 && last) {
                atomicAdd(acc + key[u], s);
                touched[key[u]] = 1;               // a flag per cell: explicit zeros survive
            }
        }
    }
}

// Variant for LONG runs (reduced axis innermost and many elements per output cell): a lane owns kAccItems
// CONSECUTIVE elements (one or two sectors), a warp 128.  The lane re-keys its
// elements and adds up those that fall into the same cell as its first one as long as they are adjacent
// (when the reduced axis is the last one the elements of a cell are contiguous, so a lane usually holds
// ONE run); a lane holding a single run hands (cell, sum) to one segmented shuffle scan per 128 elements,
// whose run ends issue one float64 atomic each; a lane holding several runs adds its pieces atomically
// on its own.
constexpr int kAccItems = 4;

__device__ __forceinline__ double warp_run_sum(int64_t cell, double s, int lane, bool& last) {
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
    last = lane == 31 || ((heads >> (lane + 1)) & 1u);
    return s;
}

template <typename T>
__global__ void __launch_bounds__(256) axis_accumulate_runs_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                              int64_t n, const int64_t* __restrict__ n_dev,
                                                              const Rekey rk, double* __restrict__ acc,
                                                              uint8_t* __restrict__ touched) {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();   // operands (and n_dev) may come from the previous launch
    if (n_dev) {                       // the producer's count never left the device; n is its upper bound
        const int64_t m = *n_dev;
        n = m < n ? m : n;
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t chunk = 32 * kAccItems;
    for (int64_t base = warp_global * chunk; base < n; base += nwarps * chunk) {
        const int64_t i0 = base + (int64_t)lane * kAccItems;
        int64_t key[kAccItems];
        double v[kAccItems];
#pragma unroll
        for (int j = 0; j < kAccItems; ++j) {
            key[j] = i0 + j < n ? __ldcs(idx + i0 + j) : -1;
            v[j] = i0 + j < n ? (double)__ldcs(val + i0 + j) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < kAccItems; ++j)
            if (key[j] >= 0) key[j] = apply_rekey(key[j], rk);
        // my runs of adjacent equal cells: all but a single whole-lane run go in atomically right here
        const bool single = key[0] >= 0 && key[0] == key[1] && key[1] == key[2] && key[2] == key[3];
        int64_t cell = -2 - lane;      // negative and distinct: takes no part in the warp's runs
        double s = 0.0;
        if (single) {
            cell = key[0];
            s = (v[0] + v[1]) + (v[2] + v[3]);
        } else {
            int64_t run = -1;
            double rs = 0.0;
#pragma unroll
            for (int j = 0; j < kAccItems; ++j) {
                if (key[j] >= 0) {
                    if (key[j] != run) {
                        if (run >= 0) {
                            atomicAdd(acc + run, rs);
                            touched[run] = 1;
                        }
                        run = key[j];
                        rs = v[j];
                    } else {
                        rs += v[j];
                    }
                }
            }
            if (run >= 0) {
                atomicAdd(acc + run, rs);
                touched[run] = 1;       // a flag per cell: explicit zeros survive
            }
        }
        bool last;
        s = warp_run_sum(cell, s, lane, last);
        if (cell >= 0 && last) {
            atomicAdd(acc + cell, s);
            touched[cell] = 1;
        }
    }
}

// Emits the touched cells in key order and puts every cell it emitted back to zero, so the
// accumulator slab is all-zero again when the call is over (no memset per call).  One thread owns
// 64 cells (four 16-byte loads of flag bytes); a tile of 256 threads covers 16 K cells, which keeps
// the look-back chain short.
__global__ void __launch_bounds__(256) select_touched_kernel(double* __restrict__ acc, uint8_t* __restrict__ touched,
                                                             int64_t cells, int64_t key_base, int64_t* __restrict__ out_idx,
                                                             double* __restrict__ out_val, int64_t* out_count,
                                                             uint64_t* state, uint32_t epoch) {
    const int64_t tile = blockIdx.x;
    const int64_t first = (tile * 256 + threadIdx.x) * 64;
    cudaGridDependencySynchronize();   // accumulators come from the previous launch
    unsigned long long bits = 0;
    if (first < cells) {               // the flag array is padded to a multiple of 64 (the padding stays zero)
        uint4* const flags = reinterpret_cast<uint4*>(touched + first);
        uint4 raw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) raw[q] = flags[q];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned w[4] = {raw[q].x, raw[q].y, raw[q].z, raw[q].w};
#pragma unroll
            for (int k = 0; k < 16; ++k)
                bits |= (unsigned long long)((w[k >> 2] >> (8 * (k & 3))) & 1u) << (16 * q + k);
            if (raw[q].x | raw[q].y | raw[q].z | raw[q].w) flags[q] = make_uint4(0, 0, 0, 0);
        }
    }
    const int cnt = __popcll(bits);
    const int64_t off = compact_offset_256<true>(cnt, state, epoch, tile, tile == gridDim.x - 1, out_count);
    // Warp-cooperative emission: the warp's outputs are one contiguous run, written 32 at a time (coalesced
    // whether the cells are nearly all touched or nearly all empty).  Output r of the warp belongs to the
    // first lane whose inclusive count exceeds r, and is that lane's (r - exclusive count)-th set bit.
    const int lane = threadIdx.x & 31;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    const int64_t warp_off = __shfl_sync(0xffffffffu, off, 0);
    const int64_t warp_first = first - (int64_t)lane * 64;
    for (int c = 0; c < warp_total; c += 32) {
        const int r = c + lane;
        int owner = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int probe = __shfl_sync(0xffffffffu, incl, owner + step - 1);
            if (probe <= r) owner += step;
        }
        owner = owner > 31 ? 31 : owner;
        const int owner_excl = __shfl_sync(0xffffffffu, incl - cnt, owner);
        const unsigned lo_bits = __shfl_sync(0xffffffffu, (unsigned)bits, owner);
        const unsigned hi_bits = __shfl_sync(0xffffffffu, (unsigned)(bits >> 32), owner);
        if (r < warp_total) {
            const int j = r - owner_excl, nlo = __popc(lo_bits);
            const int k = j < nlo ? (int)__fns(lo_bits, 0, j + 1) : 32 + (int)__fns(hi_bits, 0, j - nlo + 1);
            const int64_t cell = warp_first + (int64_t)owner * 64 + k;
            out_idx[warp_off + r] = key_base + cell;
            out_val[warp_off + r] = acc[cell];
            acc[cell] = 0.0;
        }
    }
}

}  // namespace spb

namespace spb {

static int make_rekey(Rekey& rk, int ndim, const int64_t* in_strides, const int64_t* out_strides) {
    if (ndim < 1 || ndim > kMaxRekeyDims || !in_strides || !out_strides) return SPB_ERR_BAD_ARG;
    rk.ndim = ndim;
    for (int ax = 0; ax < ndim; ++ax) {
        if (in_strides[ax] < 1 || out_strides[ax] < 0) return SPB_ERR_BAD_ARG;
        rk.in_div[ax] = make_fastdiv((uint64_t)in_strides[ax]);
        rk.out_mul[ax] = out_strides[ax];
    }
    return SPB_OK;
}

static int launch_accumulate(int dtype, const int64_t* idx, const void* val, int64_t n, const int64_t* n_dev,
                             const Rekey& rk, int64_t cells, double* acc, uint8_t* touched, cudaStream_t stream) {
    int64_t blocks = ceil_div(n, 256 * kAccUnroll);
    if (blocks > 148 * 16) blocks = 148 * 16;
    // long contiguous runs (innermost axis reduced, >= 8 elements per output cell on average): the variant
    // that pre-reduces inside a lane and scans once per 128 elements
    const bool long_runs = rk.ndim >= 1 && rk.out_mul[rk.ndim - 1] == 0 && n >= 8 * cells;
#define SPB_ACC(T)                                                                                            \
    if (long_runs) {                                                                                          \
        SPB_CUDA_TRY(launch_pdl(axis_accumulate_runs_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, stream, \
                                idx, (const T*)val, n, n_dev, rk, acc, touched));                             \
    } else {                                                                                                  \
        SPB_CUDA_TRY(launch_pdl(axis_accumulate_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, stream, idx, \
                                (const T*)val, n, n_dev, rk, acc, touched));                                  \
    }                                                                                                         \
    break
    switch (dtype) {
        case SPB_F32: SPB_ACC(float);
        case SPB_F64: SPB_ACC(double);
        case SPB_I32: SPB_ACC(int32_t);
        case SPB_I64: SPB_ACC(int64_t);
        case SPB_U8: case SPB_BOOL: SPB_ACC(uint8_t);
        case SPB_I8: SPB_ACC(int8_t);
        case SPB_I16: SPB_ACC(int16_t);
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_ACC
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

// cells of [0, cells) relative to acc / touched (touched readable up to the next multiple of 64)
static int launch_select(double* acc, uint8_t* touched, int64_t cells, int64_t key_base, int64_t* out_idx,
                         double* out_sums, int64_t* out_count, spb_workspace* ws, cudaStream_t stream) {
    const int64_t ntiles = ceil_div(cells, 256 * 64);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)ntiles * 8, stream);
    if (rc) return rc;
    uint32_t epoch;
    rc = workspace_next_epoch(ws, stream, &epoch);
    if (rc) return rc;
    SPB_CUDA_TRY(launch_pdl(select_touched_kernel, dim3((unsigned)ntiles), dim3(256), 0, stream, acc, touched, cells,
                            key_base, out_idx, out_sums, out_count, ws_state(ws), epoch));
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

}  // namespace spb

extern "C" int spb_axis_sum_dense(int dtype, const int64_t* idx, const void* val, int64_t n, const int64_t* n_dev, int ndim,
                                  const int64_t* in_strides, const int64_t* out_strides, int64_t cells,
                                  int64_t* out_idx, double* out_sums, int64_t* out_count,
                                  spb_workspace* ws, spb_stream_t stream) {
    using namespace spb;
    if (n < 0 || cells < 0 || !out_count) return SPB_ERR_BAD_ARG;
    Rekey rk{};
    int rc = make_rekey(rk, ndim, in_strides, out_strides);
    if (rc) return rc;
    if (n == 0 || cells == 0) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!idx || !val || !out_idx || !out_sums || !ws) return SPB_ERR_BAD_ARG;
    // accumulators (cells doubles) then one flag byte per cell (padded to 64) in the workspace's self-cleaning slab
    const size_t acc_bytes = ((size_t)cells * sizeof(double) + 255) & ~(size_t)255;
    const int64_t nflags = ceil_div(cells, 64) * 64;
    rc = workspace_reserve_accum(ws, acc_bytes + (size_t)nflags, stream);
    if (rc) return rc;
    double* const acc = reinterpret_cast<double*>(ws->accum);
    uint8_t* const touched = reinterpret_cast<uint8_t*>(ws->accum) + acc_bytes;
    rc = launch_accumulate(dtype, idx, val, n, n_dev, rk, cells, acc, touched, stream);
    if (rc) return rc;
    return launch_select(acc, touched, cells, 0, out_idx, out_sums, out_count, ws, stream);
}

// The two halves on caller-owned buffers, for the sharded axis sum (all-reduce of the accumulators
// between them): acc = cells doubles, touched = cells flag bytes padded to a multiple of 64, both
// zeroed by the caller, touched 16-byte aligned.
extern "C" int spb_axis_accumulate(int dtype, const int64_t* idx, const void* val, int64_t n, int ndim,
                                   const int64_t* in_strides, const int64_t* out_strides, int64_t cells, double* acc,
                                   uint8_t* touched, spb_stream_t stream) {
    using namespace spb;
    if (n < 0 || cells < 0) return SPB_ERR_BAD_ARG;
    Rekey rk{};
    const int rc = make_rekey(rk, ndim, in_strides, out_strides);
    if (rc) return rc;
    if (n == 0 || cells == 0) return SPB_OK;
    if (!idx || !val || !acc || !touched) return SPB_ERR_BAD_ARG;
    return launch_accumulate(dtype, idx, val, n, nullptr, rk, cells, acc, touched, stream);
}

// Emit the touched cells of [cell_lo, cell_hi) in key order (keys are absolute cell numbers).
// cell_lo must be a multiple of 64 (flag words are read 16 bytes at a time); the cells that are
// emitted are zeroed.  out_idx / out_sums: cell_hi - cell_lo slots.
extern "C" int spb_emit_touched(double* acc, uint8_t* touched, int64_t cell_lo, int64_t cell_hi, int64_t* out_idx,
                                double* out_sums, int64_t* out_count, spb_workspace* ws, spb_stream_t stream) {
    using namespace spb;
    if (cell_lo < 0 || cell_hi < cell_lo || (cell_lo & 63) != 0 || !out_count) return SPB_ERR_BAD_ARG;
    if (cell_hi == cell_lo) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!acc || !touched || !out_idx || !out_sums || !ws) return SPB_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(touched) & 15) != 0) return SPB_ERR_BAD_ARG;
    return launch_select(acc + cell_lo, touched + cell_lo, cell_hi - cell_lo, cell_lo, out_idx, out_sums, out_count, ws,
                         stream);
}
