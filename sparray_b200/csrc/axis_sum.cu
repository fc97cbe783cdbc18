// sparray_b200/csrc/axis_sum.cu -- sum(axis) into dense float64 accumulators (no sort).
//
// Replaces the unravel / re-ravel + np.unique + np.bincount of FlatSparray.sum (flat.py:617-625)
// whenever the reduced array is small enough to hold densely.
#include "rekey.cuh"

#include <string.h>

// ---- axis sums into a dense accumulator (no sort) ---------------------------------------------------
// When the reduced array prod(new_shape) is small enough to hold densely, sum(axis != last)
// (flat.py:617-625) does not need the re-key + radix sort at all: every stored element adds its
// value into acc[new_key] with a float64 atomic (the accumulator is L2-resident for the BASELINE
// configs), and one compaction pass over the cells emits the touched ones in key order.
//
// "Touched" cannot be "sum != 0": explicit zeros are stored elements and must survive (and so must sums
// that cancel).  Two ways to tell:
//   * flags   -- one byte per cell, set next to every atomic.  The emitting pass then scans 1 byte per cell:
//     the right thing when few cells are touched (a lazily counted product: 10^5 elements into 4 M cells).
//   * sentinel -- no flags at all.  Between calls every accumulator holds -0.0.  IEEE 754 round-to-nearest:
//     x + (-0.0) == x for every x but -0.0 itself, and sums that cancel give +0.0, never -0.0; stored -0.0
//     values are turned into +0.0 on the way in (np.bincount starts from +0.0, so it also yields +0.0 for
//     them).  A cell is untouched iff it still holds the bit pattern of -0.0.  This saves one scattered store
//     per element: with a middle or leading axis reduced, a warp's 32 elements hit 32 different sectors, and
//     the flag stores were 45 % of the accumulate pass (190 -> 125 us for a config-2 axis-0 sum).
namespace spb {

// A warp reads 32 consecutive elements (coalesced), re-keys them, combines runs of equal new keys with a
// segmented shuffle scan (when the reduced axis is the last one the elements of an output cell are
// contiguous; otherwise runs are rare and the scan is a few wasted instructions) and the last lane of
// each run adds the run's sum into its cell with one float64 atomic.
constexpr int kAccUnroll = 4;

// cell of a stored element inside the accumulator window [cell_lo, cell_lo + cells): -1 when it falls
// outside (a flat index beyond prod(shape), which resize() permits, must not touch the slab)
__device__ __forceinline__ int64_t window_cell(int64_t idx, const Rekey& rk, int64_t cell_lo, int64_t cells) {
    const int64_t c = apply_rekey(idx, rk) - cell_lo;
    return (uint64_t)c < (uint64_t)cells ? c : -1;
}

// RUNS: the finest axis group is reduced, so neighbouring elements can fall into the same cell; otherwise they
// never do (distinct flat indices that agree on every kept coordinate differ in a coarser, reduced one and lie
// far apart) and the run detection -- two shuffles and a vote per element on a kernel that is bound by its
// memory-instruction queue -- is left out
template <typename T, bool RUNS>
__global__ void __launch_bounds__(256) axis_accumulate_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                              int64_t n, const int64_t* __restrict__ n_dev,
                                                              const Rekey rk, double* __restrict__ acc,
                                                              uint8_t* __restrict__ touched, const int64_t cell_lo,
                                                              const int64_t cells, uint64_t* __restrict__ clear_state,
                                                              const int64_t clear_n) {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();   // operands (and n_dev) may come from the previous launch
    // the group counters of the emitting launches that follow start from zero (they wait for this whole grid):
    // no memset node in front of them when the chain is captured into a CUDA graph
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < clear_n; i += (int64_t)gridDim.x * blockDim.x)
        clear_state[i] = 0;
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
            const int64_t key = window_cell(__ldcs(idx + gtid), rk, cell_lo, cells);
            if (key >= 0) {
                atomicAdd(acc + key, (double)__ldcs(val + gtid) + 0.0);      // (-0.0 goes in as +0.0)
                if (touched) touched[key] = 1;
            }
        }
        return;
    }
    const int64_t warp_global = gtid >> 5;
    const int64_t nwarps = nthreads >> 5;
    const int64_t chunk = 32 * kAccUnroll;
    auto load_chunk = [&](int64_t base, int64_t (&k)[kAccUnroll], T (&t)[kAccUnroll]) {
#pragma unroll
        for (int u = 0; u < kAccUnroll; ++u) {
            const int64_t i = base + u * 32 + lane;
            k[u] = i < n ? __ldcs(idx + i) : -1;
            t[u] = i < n ? __ldcs(val + i) : T(0);
        }
    };
    // the next chunk's loads are issued before the current chunk's atomics
#ifndef SPB_ACC_PIPE2
#define SPB_ACC_PIPE2 1
#endif
    int64_t key[kAccUnroll], nkey[kAccUnroll];
    T tv[kAccUnroll], ntv[kAccUnroll];
    int64_t base = warp_global * chunk;
    if (SPB_ACC_PIPE2 && base < n) load_chunk(base, nkey, ntv);
    for (; base < n; base += nwarps * chunk) {
        if (SPB_ACC_PIPE2) {
#pragma unroll
            for (int u = 0; u < kAccUnroll; ++u) {
                key[u] = nkey[u];
                tv[u] = ntv[u];
            }
            if (base + nwarps * chunk < n) load_chunk(base + nwarps * chunk, nkey, ntv);
        } else {
            load_chunk(base, key, tv);
        }
        double v[kAccUnroll];
#pragma unroll
        for (int u = 0; u < kAccUnroll; ++u) v[u] = (double)tv[u] + 0.0;     // (-0.0 goes in as +0.0)
#pragma unroll
        for (int u = 0; u < kAccUnroll; ++u)
            if (key[u] >= 0) key[u] = window_cell(key[u], rk, cell_lo, cells);
        if constexpr (!RUNS) {
#pragma unroll
            for (int u = 0; u < kAccUnroll; ++u) {
                if (key[u] >= 0) {
                    atomicAdd(acc + key[u], v[u]);
                    if (touched) touched[key[u]] = 1;  // (sentinel mode: no flags)
                }
            }
            continue;
        }
#pragma unroll
        for (int u = 0; u < kAccUnroll; ++u) {
            // runs = maximal stretches of ADJACENT lanes with the same key (equal keys further apart are
            // separate runs and simply cost one more atomic)
            const int64_t kprev = __shfl_up_sync(0xffffffffu, key[u], 1);
            const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || kprev != key[u]);
            const int start = 31 - __clz(heads & ((2u << lane) - 1u));       // first lane of my run
            double s = v[u];
            if (heads != 0xffffffffu) {                                       // some run is longer than one lane
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double up = __shfl_up_sync(0xffffffffu, s, o);
                    if (lane - o >= start) s += up;
                }
            }
            const bool last = lane == 31 || ((heads >> (lane + 1)) & 1u);
            if (key[u] >= 0 && last) {
                atomicAdd(acc + key[u], s);
                if (touched) touched[key[u]] = 1;  // a flag per cell: explicit zeros survive
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
                                                              uint8_t* __restrict__ touched, const int64_t cell_lo,
                                                              const int64_t cells, uint64_t* __restrict__ clear_state,
                                                              const int64_t clear_n) {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();   // operands (and n_dev) may come from the previous launch
    // the group counters of the emitting launches that follow start from zero (they wait for this whole grid):
    // no memset node in front of them when the chain is captured into a CUDA graph
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < clear_n; i += (int64_t)gridDim.x * blockDim.x)
        clear_state[i] = 0;
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
            v[j] = i0 + j < n ? (double)__ldcs(val + i0 + j) + 0.0 : 0.0;    // (-0.0 goes in as +0.0)
        }
#pragma unroll
        for (int j = 0; j < kAccItems; ++j)
            if (key[j] >= 0) key[j] = window_cell(key[j], rk, cell_lo, cells);
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
                            if (touched) touched[run] = 1;
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
                if (touched) touched[run] = 1;       // a flag per cell: explicit zeros survive
            }
        }
        bool last;
        s = warp_run_sum(cell, s, lane, last);
        if (cell >= 0 && last) {
            atomicAdd(acc + cell, s);
            if (touched) touched[cell] = 1;
        }
    }
}

// The same for the plain case -- the LAST axis group is summed and every other group keeps its place, so the
// cell of an element is its flat index divided by the size of that group and nothing else: a lane's four
// ascending elements need TWO divisions (first and last element; equal quotients = one run, the usual case),
// a lane that straddles a group boundary places its middle elements by comparing them with the two group
// bounds.  Full chunks are read with 128-bit loads.  (The general kernel above re-keys every element: 116
// thread instructions per element, 80 % issue activity on the config-4 row sums -- it was bound by that.)
template <typename T, int BYTES = (int)sizeof(T)>
struct Load4;
template <typename T>
struct Load4<T, 8> {
    static __device__ __forceinline__ void run(const T* p, T (&o)[4]) {
        const ulonglong2 a = __ldcs(reinterpret_cast<const ulonglong2*>(p));
        const ulonglong2 b = __ldcs(reinterpret_cast<const ulonglong2*>(p) + 1);
        const unsigned long long w[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) memcpy(&o[j], &w[j], 8);
    }
};
template <typename T>
struct Load4<T, 4> {
    static __device__ __forceinline__ void run(const T* p, T (&o)[4]) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p));
        const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) memcpy(&o[j], &w[j], 4);
    }
};
template <typename T>
struct Load4<T, 2> {
    static __device__ __forceinline__ void run(const T* p, T (&o)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __ldcs(p + j);
    }
};
template <typename T>
struct Load4<T, 1> {
    static __device__ __forceinline__ void run(const T* p, T (&o)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __ldcs(p + j);
    }
};

template <typename T>
__global__ void __launch_bounds__(256) axis_accumulate_lastaxis_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val,
                                                                   int64_t n, const int64_t* __restrict__ n_dev,
                                                                   const FastDiv dv, double* __restrict__ acc,
                                                                   uint8_t* __restrict__ touched, const int64_t cell_lo,
                                                                   const int64_t cells, uint64_t* __restrict__ clear_state,
                                                                   const int64_t clear_n, const int aligned) {
    static_assert(kAccItems == 4, "written out for 4 elements per lane");
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();   // operands (and n_dev) may come from the previous launch
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < clear_n; i += (int64_t)gridDim.x * blockDim.x)
        clear_state[i] = 0;
    if (n_dev) {                       // the producer's count never left the device; n is its upper bound
        const int64_t m = *n_dev;
        n = m < n ? m : n;
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t chunk = 32 * kAccItems;
    auto add_run = [&](uint64_t g, double rs) {      // one run's sum into its cell, if the cell is in the window
        const uint64_t c = g - (uint64_t)cell_lo;
        if (c < (uint64_t)cells) {
            atomicAdd(acc + c, rs);
            if (touched) touched[c] = 1;             // a flag per cell: explicit zeros survive
        }
    };
    auto load_chunk = [&](int64_t base, int64_t (&key)[kAccItems], T (&tv)[kAccItems]) {
        const int64_t i0 = base + (int64_t)lane * kAccItems;
        if (aligned && base + chunk <= n) {          // (uniform) a full chunk: two 128-bit loads of indices
            Load4<int64_t>::run(idx + i0, key);
            Load4<T>::run(val + i0, tv);
        } else {
#pragma unroll
            for (int j = 0; j < kAccItems; ++j) {
                key[j] = i0 + j < n ? __ldcs(idx + i0 + j) : -1;
                tv[j] = i0 + j < n ? __ldcs(val + i0 + j) : T(0);
            }
        }
    };
    // the next chunk's loads are issued before the current chunk is worked on
    int64_t key[kAccItems], nkey[kAccItems];
    T tv[kAccItems], ntv[kAccItems];
#ifndef SPB_ACC_PIPE
#define SPB_ACC_PIPE 1
#endif
    int64_t base = warp_global * chunk;
    if (SPB_ACC_PIPE && base < n) load_chunk(base, key, tv);
    for (; base < n; base += nwarps * chunk) {
        const int64_t nbase = base + nwarps * chunk;
        if (SPB_ACC_PIPE) {
            if (nbase < n) load_chunk(nbase, nkey, ntv);
        } else {
            load_chunk(base, key, tv);
        }
        double v[kAccItems];
#pragma unroll
        for (int j = 0; j < kAccItems; ++j) v[j] = (double)tv[j] + 0.0;       // (-0.0 goes in as +0.0)
        int64_t cell = -2 - lane;      // negative and distinct: takes no part in the warp's runs
        double s = 0.0;
        if (key[0] >= 0) {
            // elements ascend and the missing ones (-1, past the end) are the last: the last real one
            int jl = 3;
#pragma unroll
            for (int j = 3; j > 0; --j)
                if (key[j] < 0) jl = j - 1;
            const uint64_t g0 = fastdiv((uint64_t)key[0], dv);
            const int64_t kl = key[3] >= 0 ? key[3] : (key[2] >= 0 ? key[2] : (key[1] >= 0 ? key[1] : key[0]));
            const uint64_t gl = jl == 0 ? g0 : fastdiv((uint64_t)kl, dv);
            if (g0 == gl) {            // one run (all four elements, or the real ones of the last chunk)
                const uint64_t c = g0 - (uint64_t)cell_lo;
                if (c < (uint64_t)cells) {
                    cell = (int64_t)c;
                    s = (v[0] + v[1]) + (v[2] + v[3]);
                }
            } else {
                // several runs: each goes in atomically right here.  A middle element lies in the first group
                // (below its end), in the last one (at or above its start), or -- a lane spanning three groups
                // or more, rare -- somewhere in between
                const uint64_t end0 = (g0 + 1) * dv.d, startl = gl * dv.d;
                uint64_t run = g0;
                double rs = v[0];
#pragma unroll
                for (int j = 1; j < kAccItems; ++j) {
                    if (j <= jl) {
                        const uint64_t k = (uint64_t)key[j];
                        const uint64_t g = k < end0 ? g0 : (k >= startl ? gl : fastdiv(k, dv));
                        if (g != run) {
                            add_run(run, rs);
                            run = g;
                            rs = v[j];
                        } else {
                            rs += v[j];
                        }
                    }
                }
                add_run(run, rs);
            }
        }
        bool last;
        s = warp_run_sum(cell, s, lane, last);
        if (cell >= 0 && last) {
            atomicAdd(acc + cell, s);
            if (touched) touched[cell] = 1;
        }
        if (SPB_ACC_PIPE) {
#pragma unroll
            for (int j = 0; j < kAccItems; ++j) {
                key[j] = nkey[j];
                tv[j] = ntv[j];
            }
        }
    }
}

// Emits the touched cells in key order and puts every cell it emitted back to zero, so the
// accumulator slab is all-zero again when the call is over (no memset per call).  One thread owns
// kSelCells cells (one 16-byte load of flag bytes); a tile of 256 threads covers 4 K cells -- thousands of
// tiles for the BASELINE shapes, enough to fill the machine.
// Two launches: count_touched_kernel reads the flags only and leaves one count per tile plus one per group of
// 64 tiles; select_touched_kernel then works out its output offset on its own (a few dozen L2 loads) -- no CTA
// waits for another.  (With a decoupled look-back in a single pass the tiles ran in lockstep waves, each one
// waiting for the slowest predecessor's count: 160 us for the 16.7 M cells of a config-2 axis-0 sum, against
// the ~60 us its traffic takes.)
constexpr int kSelCells = 16;
constexpr int kSelGroupShift = 6;          // tiles per group counter: 64

struct SelCounts {
    unsigned long long* group_cnt;         // [ceil(ntiles / 64)], zero before count_touched_kernel runs
    int* counts;                           // [ntiles]
};

__device__ __forceinline__ unsigned touched_bits(const uint8_t* touched, int64_t first, int64_t cells) {
    unsigned bits = 0;
    if (first < cells) {               // the flag array is padded to a multiple of 64 (the padding stays zero)
        const uint4 raw = *reinterpret_cast<const uint4*>(touched + first);
        const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) bits |= (((w[k] & 0x01010101u) * 0x01020408u) >> 24) << (4 * k);
    }
    return bits;
}

// sentinel mode: a cell is touched iff it no longer holds the bit pattern of -0.0 (acc is readable, and clean,
// up to the next multiple of 16 cells)
__device__ __forceinline__ unsigned sentinel_bits(const double* acc, int64_t first, int64_t cells) {
    unsigned bits = 0;
    if (first < cells) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(acc + first);
#pragma unroll
        for (int k = 0; k < kSelCells / 2; ++k) {
            const ulonglong2 w = p[k];
            bits |= (unsigned)(w.x != kNegZeroBits) << (2 * k);
            bits |= (unsigned)(w.y != kNegZeroBits) << (2 * k + 1);
        }
    }
    return bits;
}

__global__ void __launch_bounds__(256) count_touched_kernel(const double* __restrict__ acc,
                                                            const uint8_t* __restrict__ touched, int64_t cells,
                                                            const SelCounts sc) {
    __shared__ int s_w[8];
    const int64_t tile = blockIdx.x;
    const int64_t first = (tile * 256 + threadIdx.x) * kSelCells;
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();   // flags (and the zeroed group counters) come from the previous launch
    int cnt = __popc(touched ? touched_bits(touched, first, cells) : sentinel_bits(acc, first, cells));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) total += s_w[w];
        sc.counts[tile] = total;
        if (total) atomicAdd(sc.group_cnt + (tile >> kSelGroupShift), (unsigned long long)total);
    }
}

__global__ void __launch_bounds__(256) select_touched_kernel(double* __restrict__ acc, uint8_t* __restrict__ touched,
                                                             int64_t cells, int64_t key_base, int64_t* __restrict__ out_idx,
                                                             double* __restrict__ out_val, int64_t* out_count,
                                                             const SelCounts sc, const double clean) {
    __shared__ int64_t s_red[8];
    __shared__ int s_w[8];
    const int64_t tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t first = (tile * 256 + tid) * kSelCells;
    cudaGridDependencySynchronize();   // accumulators and counts come from the previous launches
    unsigned bits;
    if (touched) {
        bits = touched_bits(touched, first, cells);
        if (bits) *reinterpret_cast<uint4*>(touched + first) = make_uint4(0, 0, 0, 0);
    } else {
        bits = sentinel_bits(acc, first, cells);
    }
    const int cnt = __popc(bits);
    // output offset of the tile: the groups in front of its own + the earlier tiles of its own group
    int64_t gsum = 0;
    {
        const int64_t g = tile >> kSelGroupShift;
        for (int64_t j = tid; j < g; j += 256) gsum += (int64_t)__ldg(sc.group_cnt + j);
        for (int64_t j = (g << kSelGroupShift) + tid; j < tile; j += 256) gsum += __ldg(sc.counts + j);
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
    if (lane == 31) s_w[warp] = incl;
    if (lane == 0) s_red[warp] = gsum;
    __syncthreads();
    int64_t tile_base = 0;
    int wofs = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        tile_base += s_red[w];
        if (w < warp) wofs += s_w[w];
        tile_total += s_w[w];
    }
    if (tile == gridDim.x - 1 && tid == 0) *out_count = tile_base + tile_total;
    const int64_t off = tile_base + wofs + incl - cnt;
    // Warp-cooperative emission: the warp's outputs are one contiguous run, written 32 at a time (coalesced
    // whether the cells are nearly all touched or nearly all empty).  Output r of the warp belongs to the
    // first lane whose inclusive count exceeds r, and is that lane's (r - exclusive count)-th set bit.
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    const int64_t warp_off = __shfl_sync(0xffffffffu, off, 0);
    const int64_t warp_first = first - (int64_t)lane * kSelCells;
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
        const unsigned obits = __shfl_sync(0xffffffffu, bits, owner);
        if (r < warp_total) {
            const int k = (int)__fns(obits, 0, r - owner_excl + 1);
            const int local = owner * kSelCells + k;
            out_idx[warp_off + r] = key_base + warp_first + local;
            // one gather per emitted element (measured against loading every cell of a touched thread up
            // front and emitting from shared memory: 111 vs 126 us on the config-2 operand)
            out_val[warp_off + r] = acc[warp_first + local];
            acc[warp_first + local] = clean;       // -0.0 (the workspace's slab) or +0.0 (caller-owned buffers)
        }
    }
}

}  // namespace spb

namespace spb {

static int launch_accumulate(int dtype, const int64_t* idx, const void* val, int64_t n, const int64_t* n_dev,
                             const Rekey& rk, int64_t cell_lo, int64_t cells, double* acc, uint8_t* touched,
                             uint64_t* clear_state, int64_t clear_n, cudaStream_t stream) {
    int64_t blocks = ceil_div(n, 256 * kAccUnroll);
    if (blocks > (int64_t)device_sm_count() * 16) blocks = (int64_t)device_sm_count() * 16;
    // long contiguous runs (innermost axis reduced, >= 8 elements per output cell on average): the variant
    // that pre-reduces inside a lane and scans once per 128 elements
    const bool runs = rk.ndim >= 1 && rk.out_mul[rk.ndim - 1] == 0;      // the finest axis group is reduced
    const bool long_runs = runs && n >= 8 * cells;
    // plain last-axis sum: every kept group keeps its place, so cell == flat index / size of the last group
    bool plain = long_runs && rk.ndim >= 2;
    if (plain) {
        const uint64_t inner = rk.in_div[rk.ndim - 2].d;
        for (int ax = 0; ax < rk.ndim - 1; ++ax)
            if ((uint64_t)rk.out_mul[ax] * inner != rk.in_div[ax].d) plain = false;
    }
    const int aligned = ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(val)) & 15) == 0;
#define SPB_ACC(T)                                                                                            \
    if (plain) {                                                                                              \
        SPB_CUDA_TRY(launch_pdl(axis_accumulate_lastaxis_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, stream, \
                                idx, (const T*)val, n, n_dev, rk.in_div[rk.ndim - 2], acc, touched, cell_lo, cells, \
                                clear_state, clear_n, aligned));                                              \
    } else if (long_runs) {                                                                                          \
        SPB_CUDA_TRY(launch_pdl(axis_accumulate_runs_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, stream, \
                                idx, (const T*)val, n, n_dev, rk, acc, touched, cell_lo, cells, clear_state, clear_n)); \
    } else if (runs) {                                                                                        \
        SPB_CUDA_TRY(launch_pdl(axis_accumulate_kernel<T, true>, dim3((unsigned)blocks), dim3(256), 0, stream, idx, \
                                (const T*)val, n, n_dev, rk, acc, touched, cell_lo, cells, clear_state, clear_n)); \
    } else {                                                                                                  \
        SPB_CUDA_TRY(launch_pdl(axis_accumulate_kernel<T, false>, dim3((unsigned)blocks), dim3(256), 0, stream, idx, \
                                (const T*)val, n, n_dev, rk, acc, touched, cell_lo, cells, clear_state, clear_n)); \
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

static double __longlong_as_double_host(unsigned long long b) {
    double d;
    memcpy(&d, &b, sizeof(d));
    return d;
}

// cells of [0, cells) relative to acc / touched (touched readable up to the next multiple of 64)
static int64_t select_tiles(int64_t cells) { return ceil_div(cells, 256 * kSelCells); }

// Scratch of the two emitting launches, in the workspace's state region: [group counters u64][tile counts i32]
// (small integers there can never pass for a look-back word of a live epoch: their status bits are zero).
static int64_t select_groups(int64_t ntiles) { return ceil_div(ntiles, (int64_t)1 << kSelGroupShift); }
static size_t select_scratch_bytes(int64_t ntiles) { return (size_t)select_groups(ntiles) * 8 + (size_t)ntiles * 4; }

// precleared: the producer launch zeroed the group counters (see the accumulate kernels)
// touched == nullptr: sentinel mode.  clean: what an emitted cell is put back to.
static int launch_select(double* acc, uint8_t* touched, int64_t cells, int64_t key_base, int64_t* out_idx,
                         double* out_sums, int64_t* out_count, spb_workspace* ws, cudaStream_t stream,
                         double clean, bool precleared = false) {
    const int64_t ntiles = select_tiles(cells);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, select_scratch_bytes(ntiles), stream);
    if (rc) return rc;
    SelCounts sc;
    sc.group_cnt = reinterpret_cast<unsigned long long*>(ws_state(ws));
    sc.counts = reinterpret_cast<int*>(sc.group_cnt + select_groups(ntiles));
    if (!precleared) SPB_CUDA_TRY(cudaMemsetAsync(sc.group_cnt, 0, (size_t)select_groups(ntiles) * 8, stream));
    SPB_CUDA_TRY(launch_pdl(count_touched_kernel, dim3((unsigned)ntiles), dim3(256), 0, stream, (const double*)acc,
                            (const uint8_t*)touched, cells, sc));
    SPB_CUDA_TRY(launch_pdl(select_touched_kernel, dim3((unsigned)ntiles), dim3(256), 0, stream, acc, touched, cells,
                            key_base, out_idx, out_sums, out_count, sc, clean));
    SPB_LAUNCHED(2);
    return (int)cudaGetLastError();
}

}  // namespace spb

// cells of the reduced array in [cell_lo, cell_hi) only (every element must fall inside; others are ignored)
static int axis_sum_window(int dtype, const int64_t* idx, const void* val, int64_t n, const int64_t* n_dev, int ndim,
                           const int64_t* in_strides, const int64_t* out_strides, int64_t cell_lo, int64_t cell_hi,
                           int64_t* out_idx, double* out_sums, int64_t* out_count, spb_workspace* ws,
                           cudaStream_t stream) {
    using namespace spb;
    if (n < 0 || cell_lo < 0 || cell_hi < cell_lo || !out_count) return SPB_ERR_BAD_ARG;
    const int64_t cells = cell_hi - cell_lo;
    Rekey rk{};
    if (ndim < 1) return SPB_ERR_BAD_ARG;
    int rc = make_rekey(rk, ndim, in_strides, out_strides);
    if (rc) return rc;
    if (n == 0 || cells == 0) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!idx || !val || !out_idx || !out_sums || !ws) return SPB_ERR_BAD_ARG;
    // Flags or sentinel?  Flags when the count is only an upper bound (a lazily counted producer: usually far
    // fewer elements than that) or when few cells can be touched: scanning one byte per cell is what the
    // emitting pass costs then.  Otherwise no flags: one scattered store less per element ...
    // ... when the flag stores would scatter, that is: the finest axis group is kept and long enough that a
    // warp's 32 consecutive elements land in 32 different sectors of the flag array (config 2, axis 0 or 1:
    // runs of 2^22 / 2^14 cells).  With a short kept group behind the reduced axis (axis 2: 64 cells) the
    // flag bytes of a warp share a sector or two and cost next to nothing, while the sentinel variant's
    // emitting pass reads 8 bytes per cell instead of 1 (measured: 109 us with flags, 120 us without).
    const bool scattered = ndim >= 2 && out_strides[ndim - 1] != 0 && in_strides[ndim - 2] >= 512;
    const bool use_flags = n_dev != nullptr || n < cells / 4 || !scattered;
    // accumulators (cells doubles, readable up to a multiple of 16 cells) and one flag byte per cell (padded to
    // 64) in the workspace's self-cleaning buffers
    const size_t acc_bytes = ((size_t)cells * sizeof(double) + 255) & ~(size_t)255;
    const int64_t nflags = ceil_div(cells, 64) * 64;
    rc = workspace_reserve_accum(ws, acc_bytes, use_flags ? (size_t)nflags : 0, stream);
    if (rc) return rc;
    // The buffers are only clean again once the emitting pass has run: if an earlier call failed between
    // its launches, wipe them before anybody adds into them.
    if (ws->accum_dirty) {
        rc = workspace_wipe_accum(ws, stream);
        if (rc) return rc;
    }
    double* const acc = reinterpret_cast<double*>(ws->accum);
    uint8_t* const touched = use_flags ? reinterpret_cast<uint8_t*>(ws->flags) : nullptr;
    ws->accum_dirty = 1;
    const int64_t ntiles = select_tiles(cells);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    rc = workspace_reserve(ws, select_scratch_bytes(ntiles), stream);    // the emitting launches' counters: the group
    if (rc) return rc;                                                   // counters are zeroed by the accumulate launch
    rc = launch_accumulate(dtype, idx, val, n, n_dev, rk, cell_lo, cells, acc, touched, ws_state(ws),
                           select_groups(ntiles), stream);
    if (rc) return rc;
    rc = launch_select(acc, touched, cells, cell_lo, out_idx, out_sums, out_count, ws, stream,
                       __longlong_as_double_host(kNegZeroBits), true);
    if (rc == SPB_OK) ws->accum_dirty = 0;
    return rc;
}

extern "C" int spb_axis_sum_dense(int dtype, const int64_t* idx, const void* val, int64_t n, const int64_t* n_dev, int ndim,
                                  const int64_t* in_strides, const int64_t* out_strides, int64_t cells,
                                  int64_t* out_idx, double* out_sums, int64_t* out_count,
                                  spb_workspace* ws, spb_stream_t stream) {
    return axis_sum_window(dtype, idx, val, n, n_dev, ndim, in_strides, out_strides, 0, cells, out_idx, out_sums,
                           out_count, ws, stream);
}

extern "C" int spb_axis_sum_window(int dtype, const int64_t* idx, const void* val, int64_t n, const int64_t* n_dev,
                                   int ndim, const int64_t* in_strides, const int64_t* out_strides, int64_t cell_lo,
                                   int64_t cell_hi, int64_t* out_idx, double* out_sums, int64_t* out_count,
                                   spb_workspace* ws, spb_stream_t stream) {
    return axis_sum_window(dtype, idx, val, n, n_dev, ndim, in_strides, out_strides, cell_lo, cell_hi, out_idx,
                           out_sums, out_count, ws, stream);
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
    if (ndim < 1) return SPB_ERR_BAD_ARG;
    const int rc = make_rekey(rk, ndim, in_strides, out_strides);
    if (rc) return rc;
    if (n == 0 || cells == 0) return SPB_OK;
    if (!idx || !val || !acc || !touched) return SPB_ERR_BAD_ARG;
    return launch_accumulate(dtype, idx, val, n, nullptr, rk, 0, cells, acc, touched, nullptr, 0, stream);
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
                         stream, 0.0);
}
