// sparray_b200/csrc/slab_sum.cu -- sum(axis) by streaming over the sorted flat indices: no sort, no dense
// array of prod(new_shape) accumulators, no pass over every cell of the reduced array.
//
// Replaces the unravel / re-ravel + np.unique + np.bincount of FlatSparray.sum (flat.py:617-625) whenever
// the stored elements that feed a bounded set of output cells are contiguous in the input:
//
//   Reducing axis k of shape (d0, .., dn-1) splits the C-order index space into SLABS of
//   slab_cells = dk * suffix consecutive cells (suffix = d(k+1) * .. * d(n-1)), one per value of the
//   leading coordinates.  All elements of a slab are contiguous in the sorted input, they feed exactly
//   `suffix` output cells, and slab s owns the output keys [s * suffix, (s + 1) * suffix) -- so TILES made of
//   whole slabs can be reduced independently and their outputs concatenate in key order.
//
//   slab_partition_kernel  : one warp per tile boundary (32-ary search for the first element of a slab);
//   *_kernel<COUNT = true> : keys only -- how many output cells does each tile touch?  The count also goes into
//                            one counter per group of tiles, so that
//   *_kernel<COUNT = false>: can work out its output offset on its own (a few hundred L2 loads per CTA: no scan
//                            launch, and no CTA ever waits for another), reduces the tile in shared memory and
//                            stores the (key, sum) pairs coalesced.
//
//   Reduced axis LAST (slab_last_kernel): the elements of an output cell are adjacent.  A thread reads 8
//   consecutive elements with 128-bit loads and adds up whole runs in registers; the output slot of a run is
//   the number of run heads in front of it (block scan), so the tile's results are compact from the start.  A
//   run inside a thread's stretch is that thread's alone and is simply stored; only a thread's first and last
//   run (which the neighbours may share) are added with a shared-memory float64 atomic.
//   Other axes (slab_cells_kernel): one shared-memory float64 add per element into accumulators indexed by the
//   tile-local cell (at most 4096 cells per tile), a flag byte per cell (explicit zeros survive), then the
//   touched cells are listed in key order.
//
// Why two passes over the keys instead of one pass with a decoupled look-back: measured.  With the look-back,
// every wave of ~740 resident tiles started at the same moment, loaded in one burst (6 us) and then sat waiting
// for the slowest predecessor's count (6 us more) -- 14 us of tile lifetime for 24 KB of payload, 118 us for the
// config-2 operand whatever the tile did (profiles/r2_slab_phases_lookback.txt).  Without the chain the tiles
// drift apart and the loads stream.  Accumulating with float64 reductions into CTA-private strips of GLOBAL
// memory was 3x slower still: the strips did not stay in L2 (275 MB of DRAM writes for 62 MB of output) and
// every tile paid several L2 round trips in sequence.
//
// Algorithmic traffic n(8 + v) + nout * 16; DRAM traffic 8n more (the keys are read twice).
#include "common.cuh"

namespace spb {

constexpr int kSlabNT = 256;
constexpr int kSlabItems = 8;                  // consecutive elements per thread
constexpr int kSlabBatch = kSlabNT * kSlabItems;
constexpr int kSlabCapCells = 4096;            // output cells a tile may span: 32 KB of float64 accumulators
constexpr int kSlabTileElems = 1792;           // stored elements per tile, on average: nearly every tile fits ONE batch
constexpr int kSlabCtasPerSm = 5;              // ~45 KB of shared memory each
constexpr int64_t kSlabMaxGroups = 1024;       // group counters a CTA sums to find its output offset

struct SlabParams {
    const int64_t* idx;
    const void* val;
    int64_t n;
    FastDiv slab_div;          // key / slab_div.d = slab number
    FastDiv suf_div;           // suffix: output cells per slab
    int64_t total_slabs;
    int64_t slabs_per_tile;
    int64_t ntiles;
    int64_t* bounds;           // [ntiles + 1] first element of every tile
    int* counts;               // [ntiles] output cells per tile (pass 1)
    unsigned long long* group_cnt;   // [ngroups] the same, summed over groups of 2^group_shift tiles
    int group_shift;
    int64_t ngroups;
    int64_t* out_idx;
    double* out_sums;
    int64_t* out_count;
};

// bounds[t] = first element whose slab is >= t * slabs_per_tile (elements beyond the last slab -- flat
// indices past prod(shape), which resize() permits -- fall behind bounds[ntiles] and are ignored)
__global__ void __launch_bounds__(256) slab_partition_kernel(const SlabParams p) {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();               // the operand may come from the previous launch
    const int lane = threadIdx.x & 31;
    const int64_t gtid = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (gtid < p.ngroups) p.group_cnt[gtid] = 0;
    const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t > p.ntiles) return;
    int64_t target = t * p.slabs_per_tile;
    if (target > p.total_slabs) target = p.total_slabs;
    int64_t lo = 0, hi = p.n;
    while (lo < hi) {
        const int64_t span = hi - lo;
        const int64_t m = lo + ((span >> 5) * lane) + (((span & 31) * lane) >> 5);   // floor(span * lane / 32)
        const bool before = (int64_t)fastdiv((uint64_t)__ldg(p.idx + m), p.slab_div) < target;
        const int c = __popc(__ballot_sync(0xffffffffu, before));
        if (c == 0) {
            hi = lo;
            break;
        }
        const int64_t m_last_true = __shfl_sync(0xffffffffu, m, c - 1);
        const int64_t m_first_false = __shfl_sync(0xffffffffu, m, c < 32 ? c : 31);
        lo = m_last_true + 1;
        if (c < 32) hi = m_first_false;
    }
    if (lane == 0) p.bounds[t] = lo;
}

// kSlabItems consecutive elements of a thread: 128-bit loads when the arrays allow it (16-byte aligned bases,
// the whole group inside the array), else one by one
template <typename E>
__device__ __forceinline__ void load_items(const E* __restrict__ base, int64_t i0, int64_t n, bool vec, E (&out)[kSlabItems]) {
    constexpr int kPerVec = 16 / (int)sizeof(E);
    if constexpr (kPerVec <= kSlabItems) {
        if (vec && i0 + kSlabItems <= n) {
            const uint4* src = reinterpret_cast<const uint4*>(base + i0);
#pragma unroll
            for (int q = 0; q < kSlabItems / kPerVec; ++q) {
                const uint4 w = __ldg(src + q);
                *reinterpret_cast<uint4*>(&out[q * kPerVec]) = w;
            }
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < kSlabItems; ++k) out[k] = (i0 + k < n) ? __ldg(base + i0 + k) : E(0);
}

// pass 1, all threads: the tile's count goes out, also into its group's counter
__device__ __forceinline__ void publish_count(const SlabParams& p, int64_t tile, int mine, int* s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) s_w[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
#pragma unroll
        for (int w = 0; w < kSlabNT / 32; ++w) total += s_w[w];
        p.counts[tile] = total;
        if (total) atomicAdd(p.group_cnt + (tile >> p.group_shift), (unsigned long long)total);
    }
}

// pass 2, all threads: output offset of the tile = the groups in front of its own + the earlier tiles of its own
__device__ __forceinline__ int64_t tile_offset(const SlabParams& p, int64_t tile, int64_t* s_red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t g = tile >> p.group_shift;
    int64_t acc = 0;
    for (int64_t j = tid; j < g; j += kSlabNT) acc += (int64_t)__ldg(p.group_cnt + j);
    for (int64_t j = (g << p.group_shift) + tid; j < tile; j += kSlabNT) acc += __ldg(p.counts + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    int64_t total = 0;
#pragma unroll
    for (int w = 0; w < kSlabNT / 32; ++w) total += s_red[w];
    return total;
}

// ---- reduced axis last --------------------------------------------------------------------------------------
template <typename T, bool COUNT>
__global__ void __launch_bounds__(kSlabNT, COUNT ? 6 : kSlabCtasPerSm) slab_last_kernel(const SlabParams p) {
    __shared__ double s_sum[COUNT ? 1 : kSlabCapCells];        // the tile's results, compact, in key order
    __shared__ uint16_t s_key[COUNT ? 1 : kSlabCapCells];      // their slabs, relative to the tile's first
    __shared__ int s_w[kSlabNT / 32];
    __shared__ int64_t s_red[kSlabNT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const T* const vals = reinterpret_cast<const T*>(p.val);
    if constexpr (!COUNT) {
        const int slots = (int)p.slabs_per_tile;               // <= kSlabCapCells: a tile cannot emit more
        for (int c = tid; c < slots; c += kSlabNT) s_sum[c] = 0.0;
    }
    if constexpr (COUNT) cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();               // boundaries (and counts) come from the previous launches
    const int64_t e0 = __ldg(p.bounds + tile), e1 = __ldg(p.bounds + tile + 1);
    const int64_t slab0 = tile * p.slabs_per_tile;
    const bool vec = ((reinterpret_cast<uintptr_t>(p.idx) | reinterpret_cast<uintptr_t>(vals)) & 15) == 0;
    const int64_t a0 = e0 & ~(int64_t)3;           // thread stretches start 16-byte aligned in both arrays
    int run_base = 0;                              // run heads in the batches already done
    int nheads_all = 0;
    if constexpr (!COUNT) __syncthreads();         // the slots are zero
    for (int64_t b0 = a0; b0 < e1; b0 += kSlabBatch) {
        const int64_t i0 = b0 + (int64_t)tid * kSlabItems;
        int64_t key[kSlabItems];
        [[maybe_unused]] T v[kSlabItems];
        int64_t key_before = -1;
        if (i0 < e1) {
            load_items<int64_t>(p.idx, i0, p.n, vec, key);
            if constexpr (!COUNT) load_items<T>(vals, i0, p.n, vec, v);
            if (lane == 0 && i0 > e0) key_before = __ldg(p.idx + i0 - 1);     // the element before the warp's first
        }
        // slab of every element relative to the tile's first slab; -1 outside the tile (elements klo .. khi-1 of
        // the stretch are inside).  The kind of division is decided once, not per element.
        const int klo = (int)(e0 - i0 > 0 ? (e0 - i0 < kSlabItems ? e0 - i0 : kSlabItems) : 0);
        const int khi = (int)(e1 - i0 < kSlabItems ? (e1 - i0 > 0 ? e1 - i0 : 0) : kSlabItems);
        int rel[kSlabItems];
        int prev;
        if (p.slab_div.magic == 0) {
            const int sh = (int)p.slab_div.shift;
#pragma unroll
            for (int k = 0; k < kSlabItems; ++k)
                rel[k] = (k >= klo && k < khi) ? (int)((int64_t)((uint64_t)key[k] >> sh) - slab0) : -1;
            prev = key_before >= 0 ? (int)((int64_t)((uint64_t)key_before >> sh) - slab0) : -1;
        } else {
#pragma unroll
            for (int k = 0; k < kSlabItems; ++k)
                rel[k] = (k >= klo && k < khi) ? (int)((int64_t)fastdiv((uint64_t)key[k], p.slab_div) - slab0) : -1;
            prev = key_before >= 0 ? (int)((int64_t)fastdiv((uint64_t)key_before, p.slab_div) - slab0) : -1;
        }
        const int up = __shfl_up_sync(0xffffffffu, rel[kSlabItems - 1], 1);
        if (lane != 0) prev = up;
        unsigned heads = 0;
#pragma unroll
        for (int k = 0; k < kSlabItems; ++k) {
            heads |= (unsigned)(rel[k] >= 0 && rel[k] != (k == 0 ? prev : rel[k - 1])) << k;
        }
        const int nh = __popc(heads);
        if constexpr (COUNT) {
            nheads_all += nh;
        } else {
            // slot of a run = number of heads in front of it: block scan of the head counts
            int incl = nh;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_w[warp] = incl;
            __syncthreads();
            int wofs = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kSlabNT / 32; ++w) {
                const int s = s_w[w];
                if (w < warp) wofs += s;
                total += s;
            }
            int slot = run_base + wofs + incl - nh - 1;        // the run that is open when my stretch begins
            double run = 0.0;
            bool open = false;                                 // a run is being added up
            bool mine = false;                                 // ... and it started inside my stretch
#pragma unroll
            for (int k = 0; k < kSlabItems; ++k) {
                if (rel[k] >= 0) {
                    if ((heads >> k) & 1u) {
                        if (open) {
                            // this run ends inside my stretch: mine alone unless it came in from the left
                            if (mine) s_sum[slot] = run;
                            else atomicAdd(&s_sum[slot], run);
                        }
                        ++slot;
                        s_key[slot] = (uint16_t)rel[k];
                        run = (double)v[k] + 0.0;          // (a stored -0.0 counts as +0.0, as in np.bincount)
                        mine = true;
                    } else {
                        run += (double)v[k];
                    }
                    open = true;
                }
            }
            if (open) atomicAdd(&s_sum[slot], run);            // may go on in the next thread's stretch
            run_base += total;
            __syncthreads();                                   // s_w is re-used by the next batch
        }
    }
    if constexpr (COUNT) {
        publish_count(p, tile, nheads_all, s_w);
    } else {
        const int64_t gbase = tile_offset(p, tile, s_red);     // (its barrier also completes the results)
        if (tile == p.ntiles - 1 && tid == 0) *p.out_count = gbase + run_base;
        for (int r = tid; r < run_base; r += kSlabNT) {
            p.out_idx[gbase + r] = slab0 + s_key[r];
            p.out_sums[gbase + r] = s_sum[r];
        }
    }
}

// ---- any other axis: accumulators indexed by the tile-local cell ---------------------------------------------
template <typename T, bool COUNT>
__global__ void __launch_bounds__(kSlabNT, kSlabCtasPerSm) slab_cells_kernel(const SlabParams p) {
    __shared__ double s_acc[COUNT ? 1 : kSlabCapCells];        // the tile's output cells
    __shared__ uint16_t s_list[COUNT ? 1 : kSlabCapCells];     // touched cells in key order
    __shared__ __align__(16) uint8_t s_flag[kSlabCapCells];    // cell touched (explicit zeros survive)
    __shared__ int s_w[kSlabNT / 32];
    __shared__ int64_t s_red[kSlabNT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const T* const vals = reinterpret_cast<const T*>(p.val);
    const int64_t suffix = (int64_t)p.suf_div.d;
    const int cells_tile = (int)(p.slabs_per_tile * suffix);
    // every thread owns a fixed field of `cpt` cells (a power of two <= 16)
    int cpt = 1;
    while (cpt * kSlabNT < cells_tile) cpt <<= 1;
    const int c0 = tid * cpt;
    if constexpr (!COUNT)
        for (int c = tid; c < cells_tile; c += kSlabNT) s_acc[c] = 0.0;     // (interleaved: no bank conflicts)
    for (int c = tid; c < (cpt * kSlabNT + 3) / 4; c += kSlabNT) reinterpret_cast<uint32_t*>(s_flag)[c] = 0;
    if constexpr (COUNT) cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();               // boundaries (and counts) come from the previous launches
    const int64_t e0 = __ldg(p.bounds + tile), e1 = __ldg(p.bounds + tile + 1);
    const int64_t slab0 = tile * p.slabs_per_tile;
    const bool vec = ((reinterpret_cast<uintptr_t>(p.idx) | reinterpret_cast<uintptr_t>(vals)) & 15) == 0;
    __syncthreads();

    const int64_t a0 = e0 & ~(int64_t)3;
    for (int64_t i0 = a0 + (int64_t)tid * kSlabItems; i0 < e1; i0 += kSlabBatch) {
        int64_t key[kSlabItems];
        [[maybe_unused]] T v[kSlabItems];
        load_items<int64_t>(p.idx, i0, p.n, vec, key);
        if constexpr (!COUNT) load_items<T>(vals, i0, p.n, vec, v);
#pragma unroll
        for (int k = 0; k < kSlabItems; ++k) {
            const int64_t i = i0 + k;
            if (i >= e0 && i < e1) {
                const uint64_t slab = fastdiv((uint64_t)key[k], p.slab_div);
                const uint64_t rem = (uint64_t)key[k] - slab * p.slab_div.d;
                const uint64_t q = fastdiv(rem, p.suf_div);
                const int cell = (int)(((int64_t)slab - slab0) * suffix + (int64_t)(rem - q * p.suf_div.d));
                if constexpr (!COUNT) atomicAdd(&s_acc[cell], (double)v[k]);
                s_flag[cell] = 1;
            }
        }
    }
    __syncthreads();

    // ---- count the touched cells of my field; list them in key order ---------------------------------------
    unsigned bits = 0;
    if (cpt >= 4) {
        // four flag bytes (0 / 1) -> four bits with one multiply
        for (int c = 0; c < cpt; c += 4)
            bits |= ((*reinterpret_cast<const uint32_t*>(s_flag + c0 + c) * 0x01020408u) >> 24) << c;
    } else {
        for (int c = 0; c < cpt; ++c) bits |= (unsigned)s_flag[c0 + c] << c;
    }
    const int cnt = __popc(bits);
    if constexpr (COUNT) {
        publish_count(p, tile, cnt, s_w);
    } else {
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int wofs = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kSlabNT / 32; ++w) {
            const int s = s_w[w];
            if (w < warp) wofs += s;
            total += s;
        }
        int o = wofs + incl - cnt;
        while (bits) {
            s_list[o++] = (uint16_t)(c0 + __ffs(bits) - 1);
            bits &= bits - 1;
        }
        const int64_t gbase = tile_offset(p, tile, s_red);     // (its barrier also completes the list)
        if (tile == p.ntiles - 1 && tid == 0) *p.out_count = gbase + total;
        const int64_t key0 = slab0 * suffix;
        for (int r = tid; r < total; r += kSlabNT) {
            const int cell = s_list[r];
            p.out_idx[gbase + r] = key0 + cell;
            p.out_sums[gbase + r] = s_acc[cell];
        }
    }
}

template <typename T>
static int launch_slab_sum(const SlabParams& p, bool last, cudaStream_t stream) {
    static PerDevice configured;           // five ~45 KB CTAs per SM need the large shared-memory carve-out
    if (configured.first_use(current_device())) {
        SPB_CUDA_TRY(cudaFuncSetAttribute(slab_last_kernel<T, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        SPB_CUDA_TRY(cudaFuncSetAttribute(slab_cells_kernel<T, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    const dim3 grid((unsigned)p.ntiles), block(kSlabNT);
    // pass 1 reads the keys only (the value type plays no role: one instantiation serves every dtype)
    if (last) {
        SPB_CUDA_TRY(launch_pdl(slab_last_kernel<uint8_t, true>, grid, block, 0, stream, p));
        SPB_CUDA_TRY(launch_pdl(slab_last_kernel<T, false>, grid, block, 0, stream, p));
    } else {
        SPB_CUDA_TRY(launch_pdl(slab_cells_kernel<uint8_t, true>, grid, block, 0, stream, p));
        SPB_CUDA_TRY(launch_pdl(slab_cells_kernel<T, false>, grid, block, 0, stream, p));
    }
    SPB_LAUNCHED(2);
    return SPB_OK;
}

// Tile geometry, or false when this path does not apply (the caller then takes the dense-accumulator or the
// sort-based path).  A tile is `slabs` whole slabs: about kSlabTileElems stored elements if the array were
// uniformly filled, never more than kSlabCapCells output cells.
static bool slab_plan(int64_t n, int64_t total_cells, int64_t slab_cells, int64_t suffix, int64_t* slabs, int64_t* ntiles) {
    if (n < 1 || total_cells < 1 || slab_cells < 1 || suffix < 1 || suffix > kSlabCapCells) return false;
    if (slab_cells % suffix != 0) return false;
    const int64_t total_slabs = ceil_div(total_cells, slab_cells);
    double s = (double)kSlabTileElems * (double)total_slabs / (double)n;
    const int64_t cap = kSlabCapCells / suffix;
    int64_t S = s >= (double)cap ? cap : (int64_t)s;
    if (S < 1) {
        // slabs longer than a tile: fine as long as one CTA can chew through a slab in reasonable time
        if ((double)n / (double)total_slabs > 16.0 * kSlabTileElems) return false;
        S = 1;
    }
    const int64_t nt = ceil_div(total_slabs, S);
    // very sparse arrays: most tiles would be empty
    if (nt > 4 * ceil_div(n, kSlabTileElems) + 2048 || nt > (1LL << 30)) return false;
    *slabs = S;
    *ntiles = nt;
    return true;
}

}  // namespace spb

using namespace spb;

extern "C" int spb_axis_sum_slab_ok(int64_t n, int64_t total_cells, int64_t slab_cells, int64_t suffix) {
    int64_t s, nt;
    return slab_plan(n, total_cells, slab_cells, suffix, &s, &nt) ? 1 : 0;
}

extern "C" int spb_axis_sum_slab(int dtype, const int64_t* idx, const void* val, int64_t n, int64_t total_cells,
                                 int64_t slab_cells, int64_t suffix, int64_t* out_idx, double* out_sums,
                                 int64_t* out_count, spb_workspace* ws, spb_stream_t stream) {
    if (n < 0 || !out_count) return SPB_ERR_BAD_ARG;
    if (n == 0) return (int)cudaMemsetAsync(out_count, 0, 8, stream);
    if (!idx || !val || !out_idx || !out_sums || !ws) return SPB_ERR_BAD_ARG;
    SlabParams p{};
    if (!slab_plan(n, total_cells, slab_cells, suffix, &p.slabs_per_tile, &p.ntiles)) return SPB_ERR_UNSUPPORTED;
    p.idx = idx; p.val = val; p.n = n;
    p.slab_div = make_fastdiv((uint64_t)slab_cells);
    p.suf_div = make_fastdiv((uint64_t)suffix);
    p.total_slabs = ceil_div(total_cells, slab_cells);
    p.out_idx = out_idx; p.out_sums = out_sums; p.out_count = out_count;
    // groups of 2^shift tiles, at most kSlabMaxGroups of them: a CTA sums < 1024 group counters and < 2^shift
    // tile counts to find its offset
    p.group_shift = 6;
    while (ceil_div(p.ntiles, (int64_t)1 << p.group_shift) > kSlabMaxGroups) ++p.group_shift;
    p.ngroups = ceil_div(p.ntiles, (int64_t)1 << p.group_shift);
    const size_t bounds_bytes = (size_t)(p.ntiles + 1) * 8, group_bytes = (size_t)p.ngroups * 8;
    int rc = workspace_reserve_payload(ws, bounds_bytes + group_bytes + (size_t)p.ntiles * 4, stream);
    if (rc) return rc;
    char* const pay = reinterpret_cast<char*>(ws->payload);
    p.bounds = reinterpret_cast<int64_t*>(pay);
    p.group_cnt = reinterpret_cast<unsigned long long*>(pay + bounds_bytes);
    p.counts = reinterpret_cast<int*>(pay + bounds_bytes + group_bytes);
    int64_t pblocks = ceil_div(p.ntiles + 1, 8);
    if (pblocks * 256 < p.ngroups) pblocks = ceil_div(p.ngroups, 256);
    SPB_CUDA_TRY(launch_pdl(slab_partition_kernel, dim3((unsigned)pblocks), dim3(256), 0, stream, p));
    SPB_LAUNCHED(1);
    const bool last = suffix == 1;
    switch (dtype) {
        case SPB_F32: rc = launch_slab_sum<float>(p, last, stream); break;
        case SPB_F64: rc = launch_slab_sum<double>(p, last, stream); break;
        case SPB_I32: rc = launch_slab_sum<int32_t>(p, last, stream); break;
        case SPB_I64: rc = launch_slab_sum<int64_t>(p, last, stream); break;
        case SPB_U8: case SPB_BOOL: rc = launch_slab_sum<uint8_t>(p, last, stream); break;
        case SPB_I8: rc = launch_slab_sum<int8_t>(p, last, stream); break;
        case SPB_I16: rc = launch_slab_sum<int16_t>(p, last, stream); break;
        default: rc = SPB_ERR_UNSUPPORTED; break;
    }
    if (rc) return rc;
    return (int)cudaGetLastError();
}
