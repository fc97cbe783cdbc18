// sparray_b200/csrc/common.cuh -- shared device helpers (sm_100a only).
//
//  * TMA bulk copies (cp.async.bulk + mbarrier) for staging unaligned ranges of
//    a global array into shared memory,
//  * a decoupled look-back tile-prefix protocol whose state words carry a launch
//    epoch, so the scratch never has to be memset between launches,
//  * warp-cooperative 32-ary merge-path search on global memory,
//  * exact unsigned 64-bit division by an invariant divisor (shift or mul-hi).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/sparray_b200.h"

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "sparray_b200 kernels are written for sm_100a only (compile with -gencode arch=compute_100a,code=sm_100a)"
#endif
#endif

// Scratch owned by the caller (one per stream in flight): look-back tile state words
// (epoch-tagged, see below) followed by whatever else a kernel reserves.
struct spb_workspace {
    void* dev;            // [misc | epoch-tagged look-back state words]
    size_t bytes;
    uint32_t epoch;
    int device;
    void* payload;        // plain data that rides along with the look-back (never read as state)
    size_t payload_bytes;
    void* state16;        // 16-byte {tagged word, value} look-back entries (only ever written as such)
    size_t state16_bytes;
    void* accum;          // dense float64 accumulators: every cell holds -0.0 ("untouched") between calls -- the
    size_t accum_bytes;   // consumer puts back what it reads (see csrc/axis_sum.cu)
    void* flags;          // one "touched" byte per cell for the sums that use flags: all zero between calls
    size_t flags_bytes;
    int accum_dirty;      // an accumulate pass was enqueued whose emitting pass was not: wipe before the next use
    size_t last_state_bytes;     // look-back words / 16-byte entries the call in progress asked for: what a
    size_t last_state16_bytes;   // captured launch wipes before it runs (see workspace_next_epoch)
};

namespace spb {

// number of kernels this library has launched (read back through spb_launch_count)
extern std::atomic<unsigned long long> g_launches;
#define SPB_LAUNCHED(k) (::spb::g_launches.fetch_add((k), std::memory_order_relaxed))

// Function attributes (shared-memory carve-out, max dynamic shared memory) and occupancy answers are
// per DEVICE: a process driving several GPUs must configure a kernel once on each.  One of these per
// kernel instantiation; first_use() is true exactly once per device (and thread-safe).
struct PerDevice {
    std::atomic<uint64_t> seen{0};
    int value[64] = {};
    bool first_use(int dev) {
        const uint64_t bit = 1ull << (dev & 63);
        return (seen.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
    }
};
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

// Workspace layout: [0, kWsMiscBytes) plain scratch (reduction partials, histograms);
// [kWsMiscBytes, ...) epoch-tagged look-back state words.  Only state words are ever
// written to the second region, so a stale word can never alias a live epoch.
constexpr size_t kWsMiscBytes = 1u << 20;
// A few self-resetting counters sit behind the state words: zero whenever no launch is in flight
// (the kernel that consumes one puts it back to zero), so they need no memset per call either.
constexpr size_t kWsCounterBytes = 256;
int workspace_reserve(spb_workspace* ws, size_t state_bytes, cudaStream_t stream);   // workspace.cu
int workspace_next_epoch(spb_workspace* ws, cudaStream_t stream, uint32_t* epoch, bool wipe_in_capture = true);
int workspace_reserve_payload(spb_workspace* ws, size_t bytes, cudaStream_t stream);
int workspace_reserve_state16(spb_workspace* ws, size_t bytes, cudaStream_t stream);
int workspace_reserve_accum(spb_workspace* ws, size_t acc_bytes, size_t flag_bytes, cudaStream_t stream);
int workspace_wipe_accum(spb_workspace* ws, cudaStream_t stream);
constexpr unsigned long long kNegZeroBits = 0x8000000000000000ull;     // bit pattern of -0.0
int device_sm_count();
inline uint64_t* ws_state(spb_workspace* ws) { return reinterpret_cast<uint64_t*>((char*)ws->dev + kWsMiscBytes); }
inline char* ws_misc(spb_workspace* ws) { return (char*)ws->dev; }
inline unsigned long long* ws_counters(spb_workspace* ws) { return reinterpret_cast<unsigned long long*>((char*)ws->dev + ws->bytes); }

constexpr int kWarp = 32;
constexpr int64_t kKeyMax = INT64_MAX;
constexpr int64_t kKeyMin = INT64_MIN;

#define SPB_CUDA_TRY(expr)                          \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) return (int)_e;      \
    } while (0)

// Launch with programmatic dependent launch: the kernel may be scheduled while its predecessor on
// the stream drains; it must call cudaGridDependencySynchronize() before touching the
// predecessor's results (the kernels launched this way do so first thing).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// mbarrier + TMA bulk copy (global -> shared::cta), PTX ISA 8.x
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    // make the init visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(phase), "r"(0x989680u)     // suspend-time hint: sleep in hardware instead of re-issuing the poll
        : "memory");
}

// dst, src 16-byte aligned; bytes a non-zero multiple of 16
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Plan for staging global elements [g_lo, g_hi) of an array with element size E
// (E divides 16) into a 16-byte aligned shared buffer.  Element i of the global
// array lands at byte offset (base + i*E) - align_down(base + g_lo*E, 16) from
// the buffer start, so the aligned middle can go by TMA and the < 16-byte head
// and tail by ordinary loads.
struct StagePlan {
    uintptr_t g_al;    // global address that maps to buffer offset 0
    uintptr_t q_lo;    // TMA source range [q_lo, q_hi)
    uintptr_t q_hi;
    uintptr_t p_lo;    // full byte range [p_lo, p_hi)
    uintptr_t p_hi;
    __device__ __forceinline__ uint32_t tma_bytes() const { return q_hi > q_lo ? (uint32_t)(q_hi - q_lo) : 0u; }
    __device__ __forceinline__ uint32_t span_bytes() const { return (uint32_t)(p_hi - g_al); }
};

__device__ __forceinline__ StagePlan make_stage_plan(const void* base, int64_t g_lo, int64_t g_hi, int esize) {
    StagePlan s;
    s.p_lo = reinterpret_cast<uintptr_t>(base) + (uintptr_t)g_lo * esize;
    s.p_hi = reinterpret_cast<uintptr_t>(base) + (uintptr_t)g_hi * esize;
    s.g_al = s.p_lo & ~(uintptr_t)15;
    s.q_lo = (s.p_lo + 15) & ~(uintptr_t)15;
    s.q_hi = s.p_hi & ~(uintptr_t)15;
    if (s.q_hi < s.q_lo) s.q_hi = s.q_lo;
    return s;
}

// Issue the TMA part (one thread) -- returns bytes that will be signalled.
__device__ __forceinline__ uint32_t stage_issue_tma(const StagePlan& s, char* buf, uint64_t* bar) {
    uint32_t nb = s.tma_bytes();
    if (nb) tma_bulk_g2s(buf + (s.q_lo - s.g_al), reinterpret_cast<const void*>(s.q_lo), nb, bar);
    return nb;
}

// Copy head and tail (< 16 bytes each) with plain loads; call from up to 8 threads
// (lane = 0..7), element size E in {1,2,4,8}.
template <typename E>
__device__ __forceinline__ void stage_copy_edges(const StagePlan& s, char* buf, int lane) {
    constexpr int per = 16 / (int)sizeof(E);
    uintptr_t head_end = s.q_lo < s.p_hi ? s.q_lo : s.p_hi;
    if (lane < per) {
        uintptr_t a = s.p_lo + (uintptr_t)lane * sizeof(E);
        if (a < head_end) *reinterpret_cast<E*>(buf + (a - s.g_al)) = *reinterpret_cast<const E*>(a);
        uintptr_t tail_lo = s.q_hi > s.p_lo ? s.q_hi : s.p_lo;
        if (tail_lo < head_end) tail_lo = head_end;
        uintptr_t b = tail_lo + (uintptr_t)lane * sizeof(E);
        if (b < s.p_hi) *reinterpret_cast<E*>(buf + (b - s.g_al)) = *reinterpret_cast<const E*>(b);
    }
}

// ---------------------------------------------------------------------------
// Decoupled look-back over 64-bit tile state words:
//   [63:62] status (0 invalid, 1 tile aggregate, 2 inclusive prefix)
//   [61:48] launch epoch (stale words of earlier launches read as invalid)
//   [47:0]  count
// ---------------------------------------------------------------------------
constexpr uint64_t kStAgg = 1ull, kStPrefix = 2ull;
constexpr int kEpochBits = 14;
constexpr uint64_t kCountMask = (1ull << 48) - 1;

__device__ __forceinline__ uint64_t pack_state(uint64_t status, uint32_t epoch, uint64_t count) {
    return (status << 62) | ((uint64_t)epoch << 48) | (count & kCountMask);
}

__device__ __forceinline__ void st_state(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint64_t ld_state(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by one full warp.  Returns the exclusive prefix (sum of counts of tiles
// 0..tile-1) in every lane.
__device__ __forceinline__ uint64_t lookback_exclusive(const uint64_t* state, int64_t tile, uint32_t epoch) {
    const int lane = threadIdx.x & 31;
    uint64_t total = 0;
    int64_t base = tile - 1;
    while (base >= 0) {
        int64_t t = base - lane;
        uint64_t w;
        if (t >= 0) {
            for (;;) {
                w = ld_state(state + t);
                if ((w >> 62) != 0 && (uint32_t)((w >> 48) & ((1u << kEpochBits) - 1)) == epoch) break;
                __nanosleep(20);
            }
        } else {
            w = pack_state(kStPrefix, epoch, 0);
        }
        unsigned has_prefix = __ballot_sync(0xffffffffu, (w >> 62) == kStPrefix);
        int stop = has_prefix ? (__ffs(has_prefix) - 1) : 31;
        uint64_t v = (lane <= stop) ? (w & kCountMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        total += v;
        if (has_prefix) break;
        base -= 32;
    }
    return total;
}

// Two-level look-back.  The plain walk above moves 32 tiles per L2 round trip, and a tile can only stop at a
// predecessor whose own walk is over: once tiles start faster than about 32 per round trip (a 2048-element tile
// is ~5 ns of HBM time on a B200, a round trip ~500 ns) the distance to the nearest finished walk keeps
// growing and the chain, not the memory system, sets the pace (measured: 30-55 tiles/us whatever the kernel
// did per tile).  So the tiles are grouped in BLOCKS of 32 with one more state word each, behind the ntiles tile
// words: the last tile of a block publishes the block's AGGREGATE as soon as the 31 counts before it are out
// (no dependence on anybody's prefix), and its inclusive prefix when its own walk is over.  A walk is then:
// one window over the earlier tiles of the own block, then windows of 32 BLOCK words (1024 tiles per round
// trip).  Every wait is for a count that is published unconditionally: no chain.
// Called by one full warp after the caller has published the tile's own word (aggregate; prefix for tile 0);
// `total` is the tile's own count.  Not for tile 0.  Returns the exclusive prefix in every lane.
__device__ __forceinline__ uint64_t wait_state(const uint64_t* p, uint32_t epoch) {
    uint64_t w;
    for (;;) {
        w = ld_state(p);
        if ((w >> 62) != 0 && (uint32_t)((w >> 48) & ((1u << kEpochBits) - 1)) == epoch) break;
        __nanosleep(20);
    }
    return w;
}

__host__ __device__ inline int64_t lookback2_words(int64_t ntiles) { return ntiles + (ntiles >> 5) + 1; }

__device__ __forceinline__ uint64_t lookback_exclusive_2level(uint64_t* state, int64_t ntiles, int64_t tile,
                                                              uint32_t epoch, uint64_t total) {
    const int lane = threadIdx.x & 31;
    uint64_t* const blocks = state + ntiles;
    const int64_t b = tile >> 5;
    const int j = (int)(tile & 31);
    uint64_t excl;
    bool done;
    {   // the earlier tiles of my own block
        const bool in = lane < j;
        uint64_t w = 0;
        if (in) w = wait_state(state + (tile - 1 - lane), epoch);
        const unsigned has_prefix = __ballot_sync(0xffffffffu, in && (w >> 62) == kStPrefix);
        const int stop = has_prefix ? (__ffs(has_prefix) - 1) : 31;
        uint64_t v = (in && lane <= stop) ? (w & kCountMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl = v;
        done = has_prefix != 0;
    }
    // the block's aggregate goes out as soon as it is known
    if (!done && j == 31 && lane == 0) st_state(blocks + b, pack_state(kStAgg, epoch, excl + total));
    int64_t bb = b - 1;
    while (!done) {
        const int64_t q = bb - lane;
        uint64_t w = pack_state(kStPrefix, epoch, 0);          // in front of block 0: an empty prefix
        if (q >= 0) w = wait_state(blocks + q, epoch);
        const unsigned has_prefix = __ballot_sync(0xffffffffu, (w >> 62) == kStPrefix);
        const int stop = has_prefix ? (__ffs(has_prefix) - 1) : 31;
        uint64_t v = lane <= stop ? (w & kCountMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        done = has_prefix != 0;
        bb -= 32;
    }
    if (j == 31 && lane == 0) st_state(blocks + b, pack_state(kStPrefix, epoch, excl + total));
    return excl;
}

// Block-wide variant for 256-thread CTAs: every thread inspects one predecessor tile, so a
// whole wave of concurrently running tiles (persistent kernels: a few hundred) is covered in one
// or two rounds instead of a dozen dependent warp-wide ones.  `scratch` = 24 ints (16-byte aligned) of shared
// memory.  Must be called by all 256 threads; returns the exclusive prefix in every thread.
__device__ __forceinline__ uint64_t lookback_exclusive_block256(const uint64_t* state, int64_t tile, uint32_t epoch,
                                                                int* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* s_stop = scratch;                                         // [8] nearest prefix per warp
    unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(scratch + 8);   // [8] (16-byte aligned scratch)
    uint64_t total = 0;
    int64_t base = tile - 1;
    while (base >= 0) {
        const int64_t t = base - tid;
        uint64_t w;
        if (t >= 0) {
            for (;;) {
                w = ld_state(state + t);
                if ((w >> 62) != 0 && (uint32_t)((w >> 48) & ((1u << kEpochBits) - 1)) == epoch) break;
                __nanosleep(20);
            }
        } else {
            w = pack_state(kStPrefix, epoch, 0);
        }
        const unsigned has_prefix = __ballot_sync(0xffffffffu, (w >> 62) == kStPrefix);
        if (lane == 0) s_stop[warp] = has_prefix ? warp * 32 + (__ffs(has_prefix) - 1) : 256;
        __syncthreads();
        int stop = 256;
#pragma unroll
        for (int i = 0; i < 8; ++i) stop = s_stop[i] < stop ? s_stop[i] : stop;
        uint64_t v = tid <= stop ? (w & kCountMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_sum[warp] = v;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) total += s_sum[i];
        __syncthreads();                                           // scratch is reused by the next round
        if (stop < 256) break;
        base -= 256;
    }
    return total;
}

// Block-wide (256 threads) exclusive scan of a per-thread count + decoupled look-back over
// tiles: returns the global output offset of this thread's first emitted element.  The last
// tile stores the grand total to *out_count.  Must be called by all 256 threads of the block.
// WIDE: the whole block looks back (256 predecessors per round) -- for kernels whose tiles all run
// in one or two waves and publish their counts at about the same moment, where the warp-wide walk
// would need ntiles/32 dependent round trips for the last tile.
template <bool WIDE = false>
__device__ __forceinline__ int64_t compact_offset_256(int cnt, uint64_t* state, uint32_t epoch, int64_t tile,
                                                      bool last_tile, int64_t* out_count) {
    __shared__ int cs_warp[8];
    __shared__ int64_t cs_base;
    __shared__ __align__(16) int cs_scratch[24];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();            // protect cs_* against a previous call in the same kernel
    if (lane == 31) cs_warp[warp] = incl;
    __syncthreads();
    int tile_total = 0;
    if (warp == 0 || WIDE) {
        int w = lane < 8 ? cs_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        tile_total = __shfl_sync(0xffffffffu, wi, 7);
        if constexpr (WIDE) {
            __syncthreads();    // every warp has read the per-warp totals
        }
        if (warp == 0 && lane < 8) cs_warp[lane] = wi - w;
    }
    if constexpr (WIDE) {
        if (threadIdx.x == 0)
            st_state(state + tile, pack_state(tile == 0 ? kStPrefix : kStAgg, epoch, (uint64_t)tile_total));
        uint64_t base = 0;
        if (tile > 0) {
            base = lookback_exclusive_block256(state, tile, epoch, cs_scratch);
            if (threadIdx.x == 0) st_state(state + tile, pack_state(kStPrefix, epoch, base + (uint64_t)tile_total));
        }
        if (threadIdx.x == 0) {
            cs_base = (int64_t)base;
            if (last_tile) *out_count = (int64_t)base + tile_total;
        }
    } else if (warp == 0) {
        uint64_t base = 0;
        if (tile == 0) {
            if (lane == 0) st_state(state, pack_state(kStPrefix, epoch, (uint64_t)tile_total));
        } else {
            if (lane == 0) st_state(state + tile, pack_state(kStAgg, epoch, (uint64_t)tile_total));
            base = lookback_exclusive(state, tile, epoch);
            if (lane == 0) st_state(state + tile, pack_state(kStPrefix, epoch, base + (uint64_t)tile_total));
        }
        if (lane == 0) {
            cs_base = (int64_t)base;
            if (last_tile) *out_count = (int64_t)base + tile_total;
        }
    }
    __syncthreads();
    return cs_base + cs_warp[warp] + incl - cnt;
}



// ---------------------------------------------------------------------------
// Exact floor(n / d) for 0 <= n < 2^63, 1 <= d < 2^63, d fixed per launch.
// d a power of two -> shift; otherwise round-up magic with a 64-bit mul-hi
// (Granlund-Montgomery / libdivide "branchfree" form restricted to 63-bit n).
// ---------------------------------------------------------------------------
struct FastDiv {
    uint64_t magic;   // 0 => power of two
    uint32_t shift;
    uint64_t d;
};

inline FastDiv make_fastdiv(uint64_t d) {
    FastDiv f;
    f.d = d;
    if ((d & (d - 1)) == 0) {
        f.magic = 0;
        f.shift = 0;
        while ((1ull << f.shift) < d) ++f.shift;
        return f;
    }
    // L = ceil(log2 d); m = floor(2^(64+L-1) / d) + 1 fits 64 bits because n < 2^63
    uint32_t L = 0;
    while ((1ull << L) < d) ++L;
    unsigned __int128 num = (unsigned __int128)1 << (64 + L - 1);
    f.magic = (uint64_t)(num / d) + 1;
    f.shift = L - 1;
    return f;
}

__device__ __forceinline__ uint64_t fastdiv(uint64_t n, const FastDiv& f) {
    if (f.magic == 0) return n >> f.shift;
    return __umul64hi(n, f.magic) >> f.shift;
}

}  // namespace spb
