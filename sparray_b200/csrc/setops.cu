// sparray_b200/csrc/setops.cu -- merge-path set union / intersection of sorted
// duplicate-free int64 index arrays with the value op fused in.
//
// Replaces
//   * merge_unique   (_merge.pyx:107-140) + the three mask gathers / scatters of
//     _pairwise_sparray (flat.py:416-421)                         -> union kernels
//   * merge_intersect (_merge.pyx:145-157) + the two searchsorted of the wrapper
//     (_merge.pyx:87-88) + the gathers of _pairwise_sparray_fixed_zero
//     (flat.py:429-433)                                           -> intersection kernels
//
// Launches per call (chained with programmatic dependent launch):
//   1. merge_partition_kernel: a half-warp per tile boundary finds the merge-path split of its
//      diagonal on the two global arrays (bracketed around the proportional guess, then 16-ary);
//      one thread per tile turns two neighbouring splits into a 128-byte staging descriptor.
//   2. setop_kernel: ONE tile per CTA, in blockIdx order; 8 (intersections) or 5-6 (unions) CTAs per
//      SM hide each other's latencies.  Per tile
//        - the staging warp loads the descriptor and issues TMA bulk copies (cp.async.bulk +
//          mbarrier) of the A / B index ranges (values are never staged); the < 16-byte
//          unaligned heads / tails and the halo keys ride in registers of that warp;
//        - the staged 64-bit keys are read as 32-bit tile-local offsets whenever the tile's key
//          span allows it (64-bit path otherwise), so the merge runs on single-register keys;
//        - every thread merge-path-splits the tile in shared memory and walks kVT items with a
//          branch-free step that only records three bit masks (took-A, matched, emitted); a key
//          present in both inputs is emitted once, by its A element (the B partner is found
//          through the halo key even across a tile boundary);
//        - block scan of the output counts; then
//          unions: a helper warp has been resolving the tile's output offset by decoupled
//          look-back since the tile was staged; items are materialised into a shared staging area
//          (32-bit keys; values read from global exactly once, op applied) and stored coalesced;
//          intersections: no look-back -- the tile owns private output slots (one atomic only if
//          it overflows them), only the set bits of the match mask are visited, the two values of
//          a match are gathered straight from global memory (unmatched values are never read).
//   3. merge_reorder_kernel (intersections): chunks into tile order, writes the count.
//      No lut / mask array ever exists and nothing is read twice.
#include "common.cuh"

namespace spb {

template <typename V>
__device__ __forceinline__ V np_minimum(V a, V b) {
    return a < b ? a : b;
}
template <typename V>
__device__ __forceinline__ V np_maximum(V a, V b) {
    return a > b ? a : b;
}
// numpy: NaN-propagating, and on equal operands (signed zeros) the SECOND wins
template <>
__device__ __forceinline__ float np_minimum<float>(float a, float b) {
    return a != a ? a : (a < b ? a : b);
}
template <>
__device__ __forceinline__ double np_minimum<double>(double a, double b) {
    return a != a ? a : (a < b ? a : b);
}
template <>
__device__ __forceinline__ float np_maximum<float>(float a, float b) {
    return a != a ? a : (a > b ? a : b);
}
template <>
__device__ __forceinline__ double np_maximum<double>(double a, double b) {
    return a != a ? a : (a > b ? a : b);
}

// arithmetic keeps the operand type; comparisons give a bool byte
template <typename V, typename VO>
__device__ __forceinline__ VO value_op(int op, V x, V y) {
    if constexpr (sizeof(VO) == sizeof(V)) {
        switch (op) {
            case SPB_ADD: return (VO)(x + y);
            case SPB_MUL: return (VO)(x * y);
            case SPB_MINIMUM: return (VO)np_minimum<V>(x, y);
            case SPB_MAXIMUM: return (VO)np_maximum<V>(x, y);
            case SPB_TAKE_A: return (VO)x;
            default: return (VO)y;   // SPB_TAKE_B
        }
    } else {
        switch (op) {
            case SPB_LT: return (VO)(x < y);
            case SPB_LE: return (VO)(x <= y);
            case SPB_GT: return (VO)(x > y);
            case SPB_GE: return (VO)(x >= y);
            case SPB_EQ: return (VO)(x == y);
            default: return (VO)(x != y);   // SPB_NE
        }
    }
}

struct SetopParams {
    const int64_t* a;
    const int64_t* b;
    int64_t na, nb;
    const void* a_val;
    const void* b_val;
    int64_t* out_idx;
    void* out_val;
    uint8_t* lut;        // union seam
    uint8_t* a_only;
    uint8_t* b_only;
    int64_t* a_pos;      // intersection seam
    int64_t* b_pos;
    int64_t* out_count;
    uint64_t* state;     // one look-back word per tile
    const uint64_t* desc;    // [ntiles][kDescWords] staging descriptors written by the partition kernel
    int64_t ntiles;
    uint32_t epoch;
    int op;
    // intersections only: outputs are rare, so a tile RESERVES its output slots with one atomic and a
    // small second pass puts the chunks in tile order -- no cross-tile look-back at all
    unsigned long long* slot_counter;
    int64_t* tile_slot;      // [ntiles]
    int* tile_cnt;           // [ntiles]
    int* group_cnt;          // [ceil(ntiles / kGroupTiles)] matches per group of tiles
    int64_t* tmp_idx;        // chunked outputs, min(na, nb) slots each
    void* tmp_val;
    int64_t* tmp_apos;
    int64_t* tmp_bpos;
    int negate_b;        // a - b is evaluated as a + (-b), base.py:45-46: a missing partner is +0, not -0
};

enum { KIND_UNION = 0, KIND_INTERSECT = 1 };

// Tile shapes (struct Cfg below).  Intersections: 128 merging threads x 24 items -- outputs are rare and go
// straight to global memory, so a 3072-item tile of keys is all the shared memory a CTA needs (8 CTAs per
// SM), and a thread's 12-round merge-path search is paid once per 24 items (256 x 12 measured 62.6 us per
// config-2 call, 128 x 24 58.9).  Unions stage a full tile of outputs: 256 threads x 10 (4-byte results) or
// x 6 (8-byte results).
constexpr int key_region_bytes(int nt, int vt) { return ((nt * vt + 4) * 8 + 128 + vt * 8 + 15) & ~15; }

// ---- launch 1: merge-path split of every tile + its staging descriptor ------------------------
// Half-warps search the tile diagonals (16-ary, bracketed around the proportional guess); one thread
// per tile then works out everything the merge kernel needs to stage the tile -- so that kernel's
// staging warp only loads 128 bytes and issues the copies.
//   word 0..3 : the Geom record (a0, b0, an|bn, offA|offB)
//   word 5..6 : 16-byte aligned global source of the TMA copy of the A keys, the B keys
//   word 9    : 2 x 16 bit  destination offsets inside the tile buffer
//   word 10   : 2 x 16 bit  TMA byte counts
//   word 11   : 2 x (4 bit head count, 4 bit tail count): elements before / after the aligned middle
constexpr int kDescWords = 16;

struct PartitionParams {
    const int64_t* a;
    const int64_t* b;
    int64_t na, nb;
    int tile_items;          // merged items per tile
    int64_t ntiles;
    uint64_t* desc;          // [ntiles][kDescWords]
    int* group_cnt;          // intersections: [ngroups] match counters to zero, else null
    int64_t ngroups;
    double frac_a;           // na / (na + nb): where the proportional guess of a diagonal's split lies
};

constexpr int64_t kGuessWindow = 8192;

// Diagonals (tile boundaries) per CTA of the partition kernel: 16 warps search 16 consecutive
// diagonals, then 15 threads turn the 15 tiles between them into staging descriptors.
constexpr int kPartTiles = 15;
constexpr int kPartThreads = 512;
constexpr int kGroupTiles = 64;          // intersections: tiles per group counter (see merge_reorder_kernel)

// One warp per diagonal.  The split is where the two key sequences cross, and sorted flat indices grow
// close to linearly with their position: so the search INTERPOLATES.  A probe at the current guess g gives
// delta = a[g] - b[d-1-g]; one step of the split moves a's key up by one average gap of a and b's key down
// by one of b, so the crossing is about delta / (gap_a + gap_b) positions away.  Two or three such probes
// (two keys each) land within a few elements of the split, and one window of 32 consecutive candidates
// (coalesced) pins it down.  ~1 KB of keys per diagonal instead of the ~12 KB a 32-ary search reads -- the
// partition launch of round 1 moved 60 % as many bytes as the merge itself.  Whatever the data does to the
// interpolation, every step keeps [lo, hi] a valid bracket and the tail is a plain 32-ary search.
__device__ __forceinline__ int64_t merge_split_warp(const PartitionParams& p, int64_t w) {
    const int lane = threadIdx.x & 31;
    const int64_t total = p.na + p.nb;
    int64_t d = w * (int64_t)p.tile_items;
    if (d > total) d = total;
    int64_t lo = d > p.nb ? d - p.nb : 0;
    int64_t hi = d < p.na ? d : p.na;
    // predicate "the split lies above m": a[m] <= b[d-1-m], defined for lo <= m < hi; true below the split
    if (hi - lo > 64) {
        // average key gaps of the two arrays (the same four L2-resident keys for every diagonal); single
        // precision is plenty: the estimate only has to land within a few elements
        const float span_a = (float)(__ldg(p.a + p.na - 1) - __ldg(p.a));
        const float span_b = (float)(__ldg(p.b + p.nb - 1) - __ldg(p.b));
        const float gap = __fdividef(span_a, (float)(p.na > 1 ? p.na - 1 : 1)) +
                          __fdividef(span_b, (float)(p.nb > 1 ? p.nb - 1 : 1));
        const float inv_gap = gap > 0.0f ? __fdividef(1.0f, gap) : 0.0f;
        // proportional first guess (double keeps 2^53 positions exact enough; no 128-bit division)
        int64_t g = (int64_t)((double)d * p.frac_a);
        for (int it = 0; it < 3 && hi - lo > 64; ++it) {
            g = g < lo ? lo : (g > hi - 1 ? hi - 1 : g);
            const int64_t ka = __ldg(p.a + g), kb = __ldg(p.b + (d - 1 - g));          // uniform: one transaction each
            if (ka <= kb) lo = g + 1;                                                // the split lies above g
            else hi = g;
            const int64_t mv = (int64_t)__float2ll_rn((float)(ka - kb) * inv_gap);
            g -= mv;
            if (mv > -12 && mv < 12) break;
        }
        // one window of 32 consecutive candidates around the estimate
        if (hi - lo > 32) {
            int64_t base = g - 16;
            if (base > hi - 32) base = hi - 32;
            if (base < lo) base = lo;
            const int64_t m = base + lane;
            const bool t = m < hi ? (__ldg(p.a + m) <= __ldg(p.b + (d - 1 - m))) : false;
            const int c = __popc(__ballot_sync(0xffffffffu, t));
            if (c > 0 && c < 32) return base + c;                // the crossing is inside the window
            if (c == 0) hi = base;                               // everything here is past the split
            else lo = base + 32 < hi ? base + 32 : hi;           // everything here is before it
        }
    }
    while (lo < hi) {
        const int64_t span = hi - lo;
        const int64_t m = lo + ((span >> 5) * lane) + (((span & 31) * lane) >> 5);   // floor(span * lane / 32)
        const bool take_a = __ldg(p.a + m) <= __ldg(p.b + (d - 1 - m));
        const unsigned bal = __ballot_sync(0xffffffffu, take_a);
        const int c = __popc(bal);
        if (c == 0) {
            hi = lo;
            break;
        }
        const int64_t m_last_true = __shfl_sync(0xffffffffu, m, c - 1);
        const int64_t m_first_false = __shfl_sync(0xffffffffu, m, c < 32 ? c : 31);
        lo = m_last_true + 1;
        if (c < 32) hi = m_first_false;
    }
    return lo;
}

__global__ void __launch_bounds__(kPartThreads) merge_partition_kernel(const PartitionParams p) {
    __shared__ int64_t s_split[16];
    cudaGridDependencySynchronize();               // the operands may come from the previous launch
    // intersections: the per-group match counters start from zero (the merge kernel, which adds to them,
    // waits for this whole grid)
    if (p.group_cnt) {
        const int64_t g = (int64_t)blockIdx.x * kPartThreads + threadIdx.x;
        if (g < p.ngroups) p.group_cnt[g] = 0;
    }
    {
        int64_t w = (int64_t)blockIdx.x * kPartTiles + (threadIdx.x >> 5);   // diagonal index 0..ntiles
        if (w > p.ntiles) w = p.ntiles;              // keep the warp converged on a valid diagonal
        const int64_t sp = merge_split_warp(p, w);
        if ((threadIdx.x & 31) == 0) s_split[threadIdx.x >> 5] = sp;
    }
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * kPartTiles + threadIdx.x;
    if (threadIdx.x >= kPartTiles || t >= p.ntiles) return;
    const int64_t total = p.na + p.nb;
    const int64_t d0 = t * (int64_t)p.tile_items;
    const int64_t d1 = (d0 + p.tile_items < total) ? d0 + p.tile_items : total;
    const int64_t a0 = s_split[threadIdx.x], a1 = s_split[threadIdx.x + 1];
    const int64_t b0 = d0 - a0, b1 = d1 - a1;
    // indices [x0-1, x1] (one halo key either side); values are never staged: the merge kernel reads
    // the few it needs (intersections) or each exactly once (unions) straight from global memory
    const int64_t ga_lo = a0 > 0 ? a0 - 1 : 0, ga_hi = a1 < p.na ? a1 + 1 : p.na;
    const int64_t gb_lo = b0 > 0 ? b0 - 1 : 0, gb_hi = b1 < p.nb ? b1 + 1 : p.nb;
    StagePlan pl[2];
    int dst0[2];             // buffer offset that the plan's g_al maps to
    pl[0] = make_stage_plan(p.a, ga_lo, ga_hi, 8);
    pl[1] = make_stage_plan(p.b, gb_lo, gb_hi, 8);
    dst0[0] = 16;                                                     // front pad: slot for index -1
    dst0[1] = dst0[0] + (int)((pl[0].span_bytes() + 15) & ~15u) + 32; // A's +1 sentinel, B's -1 slot
    uint64_t* d = p.desc + t * kDescWords;
    d[0] = (uint64_t)a0;
    d[1] = (uint64_t)b0;
    d[2] = (uint64_t)(uint32_t)(a1 - a0) | ((uint64_t)(uint32_t)(b1 - b0) << 32);
    const int offA = dst0[0] + (int)(pl[0].p_lo - pl[0].g_al) + (int)(a0 - ga_lo) * 8;
    const int offB = dst0[1] + (int)(pl[1].p_lo - pl[1].g_al) + (int)(b0 - gb_lo) * 8;
    d[3] = (uint64_t)(uint32_t)offA | ((uint64_t)(uint32_t)offB << 32);
    uint64_t w_dst = 0, w_cnt = 0, w_ht = 0;
    for (int i = 0; i < 2; ++i) {
        const StagePlan& s = pl[i];
        const uintptr_t head_end = s.q_lo < s.p_hi ? s.q_lo : s.p_hi;      // == q_lo unless the range is tiny
        const uintptr_t mid = head_end;                                    // the copy starts here
        const uint32_t bytes = s.tma_bytes();
        const uintptr_t tail_lo = mid + bytes;
        w_dst |= (uint64_t)(uint32_t)(dst0[i] + (int)(mid - s.g_al)) << (16 * i);
        w_cnt |= (uint64_t)bytes << (16 * i);
        const uint32_t nh = (uint32_t)((mid - s.p_lo) / 8), nt = (uint32_t)((s.p_hi - tail_lo) / 8);
        w_ht |= (uint64_t)(nh | (nt << 4)) << (8 * i);
        d[5 + i] = (uint64_t)mid;
    }
    d[9] = w_dst;
    d[10] = w_cnt;
    d[11] = w_ht;
}

// ---- shared-memory layout ------------------------------------------------------------------------
struct Geom {            // where a staged tile lives inside its buffer
    int64_t a0, b0;
    int an, bn;
    int offA, offB;      // byte offset of As64[0] / Bs64[0] from the buffer start
};

template <int NT, int VT, int OUT_KEYS, int OUT_PAYLOAD>   // merging threads; items/thread; staging bytes (keys, values / lut)
struct Layout {
    static constexpr int kTile = NT * VT;
    static constexpr int kHeader = 256;
    static constexpr int kBuf = key_region_bytes(NT, VT);                   // the staged input keys of one tile
    static constexpr int kStgKeys = OUT_KEYS;                            // compacted outputs waiting for their offset
    static constexpr int kStg = ((OUT_KEYS + OUT_PAYLOAD) + 15) & ~15;
    static constexpr int kTotal = kHeader + kBuf + kStg;
    static_assert(kBuf % 16 == 0, "the buffer must stay 16-byte aligned for TMA");
};

template <typename V, typename VO, int KIND, bool HAS_VAL, bool SEAM>
struct Cfg {
#ifndef SPB_UNION_NT
#define SPB_UNION_NT 256
#endif
#ifndef SPB_ISECT_NT
#define SPB_ISECT_NT 128
#endif
#ifndef SPB_ISECT_VT
#define SPB_ISECT_VT 24
#endif
#ifndef SPB_ISECT_CTAS
#define SPB_ISECT_CTAS 8
#endif
#ifndef SPB_UNION_VT4
#define SPB_UNION_VT4 10
#endif
#ifndef SPB_UNION_VT8
#define SPB_UNION_VT8 6
#endif
#ifndef SPB_UNION_CTAS4
#define SPB_UNION_CTAS4 5
#endif
#ifndef SPB_UNION_CTAS8
#define SPB_UNION_CTAS8 6
#endif
    static constexpr bool kWideOut = HAS_VAL && sizeof(VO) == 8;
    // merging threads per CTA and items per thread (a thread's 12-round merge-path search is paid once per kVT items)
    static constexpr int kNTc = KIND == KIND_UNION ? SPB_UNION_NT : SPB_ISECT_NT;
    static constexpr int kVT = KIND == KIND_UNION ? (kWideOut ? SPB_UNION_VT8 : SPB_UNION_VT4) : SPB_ISECT_VT;
    static constexpr int kTile = kNTc * kVT;
    static constexpr int kFixedSlots = kTile / 32;                       // private output slots per tile (intersections)
    static constexpr int kCap = KIND == KIND_UNION ? kTile : kTile / 2;  // outputs per tile, at most
    static constexpr int kPayload = KIND == KIND_INTERSECT ? 0
                                    : (HAS_VAL ? kCap * (int)sizeof(VO) : (SEAM ? kCap : 0));
    // unions resolve their output offset by look-back: a ninth warp does it (and the staging) so that
    // no merging warp is held up by the L2 round trips
    static constexpr int kHelperWarp = KIND == KIND_UNION ? kNTc / 32 : 0;
    static constexpr int kThreads = KIND == KIND_UNION ? kNTc + 32 : kNTc;
    static constexpr int kMinCtas = KIND == KIND_INTERSECT ? SPB_ISECT_CTAS : (kWideOut ? SPB_UNION_CTAS8 : SPB_UNION_CTAS4);   // what shared memory allows
    // unions stage their output keys as 32-bit tile-local offsets (tiles spanning >= 2^32 store directly);
    // input values are never staged: every one is read exactly once, straight from global memory
    using L = Layout<kNTc, kVT, KIND == KIND_INTERSECT ? 0 : kCap * 4, kPayload>;
};

struct EdgeReg {         // an unaligned head / tail key riding in a register until its buffer is live
    uint64_t raw;
    uint32_t dst;        // shared-memory address, 0 = nothing
};

__device__ __forceinline__ void edge_store(const EdgeReg& er) {
    if (er.dst) asm volatile("st.shared.u64 [%0], %1;" ::"r"(er.dst), "l"(er.raw) : "memory");
}

// Executed by the whole staging warp: put one tile in flight.  `dw` is this lane's word of the
// tile descriptor (lanes 0..15).  Lanes 0..1 each issue one TMA bulk copy (A keys, B keys), lanes
// 16..19 each pick up one ragged-edge key (an 8-byte element in front of / behind the 16-byte
// aligned middle) in a register, lanes 0..3 publish the geometry.
__device__ __forceinline__ void prefetch_tile(uint64_t dw, char* buf, uint64_t* bar, Geom* g, EdgeReg& er, int lane) {
    if (lane < 4) reinterpret_cast<uint64_t*>(g)[lane] = dw;
    const uint64_t w_dst = __shfl_sync(0xffffffffu, dw, 9);
    const uint64_t w_cnt = __shfl_sync(0xffffffffu, dw, 10);
    const uint32_t w_ht = (uint32_t)__shfl_sync(0xffffffffu, dw, 11);
    // which array this lane serves: TMA lanes by index, edge lanes by slot (A head, A tail, B head, B tail)
    const int s = lane - 16;
    const int arr = lane >= 16 ? ((s >> 1) & 1) : (lane & 1);
    const int tail = s & 1;
    const uint64_t src = __shfl_sync(0xffffffffu, dw, 5 + arr);
    const uint32_t dst = (uint32_t)(w_dst >> (16 * arr)) & 0xffffu;
    const uint32_t bytes = (uint32_t)(w_cnt >> (16 * arr)) & 0xffffu;
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(w_cnt & 0xffffu) + (uint32_t)((w_cnt >> 16) & 0xffffu));
    __syncwarp();
    if (lane < 2 && bytes) tma_bulk_g2s(buf + dst, reinterpret_cast<const void*>(src), bytes, bar);
    er.dst = 0;
    er.raw = 0;
    if (lane >= 16 && lane < 20 && src) {
        const int nh = (int)(w_ht >> (8 * arr)) & 15, nt = (int)(w_ht >> (8 * arr + 4)) & 15;
        if ((tail ? nt : nh) > 0) {
            const int64_t delta = tail ? (int64_t)bytes : -8;
            er.dst = smem_u32(buf + dst) + (uint32_t)(int32_t)delta;
            er.raw = __ldg(reinterpret_cast<const unsigned long long*>((uintptr_t)((int64_t)src + delta)));
        }
    }
}

// Key accessors over the staged int64 keys.  When a tile's keys span less than 2^32 the merge
// runs on single-register keys: the low word of each key minus the low word of the tile's
// smallest key (exact modulo 2^32 because the true offset fits 32 bits).
struct NarrowKeys {
    const uint32_t* w;       // low word of element 0 (little endian); element i is w[2 * i]
    uint32_t base;
    __device__ __forceinline__ uint32_t operator()(int i) const { return w[2 * i] - base; }
};
struct WideKeys {
    const int64_t* p;
    __device__ __forceinline__ int64_t operator()(int i) const { return p[i]; }
};

template <typename Acc>
__device__ __forceinline__ int merge_path_smem(const Acc& a, int na, const Acc& b, int nb, int d) {
    int lo = d > nb ? d - nb : 0;
    int hi = d < na ? d : na;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a(mid) <= b(d - 1 - mid)) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Branch-free walk of kVT merged items starting at (ai, bi): records who was taken (A or B),
// where both sides held the same key, and which items are emitted.  Ties go to A; the B
// partner of a matched A item is the very next item and is suppressed (prev_eq).
template <typename Acc, int KIND, int kVT>
__device__ __forceinline__ void merge_walk(const Acc& A, const Acc& B, int ai, int bi, int steps, bool prev_eq,
                                           unsigned& m_ta, unsigned& m_eq, unsigned& m_emit) {
    auto ka = A(ai);
    auto kb = B(bi);
    m_ta = m_eq = m_emit = 0;
#pragma unroll
    for (int k = 0; k < kVT; ++k) {
        const bool ta = ka <= kb;
        const bool eq = ka == kb;
        m_ta |= (unsigned)ta << k;
        m_eq |= (unsigned)eq << k;
        if (KIND == KIND_UNION) m_emit |= (unsigned)(ta || !prev_eq) << k;
        prev_eq = eq;                     // eq implies ta
        ai += ta ? 1 : 0;
        bi += ta ? 0 : 1;
        const auto kn = ta ? A(ai) : B(bi);
        ka = ta ? kn : ka;
        kb = ta ? kb : kn;
    }
    const unsigned live = steps >= kVT ? (1u << kVT) - 1u : (1u << steps) - 1u;
    m_ta &= live;
    m_eq &= live;
    m_emit = (KIND == KIND_UNION ? m_emit : m_eq) & live;
}


__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the merging warps only (unions have a ninth helper warp that must not be waited for)
template <int KIND>
__device__ __forceinline__ void sync_mergers() {
    if constexpr (KIND == 0) asm volatile("bar.sync 1, %0;" ::"n"(SPB_UNION_NT) : "memory");
    else __syncthreads();
}

// One CTA = one tile (blockIdx order, so the decoupled look-back only ever waits on tiles that
// started earlier).  Latency is hidden across the CTAs resident per SM (8 for intersections, 5-6 for
// unions): while one waits for its TMA copies, others merge or store.
template <typename V, typename VO, int KIND, bool HAS_VAL, bool SEAM>
__global__ void __launch_bounds__(Cfg<V, VO, KIND, HAS_VAL, SEAM>::kThreads, Cfg<V, VO, KIND, HAS_VAL, SEAM>::kMinCtas)
setop_kernel(const SetopParams p) {
    using C = Cfg<V, VO, KIND, HAS_VAL, SEAM>;
    using L = typename C::L;
    constexpr int kVT = C::kVT;
    constexpr int kFixedSlots = C::kFixedSlots;
    constexpr int kHelper = C::kHelperWarp;
    extern __shared__ __align__(128) char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    Geom* geom = reinterpret_cast<Geom*>(smem + 16);                    // 40 bytes
    int* warp_sums = reinterpret_cast<int*>(smem + 64);                 // [8]
    int64_t* base_s = reinterpret_cast<int64_t*>(smem + 96);
    char* buf = smem + L::kHeader;
    char* stg = buf + L::kBuf;
    uint32_t* skey = reinterpret_cast<uint32_t*>(stg);                  // compacted outputs (keys minus the tile base)
    VO* sval = reinterpret_cast<VO*>(stg + L::kStgKeys);
    uint8_t* slut = reinterpret_cast<uint8_t*>(stg + L::kStgKeys);
    int64_t* spa = reinterpret_cast<int64_t*>(stg + L::kStgKeys);
    int64_t* spb_ = spa + C::kCap;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    cudaGridDependencySynchronize();               // descriptors come from the previous launch

    uint64_t* agg_bar = reinterpret_cast<uint64_t*>(smem + 104);       // unions: count published -> helper
    uint64_t* ready_bar = reinterpret_cast<uint64_t*>(smem + 112);     // unions: offset resolved -> mergers
    int* tot_s = reinterpret_cast<int*>(smem + 120);
    if (tid == 0) {
        mbar_init(full, KIND == KIND_UNION ? 2 : 1);   // unions: + the helper's "edges written" arrival
        if (KIND == KIND_UNION) {
            mbar_init(agg_bar, 1);
            mbar_init(ready_bar, 1);
        }
        fence_mbar_init();
    }
    // unions: every warp needs the initialised barriers before the helper (another warp than thread 0's)
    // may use them.  Intersections: thread 0 sits in the staging warp itself, so a warp-level sync is
    // enough there and the others learn about the barrier through (A).
    if constexpr (KIND == KIND_UNION) __syncthreads();
    else __syncwarp();

    if (warp == kHelper) {
        // ---- stage the tile: descriptor -> TMA bulk copies, ragged edges, sentinels -------------
        const uint64_t dw = lane < kDescWords ? __ldg(p.desc + tile * kDescWords + lane) : 0;
        EdgeReg er;
        prefetch_tile(dw, buf, full, geom, er, lane);
        edge_store(er);
        const int64_t a0 = (int64_t)__shfl_sync(0xffffffffu, dw, 0), b0 = (int64_t)__shfl_sync(0xffffffffu, dw, 1);
        const uint64_t nn = __shfl_sync(0xffffffffu, dw, 2), oo = __shfl_sync(0xffffffffu, dw, 3);
        const int an_ = (int)(uint32_t)nn, bn_ = (int)(uint32_t)(nn >> 32);
        int64_t* As64w = reinterpret_cast<int64_t*>(buf + (int)(uint32_t)oo);
        int64_t* Bs64w = reinterpret_cast<int64_t*>(buf + (int)(uint32_t)(oo >> 32));
        if (lane == 0 && a0 == 0) As64w[-1] = kKeyMin;
        if (lane == 1 && a0 + an_ == p.na) As64w[an_] = kKeyMax;
        if (lane == 2 && b0 == 0) Bs64w[-1] = kKeyMin;
        if (lane == 3 && b0 + bn_ == p.nb) Bs64w[bn_] = kKeyMax;
        if constexpr (KIND == KIND_UNION) {
            __syncwarp();
            if (lane == 0) mbar_arrive(full);                          // geometry, edges, sentinels are in place
            // The exclusive prefix of the EARLIER tiles does not depend on this tile's own count, so the
            // look-back starts now and overlaps the TMA flight and the whole merge.
            uint64_t b = 0;
            if (tile > 0) b = lookback_exclusive(p.state, tile, p.epoch);
            mbar_wait(agg_bar, 0);                                     // the mergers have published the count
            const int total = *tot_s;
            if (lane == 0) {
                if (tile > 0) st_state(p.state + tile, pack_state(kStPrefix, p.epoch, b + (uint64_t)total));
                if (tile == (int64_t)gridDim.x - 1) *p.out_count = (int64_t)b + total;
                *base_s = (int64_t)b;
                mbar_arrive(ready_bar);
            }
            return;
        }
    }
    if constexpr (KIND != KIND_UNION) __syncthreads();                 // (A) geometry / edges visible
    mbar_wait(full, 0);

    const Geom g = *geom;
    const int an = g.an, bn = g.bn, items = an + bn;
    const int64_t* As64 = reinterpret_cast<const int64_t*>(buf + g.offA);
    const int64_t* Bs64 = reinterpret_cast<const int64_t*>(buf + g.offB);

    // ---- 32-bit tile-local keys when the span allows it ---------------------------------------
    const int64_t first_a = an ? As64[0] : kKeyMax, first_b = bn ? Bs64[0] : kKeyMax;
    const int64_t last_a = an ? As64[an - 1] : kKeyMin, last_b = bn ? Bs64[bn - 1] : kKeyMin;
    const int64_t kbase = first_a < first_b ? first_a : first_b;
    const int64_t kmax = last_a > last_b ? last_a : last_b;
    const bool narrow = (uint64_t)(kmax - kbase) < 0xfffffffeull;
    const bool front_dup = bn > 0 && As64[-1] == Bs64[0];              // A's last key of the previous tile pairs B's first
    if (narrow && tid < 2) {
        // the halo / sentinel key behind each range may lie beyond the 32-bit window: clamp it to the
        // largest offset, which no key of the tile can reach
        int64_t* h = const_cast<int64_t*>(tid == 0 ? As64 + an : Bs64 + bn);
        if ((uint64_t)(*h - kbase) > 0xfffffffeull) *h = kbase + 0xffffffffll;
    }
    sync_mergers<KIND>();                                              // (B)

    // ---- per-thread merge-path split + branch-free walk ------------------------------------
    int diag = tid * kVT;
    if (diag > items) diag = items;
    const int steps = (items - diag) < kVT ? (items - diag) : kVT;
    int ai0, bi0;
    unsigned m_ta, m_eq, m_emit;
    if (narrow) {
        const NarrowKeys A{reinterpret_cast<const uint32_t*>(As64), (uint32_t)kbase};
        const NarrowKeys B{reinterpret_cast<const uint32_t*>(Bs64), (uint32_t)kbase};
        ai0 = merge_path_smem(A, an, B, bn, diag);
        bi0 = diag - ai0;
        const bool prev_eq = ai0 > 0 ? (A(ai0 - 1) == B(bi0)) : (bi0 == 0 && front_dup);
        merge_walk<NarrowKeys, KIND, kVT>(A, B, ai0, bi0, steps, prev_eq, m_ta, m_eq, m_emit);
    } else {
        const WideKeys A{As64}, B{Bs64};
        ai0 = merge_path_smem(A, an, B, bn, diag);
        bi0 = diag - ai0;
        const bool prev_eq = ai0 > 0 ? (A(ai0 - 1) == B(bi0)) : (bi0 == 0 && front_dup);
        merge_walk<WideKeys, KIND, kVT>(A, B, ai0, bi0, steps, prev_eq, m_ta, m_eq, m_emit);
    }

    // intersections: the two values of this thread's FIRST match are requested now, so that the gather
    // latency overlaps the scan and its barrier (a thread rarely has a second match)
    [[maybe_unused]] V pre_x = V(0), pre_y = V(0);
    if constexpr (KIND == KIND_INTERSECT && HAS_VAL) {
        if (m_emit) {
            const int k = __ffs(m_emit) - 1;
            const int na_before = __popc(m_ta & ((1u << k) - 1u));
            pre_x = __ldg(reinterpret_cast<const V*>(p.a_val) + g.a0 + ai0 + na_before);
            pre_y = __ldg(reinterpret_cast<const V*>(p.b_val) + g.b0 + bi0 + k - na_before);
        }
    }

    // ---- block exclusive scan of the output counts -------------------------------------------
    const int cnt = __popc(m_emit);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31 && warp < C::kNTc / 32) warp_sums[warp] = incl;
    sync_mergers<KIND>();                                              // (C)
    int wofs = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < C::kNTc / 32; ++w) {
        const int s = warp_sums[w];
        if (w < warp) wofs += s;
        tile_total += s;
    }
    int off = wofs + incl - cnt;

    // ---- where do this tile's outputs go? ---------------------------------------------------------
    if constexpr (KIND == KIND_INTERSECT) {
        // Every tile owns kFixedSlots output slots outright (matches are rare: no round trip at all);
        // a tile with more matches reserves room in the overflow area with one atomic.
        // merge_reorder_kernel moves the chunks into tile order.
        if (tid == 0) {
            int64_t slot = tile * kFixedSlots;
            if (tile_total > kFixedSlots) {
                slot = p.ntiles * kFixedSlots + (int64_t)atomicAdd(p.slot_counter, (unsigned long long)tile_total);
                *base_s = slot;
            }
            p.tile_slot[tile] = slot;
            p.tile_cnt[tile] = tile_total;
            if (tile_total) atomicAdd(p.group_cnt + tile / kGroupTiles, tile_total);
        }
    } else if (tid == 0) {
        // publish the count for the successors, and hand it to the helper warp (which has been looking
        // back since the tile was put in flight)
        st_state(p.state + tile, pack_state(tile == 0 ? kStPrefix : kStAgg, p.epoch, (uint64_t)tile_total));
        *tot_s = tile_total;
        mbar_arrive(agg_bar);
    }

    // ---- materialise the emitted items: unions into the shared staging area (stored coalesced
    //      below); intersections -- few outputs -- straight to their reserved global slots -----------
    int64_t* ek = nullptr;
    VO* ev = sval;
    int64_t *epa = spa, *epb = spb_;
    if constexpr (KIND == KIND_INTERSECT) {
        int64_t slot = tile * kFixedSlots;                             // the tile's own slots: no round trip, no barrier
        if (tile_total > kFixedSlots) {                                // (uniform) overflow: thread 0 reserved room
            __syncthreads();
            slot = *base_s;
        }
        ek = p.tmp_idx + slot;
        ev = reinterpret_cast<VO*>(p.tmp_val) + slot;
        epa = p.tmp_apos + slot;
        epb = p.tmp_bpos + slot;
    }
    if constexpr (KIND == KIND_INTERSECT) {
        // matches are rare: visit only the set bits (the item's position in A / B follows from the
        // took-A mask, no walk needed)
        unsigned m = m_emit;
        bool first = true;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const int na_before = __popc(m_ta & ((1u << k) - 1u));
            const int ai = ai0 + na_before, bi = bi0 + k - na_before;
            ek[off] = As64[ai];
            if constexpr (HAS_VAL) {
                // only matched values are read (the first pair was requested before the scan)
                const V x = first ? pre_x : __ldg(reinterpret_cast<const V*>(p.a_val) + g.a0 + ai);
                V y = first ? pre_y : __ldg(reinterpret_cast<const V*>(p.b_val) + g.b0 + bi);
                first = false;
                if (p.negate_b) y = (V)(-y);
                ev[off] = value_op<V, VO>(p.op, x, y);
            }
            if constexpr (SEAM) {
                epa[off] = g.a0 + ai;
                epb[off] = g.b0 + bi;
            }
            ++off;
        }
        return;
    } else {
        // every input value is used exactly once: read it where it is needed, straight from global
        // memory (a thread's items are contiguous runs of A and of B, neighbours share the cache lines)
        const V* const Av = reinterpret_cast<const V*>(p.a_val) + g.a0;
        const V* const Bv = reinterpret_cast<const V*>(p.b_val) + g.b0;
        // one emission loop, two sinks: the shared staging area (32-bit keys; the normal case) or, for a
        // tile whose keys span >= 2^32, global memory directly once the tile's offset is known
        auto emit = [&](auto put_key, VO* out_v, uint8_t* out_lut) {
            int ai = ai0, bi = bi0;
#pragma unroll
            for (int k = 0; k < kVT; ++k) {
                const bool ta = (m_ta >> k) & 1u, e = (m_emit >> k) & 1u;
                [[maybe_unused]] const bool eq = ta && ((m_eq >> k) & 1u);
                if (e) {
                    put_key(off, ta, ai, bi);
                    if constexpr (HAS_VAL) {
                        const V x = ta ? __ldg(Av + ai) : V(0);
                        V y = (!ta || eq) ? __ldg(Bv + bi) : V(0);
                        if (p.negate_b && (!ta || eq)) y = (V)(-y);
                        if constexpr (sizeof(VO) == sizeof(V)) {
                            // assignment semantics: the stored b wins, a survives where b is absent
                            out_v[off] = p.op == SPB_OVERRIDE ? (VO)((!ta || eq) ? y : x) : value_op<V, VO>(p.op, x, y);
                        } else {
                            out_v[off] = value_op<V, VO>(p.op, x, y);
                        }
                    }
                    if constexpr (SEAM) {
                        if (out_lut) out_lut[off] = ta ? (eq ? 2 : 0) : 1;
                    }
                    ++off;
                }
                if constexpr (SEAM) {
                    if (k < steps) {
                        if (ta) {
                            if (p.a_only) p.a_only[g.a0 + ai] = eq ? 0 : 1;
                        } else {
                            if (p.b_only) p.b_only[g.b0 + bi] = e ? 1 : 0;
                        }
                    }
                }
                ai += ta ? 1 : 0;
                bi += ta ? 0 : ((k < steps) ? 1 : 0);
            }
        };
        VO* const o_val = reinterpret_cast<VO*>(p.out_val);
        if (narrow) {
            const NarrowKeys A{reinterpret_cast<const uint32_t*>(As64), (uint32_t)kbase};
            const NarrowKeys B{reinterpret_cast<const uint32_t*>(Bs64), (uint32_t)kbase};
            emit([&](int o, bool ta, int ai, int bi) { skey[o] = ta ? A(ai) : B(bi); }, sval, slut);
            sync_mergers<KIND>();                                      // (D) staging complete
            mbar_wait(ready_bar, 0);                                   // offset resolved by the helper warp
            const int64_t base = *base_s;
            for (int i = tid; i < tile_total; i += C::kNTc) {
                p.out_idx[base + i] = kbase + (int64_t)skey[i];
                if constexpr (HAS_VAL) o_val[base + i] = sval[i];
                if constexpr (SEAM) {
                    if (p.lut) p.lut[base + i] = slut[i];
                }
            }
        } else {
            mbar_wait(ready_bar, 0);
            const int64_t base = *base_s;
            int64_t* const ok = p.out_idx + base;
            emit([&](int o, bool ta, int ai, int bi) { ok[o] = ta ? As64[ai] : Bs64[bi]; }, o_val + base,
                 p.lut ? p.lut + base : nullptr);
        }
    }
}

// ---- intersections: put the reserved chunks in tile order -----------------------------------------
// One warp per tile.  The output offset of a tile is the sum of the match counts of all earlier tiles:
// the merge kernel also added every tile's count into one counter per group of 64 tiles, so a warp sums
// (at most ntiles/64) group counters and (at most 63) tile counts -- a single round of independent L2
// loads, no scan launch, no look-back.
template <int VBYTES>
__global__ void __launch_bounds__(256) merge_reorder_kernel(const SetopParams p) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t = (int64_t)blockIdx.x * 8 + warp;
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();               // chunks and counts come from the previous launch
    if (blockIdx.x == 0 && tid == 0) *p.slot_counter = 0;   // every reservation is done: ready for the next call
    if (t >= p.ntiles) return;
    const int64_t g = t / kGroupTiles;
    int64_t acc = 0;
    for (int64_t j = lane; j < g; j += 32) acc += __ldg(p.group_cnt + j);
    for (int64_t j = g * kGroupTiles + lane; j < t; j += 32) acc += __ldg(p.tile_cnt + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const int64_t d = acc;
    const int n = p.tile_cnt[t];
    if (t == p.ntiles - 1 && lane == 0) *p.out_count = d + n;
    const int64_t src = p.tile_slot[t];
    for (int i = lane; i < n; i += 32) {
        p.out_idx[d + i] = p.tmp_idx[src + i];
        if (VBYTES == 1) reinterpret_cast<uint8_t*>(p.out_val)[d + i] = reinterpret_cast<const uint8_t*>(p.tmp_val)[src + i];
        if (VBYTES == 4) reinterpret_cast<uint32_t*>(p.out_val)[d + i] = reinterpret_cast<const uint32_t*>(p.tmp_val)[src + i];
        if (VBYTES == 8) reinterpret_cast<uint64_t*>(p.out_val)[d + i] = reinterpret_cast<const uint64_t*>(p.tmp_val)[src + i];
        if (p.a_pos) p.a_pos[d + i] = p.tmp_apos[src + i];
        if (p.b_pos) p.b_pos[d + i] = p.tmp_bpos[src + i];
    }
}

template <typename V, typename VO, int KIND, bool HAS_VAL, bool SEAM>
int launch_setop(SetopParams& p, spb_workspace* ws, cudaStream_t stream) {
    constexpr int smem = Cfg<V, VO, KIND, HAS_VAL, SEAM>::L::kTotal;
    const int64_t total = p.na + p.nb;
    if (total == 0 || (KIND == KIND_INTERSECT && (p.na == 0 || p.nb == 0))) {
        return (int)cudaMemsetAsync(p.out_count, 0, sizeof(int64_t), stream);
    }
    using C = Cfg<V, VO, KIND, HAS_VAL, SEAM>;
    constexpr int kTile = C::kTile;
    constexpr int kFixedSlots = C::kFixedSlots;
    const int64_t ntiles = ceil_div(total, kTile);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)ntiles * sizeof(uint64_t), stream);
    if (rc) return rc;
    // payload: descriptors | splits | (intersections) tile_slot, tile_dst, tile_cnt, counter, chunked outputs
    const size_t cap = (size_t)(p.na < p.nb ? p.na : p.nb) + (size_t)ntiles * kFixedSlots;
    size_t off_splits = (size_t)ntiles * kDescWords * 8;
    size_t off_slot = off_splits + (size_t)(ntiles + 1) * 8;
    size_t off_dst = off_slot + (size_t)ntiles * 8;
    size_t off_cnt = off_dst + (size_t)ntiles * 8;
    const int64_t ngroups = ceil_div(ntiles, kGroupTiles);
    size_t off_group = (off_cnt + (size_t)ntiles * 4 + 15) & ~(size_t)15;
    size_t off_counter = (off_group + (size_t)ngroups * 4 + 15) & ~(size_t)15;
    size_t off_tidx = off_counter + 16;
    size_t off_tval = off_tidx + cap * 8;
    size_t off_tapos = off_tval + cap * 8;
    size_t off_tbpos = off_tapos + (SEAM ? cap * 8 : 0);
    size_t payload_bytes = KIND == KIND_INTERSECT ? off_tbpos + (SEAM ? cap * 8 : 0) : off_slot;
    rc = workspace_reserve_payload(ws, payload_bytes, stream);
    if (rc) return rc;
    rc = workspace_next_epoch(ws, stream, &p.epoch, KIND == KIND_UNION);   // intersections use no look-back words
    if (rc) return rc;
    p.state = ws_state(ws);
    p.ntiles = ntiles;
    p.desc = reinterpret_cast<uint64_t*>(ws->payload);
    auto kern = setop_kernel<V, VO, KIND, HAS_VAL, SEAM>;
    static PerDevice configured;       // function attributes are per device
    if (configured.first_use(current_device())) {
        SPB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        SPB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    PartitionParams pp{};
    pp.a = p.a; pp.b = p.b; pp.na = p.na; pp.nb = p.nb;
    pp.ntiles = ntiles;
    pp.frac_a = (double)p.na / (double)total;
    pp.tile_items = kTile;
    pp.desc = reinterpret_cast<uint64_t*>(ws->payload);
    char* pay = reinterpret_cast<char*>(ws->payload);
    if (KIND == KIND_INTERSECT) {
        p.tile_slot = reinterpret_cast<int64_t*>(pay + off_slot);
        p.tile_cnt = reinterpret_cast<int*>(pay + off_cnt);
        p.group_cnt = reinterpret_cast<int*>(pay + off_group);
        pp.group_cnt = p.group_cnt;
        pp.ngroups = ngroups;
        p.slot_counter = ws_counters(ws);          // zero between calls: the reorder pass resets it
        p.tmp_idx = reinterpret_cast<int64_t*>(pay + off_tidx);
        p.tmp_val = pay + off_tval;
        p.tmp_apos = reinterpret_cast<int64_t*>(pay + off_tapos);
        p.tmp_bpos = reinterpret_cast<int64_t*>(pay + off_tbpos);
    }
    SPB_CUDA_TRY(launch_pdl(merge_partition_kernel, dim3((unsigned)ceil_div(ntiles, kPartTiles)), dim3(kPartThreads), 0, stream, pp));
    SPB_LAUNCHED(1);
    SPB_CUDA_TRY(launch_pdl(kern, dim3((unsigned)ntiles), dim3(C::kThreads), (size_t)smem, stream, p));
    SPB_LAUNCHED(1);
    if (KIND == KIND_INTERSECT) {
        const unsigned rblocks = (unsigned)ceil_div(ntiles, 8);
        constexpr int VBY = HAS_VAL ? (int)sizeof(VO) : 0;
        SPB_CUDA_TRY(launch_pdl(merge_reorder_kernel<VBY>, dim3(rblocks), dim3(256), 0, stream, p));
        SPB_LAUNCHED(1);
    }
    return (int)cudaGetLastError();
}

template <int KIND>
int dispatch_binop(int op, int dtype, SetopParams& p, spb_workspace* ws, cudaStream_t stream) {
    if (op < 0 || op > SPB_OVERRIDE) return SPB_ERR_BAD_ARG;        // divide ops exist for spb_binary_map only
    if (op == SPB_OVERRIDE && KIND != KIND_UNION) return SPB_ERR_BAD_ARG;
    const bool cmp = op >= SPB_LT && op <= SPB_NE;
    p.op = op == SPB_SUB ? SPB_ADD : op;
    p.negate_b = op == SPB_SUB;
#define SPB_CASE(DT, V)                                                                  \
    case DT:                                                                             \
        return cmp ? launch_setop<V, uint8_t, KIND, true, false>(p, ws, stream)          \
                   : launch_setop<V, V, KIND, true, false>(p, ws, stream);
    switch (dtype) {
        SPB_CASE(SPB_F32, float)
        SPB_CASE(SPB_F64, double)
        SPB_CASE(SPB_I32, int32_t)
        SPB_CASE(SPB_I64, int64_t)
        default: return SPB_ERR_UNSUPPORTED;   // bool / 8-bit operands are widened by the host
    }
#undef SPB_CASE
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_union1d_sorted(const int64_t* a, int64_t na, const int64_t* b, int64_t nb, int64_t* out, uint8_t* lut,
                       uint8_t* a_only, uint8_t* b_only, int64_t* out_count, spb_workspace* ws,
                       spb_stream_t stream) {
    if (na < 0 || nb < 0 || !out_count || (na + nb > 0 && !out) || (na && !a) || (nb && !b)) return SPB_ERR_BAD_ARG;
    SetopParams p{};
    p.a = a; p.b = b; p.na = na; p.nb = nb;
    p.out_idx = out; p.lut = lut; p.a_only = a_only; p.b_only = b_only; p.out_count = out_count;
    if (lut || a_only || b_only) return launch_setop<int64_t, int64_t, KIND_UNION, false, true>(p, ws, stream);
    return launch_setop<int64_t, int64_t, KIND_UNION, false, false>(p, ws, stream);
}

int spb_intersect1d_sorted(const int64_t* a, int64_t na, const int64_t* b, int64_t nb, int64_t* out,
                           int64_t* a_inds, int64_t* b_inds, int64_t* out_count, spb_workspace* ws,
                           spb_stream_t stream) {
    if (na < 0 || nb < 0 || !out_count || (na && nb && !out) || (na && !a) || (nb && !b)) return SPB_ERR_BAD_ARG;
    SetopParams p{};
    p.a = a; p.b = b; p.na = na; p.nb = nb;
    p.out_idx = out; p.a_pos = a_inds; p.b_pos = b_inds; p.out_count = out_count;
    if (a_inds || b_inds) return launch_setop<int64_t, int64_t, KIND_INTERSECT, false, true>(p, ws, stream);
    return launch_setop<int64_t, int64_t, KIND_INTERSECT, false, false>(p, ws, stream);
}

int spb_union_binop(int op, int dtype, const int64_t* a_idx, const void* a_val, int64_t na, const int64_t* b_idx,
                    const void* b_val, int64_t nb, int64_t* out_idx, void* out_val, int64_t* out_count,
                    spb_workspace* ws, spb_stream_t stream) {
    if (na < 0 || nb < 0 || !out_count || (na + nb > 0 && (!out_idx || !out_val))) return SPB_ERR_BAD_ARG;
    if ((na && (!a_idx || !a_val)) || (nb && (!b_idx || !b_val))) return SPB_ERR_BAD_ARG;
    SetopParams p{};
    p.a = a_idx; p.b = b_idx; p.na = na; p.nb = nb; p.a_val = a_val; p.b_val = b_val;
    p.out_idx = out_idx; p.out_val = out_val; p.out_count = out_count;
    return dispatch_binop<KIND_UNION>(op, dtype, p, ws, stream);
}

int spb_intersect_binop(int op, int dtype, const int64_t* a_idx, const void* a_val, int64_t na,
                        const int64_t* b_idx, const void* b_val, int64_t nb, int64_t* out_idx, void* out_val,
                        int64_t* out_count, spb_workspace* ws, spb_stream_t stream) {
    if (na < 0 || nb < 0 || !out_count || (na && nb && (!out_idx || !out_val))) return SPB_ERR_BAD_ARG;
    if ((na && (!a_idx || !a_val)) || (nb && (!b_idx || !b_val))) return SPB_ERR_BAD_ARG;
    SetopParams p{};
    p.a = a_idx; p.b = b_idx; p.na = na; p.nb = nb; p.a_val = a_val; p.b_val = b_val;
    p.out_idx = out_idx; p.out_val = out_val; p.out_count = out_count;
    return dispatch_binop<KIND_INTERSECT>(op, dtype, p, ws, stream);
}

}  // extern "C"
