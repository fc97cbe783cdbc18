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
// Two launches per call:
//   1. merge_partition_kernel: one warp per tile boundary finds the merge-path split of
//      its diagonal on the two global arrays (32-ary cooperative search, ~5 round trips).
//   2. setop_kernel: persistent CTAs (a few per SM) walk the tiles in stride.  Per tile
//        - warp 0 PREFETCHES the next tile while the current one is merged: TMA bulk
//          copies (cp.async.bulk + mbarrier) of the A / B index ranges and value ranges
//          into the other shared-memory buffer; the < 16-byte unaligned heads / tails
//          and the halo keys ride in registers of that warp until the buffer is used;
//        - the staged 64-bit keys are re-based to 32-bit tile-local offsets whenever the
//          tile's key span allows it (64-bit path otherwise), so the merge runs on
//          single-register keys;
//        - every thread merge-path-splits the tile in shared memory and walks kVT items
//          with a branch-free step that only records three bit masks (took-A, matched,
//          emitted); a key present in both inputs is emitted once, by its A element
//          (the B partner is found through the halo key even across a tile boundary);
//        - block scan of the output counts, decoupled look-back across tiles, then the
//          emitted items are materialised (value op applied; for intersections the two
//          values are gathered straight from global memory, so unmatched values are
//          never read), compacted through shared memory and stored coalesced.
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
    const int64_t* splits;   // [ntiles + 1] merge-path split (A index) of every tile boundary
    int64_t ntiles;
    uint32_t epoch;
    int op;
    int negate_b;        // a - b is evaluated as a + (-b), base.py:45-46: a missing partner is +0, not -0
};

enum { KIND_UNION = 0, KIND_INTERSECT = 1 };

constexpr int kNT = 256;   // threads per CTA
constexpr int kVT = 8;     // merged items per thread
constexpr int kTile = kNT * kVT;

// ---- launch 1: merge-path split of every tile boundary ---------------------------------------
__global__ void __launch_bounds__(256) merge_partition_kernel(const int64_t* __restrict__ a, int64_t na,
                                                              const int64_t* __restrict__ b, int64_t nb,
                                                              int64_t ntiles, int64_t* __restrict__ splits) {
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (w > ntiles) return;
    const int64_t total = na + nb;
    int64_t d = w * kTile;
    if (d > total) d = total;
    const int64_t s = merge_path_warp(a, na, b, nb, d);
    if ((threadIdx.x & 31) == 0) splits[w] = s;
}

// ---- shared-memory layout ------------------------------------------------------------------------
struct Geom {            // where a staged tile lives inside its buffer
    int64_t a0, b0;
    int an, bn;
    int offA, offB;      // byte offset of As64[0] / Bs64[0] from the buffer start
    int offAv, offBv;    // byte offset of Av[0] / Bv[0]
};

template <int VBYTES, int OUT_KEYS, int OUT_PAYLOAD>   // staged value bytes; staging capacity in bytes
struct Layout {
    static constexpr int kHeader = 256;
    static constexpr int kKeyReg = (kTile + 4) * 8 + 128 + kVT * 8;
    static constexpr int kValReg = VBYTES ? kTile * VBYTES + 96 : 0;
    static constexpr int kBuf = kKeyReg + kValReg;                       // the staged input tile
    static constexpr int kK32 = ((kTile + 2 * kVT + 24) * 4 + 15) & ~15; // re-based 32-bit keys
    static constexpr int kStgKeys = OUT_KEYS;                            // compacted outputs waiting for their offset
    static constexpr int kStg = ((OUT_KEYS + OUT_PAYLOAD) + 15) & ~15;
    static constexpr int kTotal = kHeader + kBuf + kK32 + kStg;
    static_assert(kKeyReg % 16 == 0 && kValReg % 16 == 0, "buffers must stay 16-byte aligned for TMA");
};

template <typename V, typename VO, int KIND, bool HAS_VAL, bool SEAM>
struct Cfg {
    static constexpr bool kStageVals = HAS_VAL && KIND == KIND_UNION;
    static constexpr int kCap = KIND == KIND_UNION ? kTile : kTile / 2;  // outputs per tile, at most
    static constexpr int kPayload = HAS_VAL ? kCap * (int)sizeof(VO)
                                            : (SEAM ? (KIND == KIND_UNION ? kCap : 2 * kCap * 8) : 0);
    using L = Layout<kStageVals ? (int)sizeof(V) : 0, kCap * 8, kPayload>;
};

struct EdgeReg {         // an unaligned head / tail element riding in a register until its buffer is live
    uint64_t raw;
    uint32_t dst;        // shared-memory address, 0 = nothing
    uint32_t size;
};

__device__ __forceinline__ void edge_load(EdgeReg& er, const StagePlan& s, char* buf, int esize, bool tail, int j) {
    const uintptr_t head_end = s.q_lo < s.p_hi ? s.q_lo : s.p_hi;
    uintptr_t lo = s.p_lo, hi = head_end;
    if (tail) {
        lo = s.q_hi > head_end ? s.q_hi : head_end;
        hi = s.p_hi;
    }
    const uintptr_t addr = lo + (uintptr_t)j * esize;
    if (addr < hi) {
        er.dst = smem_u32(buf + (addr - s.g_al));
        er.size = (uint32_t)esize;
        if (esize == 8) er.raw = __ldg(reinterpret_cast<const unsigned long long*>(addr));
        else er.raw = __ldg(reinterpret_cast<const unsigned int*>(addr));
    }
}

__device__ __forceinline__ void edge_store(const EdgeReg& er) {
    if (er.dst) {
        if (er.size == 8) asm volatile("st.shared.u64 [%0], %1;" ::"r"(er.dst), "l"(er.raw) : "memory");
        else asm volatile("st.shared.u32 [%0], %1;" ::"r"(er.dst), "r"((uint32_t)er.raw) : "memory");
    }
}

// Executed by the whole of warp 0: stage tile `tile` into `buf` (TMA for the aligned middle,
// registers for the ragged edges), describe it in *g.
template <typename V, bool STAGE_VALS>
__device__ __forceinline__ void prefetch_tile(const SetopParams& p, int64_t tile, char* buf, uint64_t* bar, Geom* g,
                                              EdgeReg& er, int lane) {
    constexpr int kKeyReg = (kTile + 4) * 8 + 128 + kVT * 8;
    const int64_t total = p.na + p.nb;
    const int64_t a0 = __ldg(p.splits + tile), a1 = __ldg(p.splits + tile + 1);
    const int64_t d0 = tile * kTile;
    const int64_t d1 = (d0 + kTile < total) ? d0 + kTile : total;
    const int64_t b0 = d0 - a0, b1 = d1 - a1;
    // indices [x0-1, x1] (one halo key either side), values [x0, x1) (+1 for B: partner of the last A)
    const int64_t ga_lo = a0 > 0 ? a0 - 1 : 0, ga_hi = a1 < p.na ? a1 + 1 : p.na;
    const int64_t gb_lo = b0 > 0 ? b0 - 1 : 0, gb_hi = b1 < p.nb ? b1 + 1 : p.nb;
    const StagePlan pa = make_stage_plan(p.a, ga_lo, ga_hi, 8);
    const StagePlan pb = make_stage_plan(p.b, gb_lo, gb_hi, 8);
    char* bufA = buf + 16;                                            // front pad: slot for index -1
    char* bufB = bufA + ((pa.span_bytes() + 15) & ~15u) + 32;         // A's +1 sentinel, B's -1 slot
    StagePlan pav{}, pbv{};
    char *bufAv = buf + kKeyReg, *bufBv = nullptr;
    if constexpr (STAGE_VALS) {
        pav = make_stage_plan(p.a_val, a0, a1, sizeof(V));
        pbv = make_stage_plan(p.b_val, b0, gb_hi > b1 ? b1 + 1 : b1, sizeof(V));
        bufBv = bufAv + ((pav.span_bytes() + 15) & ~15u) + 16;
    }
    if (lane == 0) {
        g->a0 = a0;
        g->b0 = b0;
        g->an = (int)(a1 - a0);
        g->bn = (int)(b1 - b0);
        g->offA = (int)(bufA - buf) + (int)(pa.p_lo - pa.g_al) + (int)(a0 - ga_lo) * 8;
        g->offB = (int)(bufB - buf) + (int)(pb.p_lo - pb.g_al) + (int)(b0 - gb_lo) * 8;
        uint32_t bytes = pa.tma_bytes() + pb.tma_bytes();
        if constexpr (STAGE_VALS) {
            g->offAv = (int)(bufAv - buf) + (int)(pav.p_lo - pav.g_al);
            g->offBv = (int)(bufBv - buf) + (int)(pbv.p_lo - pbv.g_al);
            bytes += pav.tma_bytes() + pbv.tma_bytes();
        }
        mbar_arrive_expect_tx(bar, bytes);
        stage_issue_tma(pa, bufA, bar);
        stage_issue_tma(pb, bufB, bar);
        if constexpr (STAGE_VALS) {
            stage_issue_tma(pav, bufAv, bar);
            stage_issue_tma(pbv, bufBv, bar);
        }
    }
    er.dst = 0;
    er.size = 0;
    er.raw = 0;
    if (lane == 1) edge_load(er, pa, bufA, 8, false, 0);
    else if (lane == 2) edge_load(er, pa, bufA, 8, true, 0);
    else if (lane == 3) edge_load(er, pb, bufB, 8, false, 0);
    else if (lane == 4) edge_load(er, pb, bufB, 8, true, 0);
    if constexpr (STAGE_VALS) {
        constexpr int E = (int)sizeof(V);
        if (lane >= 5 && lane <= 7) edge_load(er, pav, bufAv, E, false, lane - 5);
        else if (lane >= 8 && lane <= 10) edge_load(er, pav, bufAv, E, true, lane - 8);
        else if (lane >= 11 && lane <= 13) edge_load(er, pbv, bufBv, E, false, lane - 11);
        else if (lane >= 14 && lane <= 16) edge_load(er, pbv, bufBv, E, true, lane - 14);
    }
}

template <typename K>
__device__ __forceinline__ int merge_path_smem(const K* a, int na, const K* b, int nb, int d) {
    int lo = d > nb ? d - nb : 0;
    int hi = d < na ? d : na;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= b[d - 1 - mid]) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Branch-free walk of kVT merged items starting at (ai, bi): records who was taken (A or B),
// where both sides held the same key, and which items are emitted.  Ties go to A; the B
// partner of a matched A item is the very next item and is suppressed (prev_eq).
template <typename K, int KIND>
__device__ __forceinline__ void merge_walk(const K* KA, const K* KB, int ai, int bi, int steps, bool prev_eq,
                                           unsigned& m_ta, unsigned& m_eq, unsigned& m_emit) {
    K ka = KA[ai], kb = KB[bi];
    m_ta = m_eq = m_emit = 0;
#pragma unroll
    for (int k = 0; k < kVT; ++k) {
        const bool ta = ka <= kb;
        const bool eq = ka == kb;
        const bool e = KIND == KIND_UNION ? (ta || !prev_eq) : eq;
        prev_eq = ta && eq;
        const bool live = k < steps;
        m_ta |= (unsigned)(ta && live) << k;
        m_eq |= (unsigned)(eq && ta && live) << k;
        m_emit |= (unsigned)(e && (ta || KIND == KIND_UNION) && live) << k;
        ai += ta ? 1 : 0;
        bi += ta ? 0 : 1;
        const K kn = ta ? KA[ai] : KB[bi];
        ka = ta ? kn : ka;
        kb = ta ? kb : kn;
    }
}

template <typename V, typename VO, int KIND, bool HAS_VAL, bool SEAM>
__global__ void __launch_bounds__(kNT, 3) setop_kernel(const SetopParams p) {
    using C = Cfg<V, VO, KIND, HAS_VAL, SEAM>;
    using L = typename C::L;
    constexpr bool STAGE_VALS = C::kStageVals;
    extern __shared__ __align__(128) char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);                 // mbarrier of the input buffer
    Geom* geom = reinterpret_cast<Geom*>(smem + 16);
    int* warp_sums = reinterpret_cast<int*>(smem + 112);                // [8]
    int* lb_scratch = reinterpret_cast<int*>(smem + 144);               // [8] ints + [8] u64
    char* buf = smem + L::kHeader;
    uint32_t* k32 = reinterpret_cast<uint32_t*>(buf + L::kBuf);
    char* stg = buf + L::kBuf + L::kK32;
    // compacted outputs of the PREVIOUS tile wait here for their global offset
    int64_t* skey = reinterpret_cast<int64_t*>(stg);
    VO* sval = reinterpret_cast<VO*>(stg + L::kStgKeys);
    uint8_t* slut = reinterpret_cast<uint8_t*>(stg + L::kStgKeys);
    int64_t* spa = reinterpret_cast<int64_t*>(stg + L::kStgKeys);
    int64_t* spb_ = spa + C::kCap;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t total_items = p.na + p.nb;
    const int64_t stride = gridDim.x;

    if (tid == 0) {
        mbar_init(full, 1);
        fence_mbar_init();
    }
    __syncthreads();

    EdgeReg er{0, 0, 0};
    int64_t tile = blockIdx.x;
    if (warp == 0 && tile < p.ntiles) prefetch_tile<V, STAGE_VALS>(p, tile, buf, full, geom, er, lane);
    uint32_t phase = 0;
    int64_t prev_tile = -1;      // tile whose outputs sit in the staging area
    int prev_total = 0;

    // Store the staged outputs of `prev_tile`: its predecessors published their counts at least
    // one tile-iteration ago, so the look-back does not wait.
    auto flush_prev = [&]() {
        uint64_t base = 0;
        if (prev_tile > 0) {
            base = lookback_exclusive_block256(p.state, prev_tile, p.epoch, lb_scratch);
            if (tid == 0) st_state(p.state + prev_tile, pack_state(kStPrefix, p.epoch, base + (uint64_t)prev_total));
        }
        if (tid == 0 && prev_tile == p.ntiles - 1) *p.out_count = (int64_t)base + prev_total;
        for (int i = tid; i < prev_total; i += kNT) {
            if constexpr (SEAM && KIND == KIND_INTERSECT) {
                p.out_idx[base + i] = skey[i];
                if (p.a_pos) p.a_pos[base + i] = spa[i];
                if (p.b_pos) p.b_pos[base + i] = spb_[i];
            } else {
                p.out_idx[base + i] = skey[i];
                if constexpr (HAS_VAL) reinterpret_cast<VO*>(p.out_val)[base + i] = sval[i];
                if constexpr (SEAM && KIND == KIND_UNION) {
                    if (p.lut) p.lut[base + i] = slut[i];
                }
            }
        }
    };

    for (; tile < p.ntiles; tile += stride) {
        if (warp == 0) {
            edge_store(er);                                            // ragged edges of the current tile
            __syncwarp();                                              // lane 0 wrote *geom when it prefetched
            int64_t* As64w = reinterpret_cast<int64_t*>(buf + geom->offA);
            int64_t* Bs64w = reinterpret_cast<int64_t*>(buf + geom->offB);
            if (lane == 17 && geom->a0 == 0) As64w[-1] = kKeyMin;
            if (lane == 18 && geom->a0 + geom->an == p.na) As64w[geom->an] = kKeyMax;
            if (lane == 19 && geom->b0 == 0) Bs64w[-1] = kKeyMin;
            if (lane == 20 && geom->b0 + geom->bn == p.nb) Bs64w[geom->bn] = kKeyMax;
        }
        mbar_wait(full, phase);
        phase ^= 1u;
        __syncthreads();                                               // (A) tile, edges, sentinels, geometry visible

        const Geom g = *geom;
        const int an = g.an, bn = g.bn, items = an + bn;
        const int64_t* As64 = reinterpret_cast<const int64_t*>(buf + g.offA);
        const int64_t* Bs64 = reinterpret_cast<const int64_t*>(buf + g.offB);

        // ---- re-base the keys to 32-bit tile-local offsets when the span allows it -----------
        const int64_t first_a = an ? As64[0] : kKeyMax, first_b = bn ? Bs64[0] : kKeyMax;
        const int64_t last_a = an ? As64[an - 1] : kKeyMin, last_b = bn ? Bs64[bn - 1] : kKeyMin;
        const int64_t kbase = first_a < first_b ? first_a : first_b;
        const int64_t kmax = last_a > last_b ? last_a : last_b;
        const bool narrow = (uint64_t)(kmax - kbase) < 0xfffffffeull;
        const bool front_dup = bn > 0 && As64[-1] == Bs64[0];          // A's last key of the previous tile pairs B's first
        uint32_t* KA32 = k32;
        uint32_t* KB32 = k32 + an + kVT + 4;
        if (narrow) {
            for (int e = tid; e <= an + kVT + 1; e += kNT) {
                uint32_t v = 0xffffffffu;
                if (e <= an) {
                    const uint64_t dlt = (uint64_t)(As64[e] - kbase);
                    v = dlt > 0xfffffffeull ? 0xffffffffu : (uint32_t)dlt;
                }
                KA32[e] = v;
            }
            for (int e = tid; e <= bn + kVT + 1; e += kNT) {
                uint32_t v = 0xffffffffu;
                if (e <= bn) {
                    const uint64_t dlt = (uint64_t)(Bs64[e] - kbase);
                    v = dlt > 0xfffffffeull ? 0xffffffffu : (uint32_t)dlt;
                }
                KB32[e] = v;
            }
        }
        __syncthreads();                                               // (B)

        // ---- per-thread merge-path split + branch-free walk ------------------------------------
        int diag = tid * kVT;
        if (diag > items) diag = items;
        const int steps = (items - diag) < kVT ? (items - diag) : kVT;
        int ai0, bi0;
        unsigned m_ta, m_eq, m_emit;
        if (narrow) {
            ai0 = merge_path_smem<uint32_t>(KA32, an, KB32, bn, diag);
            bi0 = diag - ai0;
            const bool prev_eq = ai0 > 0 ? (KA32[ai0 - 1] == KB32[bi0]) : (bi0 == 0 && front_dup);
            merge_walk<uint32_t, KIND>(KA32, KB32, ai0, bi0, steps, prev_eq, m_ta, m_eq, m_emit);
        } else {
            ai0 = merge_path_smem<int64_t>(As64, an, Bs64, bn, diag);
            bi0 = diag - ai0;
            const bool prev_eq = ai0 > 0 ? (As64[ai0 - 1] == Bs64[bi0]) : (bi0 == 0 && front_dup);
            merge_walk<int64_t, KIND>(As64, Bs64, ai0, bi0, steps, prev_eq, m_ta, m_eq, m_emit);
        }

        // ---- block exclusive scan of the output counts -------------------------------------------
        const int cnt = __popc(m_emit);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;

        // ---- materialise the emitted items in registers -------------------------------------------
        int64_t okey[kVT];
        VO oval[kVT];
        int64_t opa[SEAM && KIND == KIND_INTERSECT ? kVT : 1];
        int64_t opb[SEAM && KIND == KIND_INTERSECT ? kVT : 1];
        uint8_t olut[SEAM && KIND == KIND_UNION ? kVT : 1];
        if (KIND == KIND_UNION || m_emit != 0) {
            const V* Av = STAGE_VALS ? reinterpret_cast<const V*>(buf + g.offAv) : nullptr;
            const V* Bv = STAGE_VALS ? reinterpret_cast<const V*>(buf + g.offBv) : nullptr;
            int ai = ai0, bi = bi0;
#pragma unroll
            for (int k = 0; k < kVT; ++k) {
                const bool ta = (m_ta >> k) & 1u, eq = (m_eq >> k) & 1u, e = (m_emit >> k) & 1u;
                if (e) {
                    okey[k] = narrow ? kbase + (int64_t)(ta ? KA32[ai] : KB32[bi]) : (ta ? As64[ai] : Bs64[bi]);
                    if constexpr (HAS_VAL) {
                        V x, y;
                        if constexpr (STAGE_VALS) {
                            x = ta ? Av[ai] : V(0);
                            y = (!ta || eq) ? Bv[bi] : V(0);
                            if (p.negate_b && (!ta || eq)) y = (V)(-y);
                        } else {   // intersection: only matched values are ever read
                            x = __ldg(reinterpret_cast<const V*>(p.a_val) + g.a0 + ai);
                            y = __ldg(reinterpret_cast<const V*>(p.b_val) + g.b0 + bi);
                            if (p.negate_b) y = (V)(-y);
                        }
                        oval[k] = value_op<V, VO>(p.op, x, y);
                    }
                    if constexpr (SEAM && KIND == KIND_UNION) olut[k] = ta ? (eq ? 2 : 0) : 1;
                    if constexpr (SEAM && KIND == KIND_INTERSECT) {
                        opa[k] = g.a0 + ai;
                        opb[k] = g.b0 + bi;
                    }
                }
                if constexpr (SEAM && KIND == KIND_UNION) {
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
        }
        fence_proxy_async_smem();      // my generic reads of the input buffer happen-before the TMA that refills it
        __syncthreads();                                               // (C) warp sums ready; the input tile is dead

        // ---- the next tile goes in flight while this one's outputs are handled -------------------
        if (warp == 0 && tile + stride < p.ntiles)
            prefetch_tile<V, STAGE_VALS>(p, tile + stride, buf, full, geom, er, lane);

        int wofs = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kNT / 32; ++w) {
            const int s = warp_sums[w];
            if (w < warp) wofs += s;
            tile_total += s;
        }
        int off = wofs + incl - cnt;
        // publish this tile's count right away: successors look back for it a whole iteration later
        if (tid == 0)
            st_state(p.state + tile, pack_state(tile == 0 ? kStPrefix : kStAgg, p.epoch, (uint64_t)tile_total));

        // ---- store the previous tile (its look-back no longer waits), then stage this one -----------
        if (prev_tile >= 0) flush_prev();
        __syncthreads();                                               // (D) staging area free
#pragma unroll
        for (int k = 0; k < kVT; ++k) {
            if ((m_emit >> k) & 1u) {
                skey[off] = okey[k];
                if constexpr (HAS_VAL) sval[off] = oval[k];
                if constexpr (SEAM && KIND == KIND_UNION) slut[off] = olut[k];
                if constexpr (SEAM && KIND == KIND_INTERSECT) {
                    spa[off] = opa[k];
                    spb_[off] = opb[k];
                }
                ++off;
            }
        }
        prev_tile = tile;
        prev_total = tile_total;
    }
    __syncthreads();
    if (prev_tile >= 0) flush_prev();
}

template <typename V, typename VO, int KIND, bool HAS_VAL, bool SEAM>
int launch_setop(SetopParams& p, spb_workspace* ws, cudaStream_t stream) {
    constexpr int smem = Cfg<V, VO, KIND, HAS_VAL, SEAM>::L::kTotal;
    const int64_t total = p.na + p.nb;
    if (total == 0 || (KIND == KIND_INTERSECT && (p.na == 0 || p.nb == 0))) {
        return (int)cudaMemsetAsync(p.out_count, 0, sizeof(int64_t), stream);
    }
    const int64_t ntiles = ceil_div(total, kTile);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)ntiles * sizeof(uint64_t), stream);
    if (rc) return rc;
    rc = workspace_reserve_payload(ws, (size_t)(ntiles + 1) * sizeof(int64_t), stream);
    if (rc) return rc;
    rc = workspace_next_epoch(ws, stream, &p.epoch);
    if (rc) return rc;
    p.state = ws_state(ws);
    p.ntiles = ntiles;
    int64_t* splits = reinterpret_cast<int64_t*>(ws->payload);
    p.splits = splits;
    auto kern = setop_kernel<V, VO, KIND, HAS_VAL, SEAM>;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        SPB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        SPB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        SPB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, kNT, smem));
        if (ctas_per_sm < 1) return SPB_ERR_UNSUPPORTED;
    }
    const int64_t pwarps = ntiles + 1;
    merge_partition_kernel<<<(unsigned)ceil_div(pwarps * 32, 256), 256, 0, stream>>>(p.a, p.na, p.b, p.nb, ntiles,
                                                                                      splits);
    SPB_LAUNCHED(1);
    int64_t grid = (int64_t)device_sm_count() * ctas_per_sm;          // persistent: every CTA is resident
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, kNT, smem, stream>>>(p);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

template <int KIND>
int dispatch_binop(int op, int dtype, SetopParams& p, spb_workspace* ws, cudaStream_t stream) {
    if (op < 0 || op >= SPB_BINOP_COUNT) return SPB_ERR_BAD_ARG;
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
