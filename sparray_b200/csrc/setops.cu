// sparray_b200/csrc/setops.cu -- merge-path set union / intersection of sorted
// duplicate-free int64 index arrays with the value op fused in.
//
// Replaces, in ONE kernel launch per call:
//   * merge_unique   (_merge.pyx:107-140) + the three mask gathers / scatters of
//     _pairwise_sparray (flat.py:416-421)                         -> union kernels
//   * merge_intersect (_merge.pyx:145-157) + the two searchsorted of the wrapper
//     (_merge.pyx:87-88) + the gathers of _pairwise_sparray_fixed_zero
//     (flat.py:429-433)                                           -> intersection kernels
//
// Design (one CTA = one tile of kTile consecutive elements of the merged order):
//   1. two warps find the tile's merge-path split on the two global arrays
//      (32-ary cooperative search, ~5 dependent round trips);
//   2. one thread stages the A / B index ranges and value ranges into shared
//      memory with TMA bulk copies (cp.async.bulk + mbarrier); the < 16-byte
//      unaligned heads / tails and the +-1 halo keys go by ordinary loads;
//   3. every thread merge-path-splits the tile in shared memory and serially
//      merges kVT items, applying the value op; a key present in both inputs is
//      emitted once, by the A element (its B partner is found through the halo
//      even when the pair straddles a tile boundary);
//   4. block scan of the per-thread output counts, decoupled look-back across
//      tiles for the output offset, compaction through shared memory, coalesced
//      stores.  Nothing is read twice and no lut / mask array ever exists.
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

template <typename V>
__device__ __forceinline__ V arith_op(int op, V x, V y) {
    switch (op) {
        case SPB_ADD: return x + y;
        case SPB_MUL: return x * y;
        case SPB_MINIMUM: return np_minimum<V>(x, y);
        case SPB_MAXIMUM: return np_maximum<V>(x, y);
        case SPB_TAKE_A: return x;
        default: return y;   // SPB_TAKE_B
    }
}

// value op of the fused kernels: arithmetic keeps the operand type, comparisons give a bool byte
template <typename V, typename VO>
__device__ __forceinline__ VO value_op(int op, int normalise_bool, V x, V y);

template <typename V>
__device__ __forceinline__ uint8_t compare_op(int op, V x, V y) {
    switch (op) {
        case SPB_LT: return x < y;
        case SPB_LE: return x <= y;
        case SPB_GT: return x > y;
        case SPB_GE: return x >= y;
        case SPB_EQ: return x == y;
        default: return x != y;   // SPB_NE
    }
}

template <typename V, typename VO>
__device__ __forceinline__ VO value_op(int op, int normalise_bool, V x, V y) {
    if constexpr (sizeof(VO) == sizeof(V)) {
        V r = arith_op<V>(op, x, y);
        // bool arithmetic (numpy: + is OR, * is AND): clamp the byte result to {0,1}
        if constexpr (sizeof(V) == 1) {
            if (normalise_bool) r = (V)(r != V(0));
        }
        return (VO)r;
    } else {
        return (VO)compare_op<V>(op, x, y);
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
    uint64_t* state;     // one word per tile
    uint32_t epoch;
    int op;
    int normalise_bool;
    int negate_b;        // a - b is evaluated as a + (-b), base.py:45-46: a missing partner is +0, not -0
};

enum { KIND_UNION = 0, KIND_INTERSECT = 1 };

constexpr int kNT = 256;   // threads per CTA
constexpr int kVT = 8;     // merged items per thread
constexpr int kTile = kNT * kVT;

// shared-memory carve-up (bytes)
template <int VBYTES>
struct SetopSmem {
    static constexpr int kHeader = 128;                               // mbarrier, splits, scan scratch
    static constexpr int kKeyBytes = (kTile + 4) * 8 + 2 * 64;        // A and B key regions incl. pads
    static constexpr int kValBytes = VBYTES ? kTile * VBYTES + 2 * 48 : 0;
    static constexpr int kTotal = kHeader + kKeyBytes + kValBytes;
};

template <typename V, typename VO, int KIND, bool HAS_VAL, bool SEAM>
__global__ void __launch_bounds__(kNT) setop_kernel(const SetopParams p) {
    constexpr int VB = HAS_VAL ? (int)sizeof(V) : (SEAM ? 8 : 0);
    using Smem = SetopSmem<VB>;
    extern __shared__ __align__(128) char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    int64_t* split = reinterpret_cast<int64_t*>(smem + 8);            // [0]=a0 [1]=a1
    int* warp_sums = reinterpret_cast<int*>(smem + 32);               // kNT/32 ints
    int64_t* tile_base_s = reinterpret_cast<int64_t*>(smem + 64);
    int* tile_total_s = reinterpret_cast<int*>(smem + 72);
    char* kbuf = smem + Smem::kHeader;
    char* vbuf = kbuf + Smem::kKeyBytes;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int64_t total_items = p.na + p.nb;
    const int64_t d0 = tile * kTile;
    const int64_t d1 = (d0 + kTile < total_items) ? d0 + kTile : total_items;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        int64_t s = merge_path_warp(p.a, p.na, p.b, p.nb, d0);
        if (lane == 0) split[0] = s;
    } else if (warp == 1) {
        int64_t s = merge_path_warp(p.a, p.na, p.b, p.nb, d1);
        if (lane == 0) split[1] = s;
    }
    __syncthreads();
    const int64_t a0 = split[0], a1 = split[1];
    const int64_t b0 = d0 - a0, b1 = d1 - a1;
    const int an = (int)(a1 - a0), bn = (int)(b1 - b0);
    const int items = an + bn;

    // ---- stage: indices [x0-1, x1] (halo of one key either side), values [x0, x1(+1 for B)) ----
    const int64_t ga_lo = a0 > 0 ? a0 - 1 : 0, ga_hi = a1 < p.na ? a1 + 1 : p.na;
    const int64_t gb_lo = b0 > 0 ? b0 - 1 : 0, gb_hi = b1 < p.nb ? b1 + 1 : p.nb;
    const StagePlan pa = make_stage_plan(p.a, ga_lo, ga_hi, 8);
    const StagePlan pb = make_stage_plan(p.b, gb_lo, gb_hi, 8);
    char* bufA = kbuf + 16;                                            // 16-byte front pad: slot for index -1
    char* bufB = bufA + ((pa.span_bytes() + 15) & ~15u) + 32;          // room for A's +1 sentinel, B's -1 slot
    // As[i] / Bs[i] are indexed by the position inside the tile's range: i in [-1, n]
    int64_t* As = reinterpret_cast<int64_t*>(bufA + (pa.p_lo - pa.g_al)) + (a0 - ga_lo);
    int64_t* Bs = reinterpret_cast<int64_t*>(bufB + (pb.p_lo - pb.g_al)) + (b0 - gb_lo);

    StagePlan pav, pbv;
    char *bufAv = nullptr, *bufBv = nullptr;
    V *Av = nullptr, *Bv = nullptr;
    if constexpr (HAS_VAL) {
        pav = make_stage_plan(p.a_val, a0, a1, sizeof(V));
        pbv = make_stage_plan(p.b_val, b0, gb_hi > b1 ? b1 + 1 : b1, sizeof(V));
        bufAv = vbuf;
        bufBv = bufAv + ((pav.span_bytes() + 15) & ~15u) + 16;
        Av = reinterpret_cast<V*>(bufAv + (pav.p_lo - pav.g_al));
        Bv = reinterpret_cast<V*>(bufBv + (pbv.p_lo - pbv.g_al));
    }

    if (tid == 0) {
        uint32_t bytes = pa.tma_bytes() + pb.tma_bytes();
        if constexpr (HAS_VAL) bytes += pav.tma_bytes() + pbv.tma_bytes();
        mbar_arrive_expect_tx(bar, bytes);
        stage_issue_tma(pa, bufA, bar);
        stage_issue_tma(pb, bufB, bar);
        if constexpr (HAS_VAL) {
            stage_issue_tma(pav, bufAv, bar);
            stage_issue_tma(pbv, bufBv, bar);
        }
    }
    if (warp == 1) {
        stage_copy_edges<int64_t>(pa, bufA, lane);
        if (lane == 8 && a0 == 0) As[-1] = kKeyMin;
        if (lane == 9 && a1 == p.na) As[an] = kKeyMax;
    } else if (warp == 2) {
        stage_copy_edges<int64_t>(pb, bufB, lane);
        if (lane == 8 && b0 == 0) Bs[-1] = kKeyMin;
        if (lane == 9 && b1 == p.nb) Bs[bn] = kKeyMax;
    } else if (warp == 3) {
        if constexpr (HAS_VAL) stage_copy_edges<V>(pav, bufAv, lane);
    } else if (warp == 4) {
        if constexpr (HAS_VAL) stage_copy_edges<V>(pbv, bufBv, lane);
    }
    mbar_wait(bar, 0);
    __syncthreads();

    // ---- per-thread serial merge of kVT items -----------------------------------------------
    int diag = tid * kVT;
    if (diag > items) diag = items;
    int ai = merge_path_thread(As, an, Bs, bn, diag);
    int bi = diag - ai;
    const int steps = (items - diag) < kVT ? (items - diag) : kVT;

    int64_t ka = As[ai], kb = Bs[bi], kprev = As[ai - 1];
    int64_t okey[kVT];
    VO oval[kVT];
    int64_t opa[SEAM && KIND == KIND_INTERSECT ? kVT : 1];
    int64_t opb[SEAM && KIND == KIND_INTERSECT ? kVT : 1];
    uint8_t olut[SEAM && KIND == KIND_UNION ? kVT : 1];
    unsigned emit = 0;

#pragma unroll
    for (int k = 0; k < kVT; ++k) {
        if (k < steps) {
            const bool take_a = ka <= kb;
            const bool match = ka == kb;
            bool e;
            if (take_a) {
                e = (KIND == KIND_UNION) || match;
                okey[k] = ka;
                if constexpr (HAS_VAL) {
                    V x = Av[ai];
                    V y = match ? (p.negate_b ? (V)(-Bv[bi]) : Bv[bi]) : V(0);
                    oval[k] = value_op<V, VO>(p.op, p.normalise_bool, x, y);
                }
                if constexpr (SEAM && KIND == KIND_UNION) {
                    olut[k] = match ? 2 : 0;
                    if (p.a_only) p.a_only[a0 + ai] = match ? 0 : 1;
                }
                if constexpr (SEAM && KIND == KIND_INTERSECT) {
                    opa[k] = a0 + ai;
                    opb[k] = b0 + bi;
                }
                kprev = ka;
                ++ai;
                ka = As[ai];
            } else {
                const bool dup = kb == kprev;
                e = (KIND == KIND_UNION) && !dup;
                okey[k] = kb;
                if constexpr (HAS_VAL) {
                    V y = p.negate_b ? (V)(-Bv[bi]) : Bv[bi];
                    oval[k] = value_op<V, VO>(p.op, p.normalise_bool, V(0), y);
                }
                if constexpr (SEAM && KIND == KIND_UNION) {
                    olut[k] = 1;
                    if (p.b_only) p.b_only[b0 + bi] = dup ? 0 : 1;
                }
                ++bi;
                kb = Bs[bi];
            }
            emit |= (e ? 1u : 0u) << k;
        }
    }

    // ---- block exclusive scan of counts ------------------------------------------------------
    const int cnt = __popc(emit);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();   // also: every thread is done reading the staged inputs
    if (warp == 0) {
        int w = lane < kNT / 32 ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < kNT / 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kNT / 32) warp_sums[lane] = wi - w;   // exclusive warp offsets
        const int tile_total = __shfl_sync(0xffffffffu, wi, kNT / 32 - 1);
        // ---- decoupled look-back -------------------------------------------------------------
        uint64_t base = 0;
        if (tile == 0) {
            if (lane == 0) st_state(p.state, pack_state(kStPrefix, p.epoch, (uint64_t)tile_total));
        } else {
            if (lane == 0) st_state(p.state + tile, pack_state(kStAgg, p.epoch, (uint64_t)tile_total));
            base = lookback_exclusive(p.state, tile, p.epoch);
            if (lane == 0) st_state(p.state + tile, pack_state(kStPrefix, p.epoch, base + (uint64_t)tile_total));
        }
        if (lane == 0) {
            *tile_base_s = (int64_t)base;
            *tile_total_s = tile_total;
            if (d1 == total_items) *p.out_count = (int64_t)base + tile_total;
        }
    }
    __syncthreads();
    const int64_t tile_base = *tile_base_s;
    const int tile_total = *tile_total_s;
    int off = warp_sums[warp] + incl - cnt;

    // ---- compact through shared memory (re-using the input staging space) ---------------------
    int64_t* skey = reinterpret_cast<int64_t*>(kbuf);
    VO* sval = reinterpret_cast<VO*>(vbuf);
    uint8_t* slut = reinterpret_cast<uint8_t*>(vbuf);
    int64_t* spa = reinterpret_cast<int64_t*>(vbuf);
    int64_t* spb_ = spa + kTile / 2;
#pragma unroll
    for (int k = 0; k < kVT; ++k) {
        if (emit & (1u << k)) {
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
    __syncthreads();
    for (int i = tid; i < tile_total; i += kNT) {
        p.out_idx[tile_base + i] = skey[i];
        if constexpr (HAS_VAL) reinterpret_cast<VO*>(p.out_val)[tile_base + i] = sval[i];
        if constexpr (SEAM && KIND == KIND_UNION) {
            if (p.lut) p.lut[tile_base + i] = slut[i];
        }
        if constexpr (SEAM && KIND == KIND_INTERSECT) {
            if (p.a_pos) p.a_pos[tile_base + i] = spa[i];
            if (p.b_pos) p.b_pos[tile_base + i] = spb_[i];
        }
    }
}

}  // namespace spb

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
namespace spb {

int workspace_reserve(spb_workspace* ws, size_t state_bytes, cudaStream_t stream) {
    if (!ws) return SPB_ERR_BAD_ARG;
    const size_t bytes = kWsMiscBytes + state_bytes;
    if (ws->bytes >= bytes && ws->dev) return SPB_OK;
    size_t want = bytes + state_bytes / 2 + (1u << 20);
    if (ws->dev) {
        SPB_CUDA_TRY(cudaStreamSynchronize(stream));
        SPB_CUDA_TRY(cudaFree(ws->dev));
        ws->dev = nullptr;
        ws->bytes = 0;
    }
    SPB_CUDA_TRY(cudaMalloc(&ws->dev, want));
    SPB_CUDA_TRY(cudaMemsetAsync(ws->dev, 0, want, stream));
    ws->bytes = want;
    ws->epoch = 0;
    return SPB_OK;
}

int workspace_reserve_payload(spb_workspace* ws, size_t bytes, cudaStream_t stream) {
    if (!ws) return SPB_ERR_BAD_ARG;
    if (ws->payload_bytes >= bytes && ws->payload) return SPB_OK;
    if (ws->payload) {
        SPB_CUDA_TRY(cudaStreamSynchronize(stream));
        SPB_CUDA_TRY(cudaFree(ws->payload));
        ws->payload = nullptr;
        ws->payload_bytes = 0;
    }
    const size_t want = bytes + bytes / 2 + (1u << 20);
    SPB_CUDA_TRY(cudaMalloc(&ws->payload, want));
    ws->payload_bytes = want;
    return SPB_OK;
}

// next launch epoch (1 .. 2^14-1); re-zero the state words when the counter wraps
int workspace_next_epoch(spb_workspace* ws, cudaStream_t stream, uint32_t* epoch) {
    ws->epoch += 1;
    if (ws->epoch >= (1u << kEpochBits)) {
        SPB_CUDA_TRY(cudaMemsetAsync(ws->dev, 0, ws->bytes, stream));
        ws->epoch = 1;
    }
    *epoch = ws->epoch;
    return SPB_OK;
}

template <typename V, typename VO, int KIND, bool HAS_VAL, bool SEAM>
int launch_setop(SetopParams& p, spb_workspace* ws, cudaStream_t stream) {
    constexpr int VB = HAS_VAL ? (int)sizeof(V) : (SEAM ? 8 : 0);
    constexpr int smem = SetopSmem<VB>::kTotal;
    const int64_t total = p.na + p.nb;
    if (total == 0 || (KIND == KIND_INTERSECT && (p.na == 0 || p.nb == 0))) {
        return (int)cudaMemsetAsync(p.out_count, 0, sizeof(int64_t), stream);
    }
    const int64_t ntiles = ceil_div(total, kTile);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int rc = workspace_reserve(ws, (size_t)ntiles * sizeof(uint64_t), stream);
    if (rc) return rc;
    rc = workspace_next_epoch(ws, stream, &p.epoch);
    if (rc) return rc;
    p.state = ws_state(ws);
    auto kern = setop_kernel<V, VO, KIND, HAS_VAL, SEAM>;
    static bool configured = false;
    if (!configured) {
        SPB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    kern<<<(unsigned)ntiles, kNT, smem, stream>>>(p);
    SPB_LAUNCHED(1);
    return (int)cudaGetLastError();
}

template <int KIND>
int dispatch_binop(int op, int dtype, SetopParams& p, spb_workspace* ws, cudaStream_t stream) {
    if (op < 0 || op >= SPB_BINOP_COUNT) return SPB_ERR_BAD_ARG;
    const bool cmp = op >= SPB_LT && op <= SPB_NE;
    p.op = op == SPB_SUB ? SPB_ADD : op;
    p.negate_b = op == SPB_SUB;
    p.normalise_bool = (dtype == SPB_BOOL && !cmp) ? 1 : 0;
#define SPB_CASE(DT, V)                                                                  \
    case DT:                                                                             \
        return cmp ? launch_setop<V, uint8_t, KIND, true, false>(p, ws, stream)          \
                   : launch_setop<V, V, KIND, true, false>(p, ws, stream);
    switch (dtype) {
        SPB_CASE(SPB_F32, float)
        SPB_CASE(SPB_F64, double)
        SPB_CASE(SPB_I32, int32_t)
        SPB_CASE(SPB_I64, int64_t)
        SPB_CASE(SPB_U8, uint8_t)
        SPB_CASE(SPB_BOOL, uint8_t)
        default: return SPB_ERR_UNSUPPORTED;
    }
#undef SPB_CASE
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_workspace_create(spb_workspace** ws) {
    if (!ws) return SPB_ERR_BAD_ARG;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SPB_ERR_NO_DEVICE;
    spb_workspace* w = new spb_workspace();
    w->dev = nullptr;
    w->bytes = 0;
    w->epoch = 0;
    w->device = dev;
    w->payload = nullptr;
    w->payload_bytes = 0;
    *ws = w;
    return SPB_OK;
}

int spb_workspace_destroy(spb_workspace* ws) {
    if (!ws) return SPB_OK;
    if (ws->dev) cudaFree(ws->dev);
    if (ws->payload) cudaFree(ws->payload);
    delete ws;
    return SPB_OK;
}

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
