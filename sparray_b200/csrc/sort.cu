// sparray_b200/csrc/sort.cu -- index re-keying (unravel / permute or drop axes / re-ravel) fused into
// a one-sweep LSD radix sort of (int64 key, value) pairs.
//
// Replaces the np.unravel_index / np.ravel_multi_index temporaries and the argsort-based np.unique of
// FlatSparray.transpose (flat.py:109-112 + ctor :34), of sum(axis != last) (flat.py:617-620) and of
// the non-canonical ctor (flat.py:32-35).
//
//   radix_hist_kernel : ONE read of the keys (re-keyed on the fly), the digit histograms of all passes;
//   radix_scan_kernel : counts -> exclusive offsets (one tiny block per pass);
//   radix_pass_kernel : one launch per digit (<= 8 bits).  A CTA owns a tile of 4096 (8-byte values:
//     3072) consecutive elements: stable tile-local ranking by warp MATCH + per-warp counters, per-digit
//     decoupled look-back across tiles, exchange through shared memory, run-coalesced scatter.  The
//     tile's values are fetched by ONE TMA bulk copy (cp.async.bulk + mbarrier) at CTA start and never
//     pass through registers: the scatter reads them from the staged tile through a 16-bit source map.
//
// Traffic is what bounds a sort, so the passes move as few bytes as the key allows:
//   * only the bits that matter are sorted (the host strips the trailing axes that already arrive in
//     order: an LSD sort is stable), in as few, equally wide digits as possible;
//   * between passes keys travel as 32-bit PACKED keys whenever they fit: P = (S << lowbits) | L with
//     S = new_key / sort_divisor (the part being sorted) and L = new_key % sort_divisor; the first
//     pass re-keys and packs, the last pass unpacks to the final int64 key;
//   * when P needs 33..40 bits (BASELINE configs 3, 4, 5) the digit a pass has just sorted is SQUEEZED
//     out of the intermediate key -- the next pass recovers it from the element's position (the input
//     of pass p+1 is sorted by digit p, and the global digit offsets are known) and squeezes out its
//     own digit instead.  32-bit intermediates up to 2^40 cells;
//   * beyond that, 64-bit packed keys.
#include "rekey.cuh"

namespace spb {

constexpr int kSNT = 256;                 // threads per CTA (= max bins: one digit per thread in the scan phases)
constexpr int kSWarps = kSNT / 32;
constexpr int kMaxPasses = 8;

enum { MODE_K32 = 0, MODE_SQUEEZE = 1, MODE_WIDE = 2 };

// How new keys are packed for sorting: P = (S << lowbits) | L
struct Packing {
    FastDiv sort_div;       // S = key / sort_div.d
    int lowbits;            // bits of L (0 when sort_div == 1)
    int pow2;               // sort_div is a power of two: P == key
};

__device__ __forceinline__ uint64_t pack_key(int64_t key, const Packing& pk) {
    if (pk.pow2) return (uint64_t)key;
    const uint64_t s = fastdiv((uint64_t)key, pk.sort_div);
    const uint64_t l = (uint64_t)key - s * pk.sort_div.d;
    return (s << pk.lowbits) | l;
}
__device__ __forceinline__ int64_t unpack_key(uint64_t p, const Packing& pk) {
    if (pk.pow2) return (int64_t)p;
    const uint64_t s = p >> pk.lowbits;
    const uint64_t l = p & ((1ull << pk.lowbits) - 1ull);
    return (int64_t)(s * pk.sort_div.d + l);
}

struct HistParams {
    const int64_t* keys_in;
    int64_t n;
    Rekey rk;
    Packing pk;
    int npasses;
    int shift[kMaxPasses];      // bit position of the pass's digit inside P
    int bins[kMaxPasses];
    uint64_t* hist;             // [kMaxPasses][256] counts (zeroed by the caller)
    int p32;                    // packed keys are re-keyed in 32-bit registers (power-of-two shape, < 2^32 cells)
    // Pre-keying: this kernel re-keys every key anyway, so it also leaves what the FIRST pass would otherwise have
    // to compute again from the 8-byte keys -- the 32-bit packed key (squeeze mode: with the first digit taken
    // out, and that digit in a byte beside it).  The first pass then costs what a middle pass costs.
    uint32_t* pre_keys;         // [n], or null
    uint8_t* pre_dig;           // [n] squeeze mode, else null
    int pre_width;              // squeeze mode: width of the first digit
};

// ---- all digit histograms in one read --------------------------------------------------------
// Four keys per thread are in flight per trip; a warp whose 32 keys share a digit (sorted input: the digit
// is a function of a slowly changing coordinate) adds 32 with one shared-memory atomic instead of a 32-way
// bank conflict.
constexpr int kHistUnroll = 4;

// NP passes known at compile time (0: any number): the digit loop is straight-line code
template <int NP, bool P32>
__global__ void __launch_bounds__(kSNT) radix_hist_kernel(const HistParams p) {
    __shared__ unsigned h[kMaxPasses * 256];
    const int npasses = NP ? NP : p.npasses;
    for (int i = threadIdx.x; i < npasses * 256; i += kSNT) h[i] = 0;
    __syncthreads();
    cudaGridDependencySynchronize();               // the keys may come from the previous launch
    using PT = typename std::conditional<P32, uint32_t, uint64_t>::type;
    const int lane = threadIdx.x & 31;
    const int64_t nthreads = (int64_t)gridDim.x * kSNT;
    const int64_t gtid = blockIdx.x * (int64_t)kSNT + threadIdx.x;
    // whole warps stay converged so the votes below are legal: iterate on a warp-aligned bound
    const int64_t n_round = (p.n + 31) & ~(int64_t)31;
    for (int64_t i0 = gtid; i0 < n_round; i0 += kHistUnroll * nthreads) {
        int64_t raw[kHistUnroll];
        PT P[kHistUnroll];
        bool valid[kHistUnroll];
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
            const int64_t i = i0 + u * nthreads;
            valid[u] = i < p.n;
            raw[u] = valid[u] ? __ldcs(p.keys_in + i) : 0;
        }
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
            if constexpr (P32) P[u] = apply_rekey32(raw[u], p.rk);
            else P[u] = pack_key(apply_rekey(raw[u], p.rk), p.pk);
        }
        if (p.pre_keys) {                              // (uniform)
#pragma unroll
            for (int u = 0; u < kHistUnroll; ++u) {
                const int64_t i = i0 + u * nthreads;
                if (valid[u]) {
                    if (p.pre_dig) {
                        const uint64_t P64 = (uint64_t)P[u];
                        const int sh = p.shift[0];
                        const uint64_t lo = P64 & ((1ull << sh) - 1ull);
                        p.pre_keys[i] = (uint32_t)(((P64 >> (sh + p.pre_width)) << sh) | lo);
                        p.pre_dig[i] = (uint8_t)((P64 >> sh) & ((1u << p.pre_width) - 1u));
                    } else {
                        p.pre_keys[i] = (uint32_t)P[u];
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kHistUnroll; ++u) {
            if (i0 + u * nthreads - lane >= n_round) break;          // (warp-uniform) this row lies past the end
            const bool full = __all_sync(0xffffffffu, valid[u]);
#pragma unroll
            for (int ps = 0; ps < (NP ? NP : kMaxPasses); ++ps) {
                if (ps < npasses) {
                    const unsigned d = (unsigned)(P[u] >> p.shift[ps]) & ((unsigned)p.bins[ps] - 1u);
                    const unsigned d0 = __shfl_sync(0xffffffffu, d, 0);
                    if (full && __all_sync(0xffffffffu, d == d0)) {
                        if (lane == 0) atomicAdd(&h[ps * 256 + d], 32u);
                    } else if (valid[u]) {
                        atomicAdd(&h[ps * 256 + d], 1u);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npasses * 256; i += kSNT)
        if (h[i]) atomicAdd(reinterpret_cast<unsigned long long*>(p.hist) + i, (unsigned long long)h[i]);
}

// counts -> exclusive offsets, one block per pass
__global__ void __launch_bounds__(256) radix_scan_kernel(uint64_t* hist) {
    __shared__ uint64_t s[8];
    cudaGridDependencySynchronize();
    uint64_t* h = hist + blockIdx.x * 256;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint64_t mine = h[t];
    uint64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s[warp] = incl;
    __syncthreads();
    uint64_t wofs = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
        if (w < warp) wofs += s[w];
    h[t] = wofs + incl - mine;
}

// Optional phase timers (-DSPB_SORT_TIMING, profiles/time_sort_phases.py): thread 0 of every CTA stamps the
// global nanosecond timer at phase boundaries into a buffer registered with spb_debug_sort_timing().
#ifdef SPB_SORT_TIMING
__device__ unsigned long long* g_sort_timing = nullptr;     // [ntiles][16]
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define SPB_STAMP(slot)                                                                        \
    do {                                                                                       \
        if (threadIdx.x == 0 && g_sort_timing) g_sort_timing[blockIdx.x * 16 + (slot)] = gtimer(); \
    } while (0)
#else
#define SPB_STAMP(slot) do { } while (0)
#endif

// ---- one digit pass -----------------------------------------------------------------------------
struct PassParams {
    const void* keys_in;        // FIRST: int64 old keys; else packed keys (u32 / u64)
    const void* vals_in;
    void* keys_out;             // LAST: int64 final keys; else packed keys
    void* vals_out;
    int64_t n;
    Rekey rk;
    Packing pk;
    int in_shift;               // bit position of this pass's digit in the representation the pass reads
                                // (FIRST: in P)
    int bins;                   // 1 << width of this pass's digit
    int squeeze;                // MODE_SQUEEZE: the intermediate keys lack the digit sorted last
    int sq_width;               // squeeze: width of every non-last digit
    int prev_bins;
    const uint64_t* off;        // [256] exclusive global offsets of this pass's digits
    const uint64_t* prev_off;   // squeeze, not FIRST: those of the previous pass
    uint64_t* state;            // [ntiles][bins] look-back words
    uint32_t epoch;
    int vals_aligned;           // vals_in is 16-byte aligned: full tiles are fetched by TMA
    int p32;                    // FIRST: re-key in 32-bit registers (see run_sort)
    const uint32_t* pre_keys;   // FIRST: the histogram kernel left the packed keys here (and, squeeze mode, the
    const uint8_t* pre_dig;     // digits of this pass): nothing is re-keyed
};

template <int VB> struct ValType;
template <> struct ValType<0> { using type = uint8_t; };
template <> struct ValType<1> { using type = uint8_t; };
template <> struct ValType<2> { using type = uint16_t; };
template <> struct ValType<4> { using type = uint32_t; };
template <> struct ValType<8> { using type = uint64_t; };

#ifndef SPB_SORT_IPT
#define SPB_SORT_IPT 15
#endif
#ifndef SPB_SORT_IPT8
#define SPB_SORT_IPT8 10
#endif
#ifndef SPB_SORT_LBW
#define SPB_SORT_LBW 4
#endif

template <bool K32, int VB>
struct PassCfg {
    // items per thread: the largest tile that still lets 4 CTAs (32 warps) share an SM's shared memory --
    // 3840 elements with values of up to 4 bytes, 2560 with 8-byte values
    static constexpr int kIPT = (VB == 8 || !K32) ? SPB_SORT_IPT8 : SPB_SORT_IPT;
    static constexpr int kTile = kSNT * kIPT;
    static constexpr int kKeyBytes = K32 ? 4 : 8;
    // dynamic shared memory layout (byte offsets; every array 16-byte aligned)
    static constexpr int oVal = 0;                                  // staged values (TMA destination)
    static constexpr int oKey = oVal + kTile * VB;
    static constexpr int oSrc = oKey + kTile * kKeyBytes;           // u16 source position per sorted slot
    static constexpr int oDig = oSrc + kTile * 2;                   // u8 digit per sorted slot (first pass of a squeezed sort)
    static constexpr int oWhist = oDig + kTile;                     // u32 [8][256]: counters, then slot offsets
    static constexpr int oGoff = oWhist + kSWarps * 256 * 4;        // i64 [256]
    static constexpr int oPoff = oGoff + 256 * 8;                   // u64 [256] previous pass's offsets (squeeze)
    static constexpr int oMisc = oPoff + 256 * 8;                   // scan partials, mbarrier
    static constexpr int kTotal = oMisc + 128;
    static constexpr int kFit = (227 * 1024) / (kTotal + 1024);     // CTAs per SM that shared memory allows
#ifndef SPB_SORT_MAXCTAS
#define SPB_SORT_MAXCTAS 4
#endif
    static constexpr int kCap = K32 ? SPB_SORT_MAXCTAS : 3;         // 64-bit packed keys take more registers
    static constexpr int kMinCtas = kFit < kCap ? kFit : kCap;
};

// digit an element at global position `pos` of a pass's input had in the PREVIOUS pass: the input is
// sorted by that digit, so it is the last bin whose exclusive offset is <= pos
__device__ __forceinline__ unsigned digit_of_position(const uint64_t* s_poff, int nbins, uint64_t pos) {
    int lo = 0, hi = nbins;                      // count of bins with off <= pos, minus one
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_poff[mid] <= pos) lo = mid + 1;
        else hi = mid;
    }
    return (unsigned)(lo - 1);
}

// lanes of the warp holding the same digit of `nbits` <= 8 bits: one ballot per bit (MATCH.ANY issues about
// once per 20 cycles per SM, which is what bounded the ranking when it was used here)
__device__ __forceinline__ unsigned match_digit(unsigned d, int nbits) {
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (b < nbits) {                           // uniform
            const bool bit = (d >> b) & 1u;
            const unsigned bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
    }
    return peers;
}

template <bool FIRST, bool LAST, bool K32, int VB>
__global__ void __launch_bounds__(kSNT, PassCfg<K32, VB>::kMinCtas) radix_pass_kernel(const PassParams p) {
    using C = PassCfg<K32, VB>;
    using KS = typename std::conditional<K32, uint32_t, uint64_t>::type;     // key as staged / as written between passes
    using VT = typename ValType<VB>::type;
    constexpr int kIPT = C::kIPT, kTile = C::kTile;
    extern __shared__ __align__(128) char smem[];
    VT* const s_val = reinterpret_cast<VT*>(smem + C::oVal);
    KS* const s_key = reinterpret_cast<KS*>(smem + C::oKey);
    uint16_t* const s_src = reinterpret_cast<uint16_t*>(smem + C::oSrc);
    uint8_t* const s_dig = reinterpret_cast<uint8_t*>(smem + C::oDig);
    uint32_t* const s_whist = reinterpret_cast<uint32_t*>(smem + C::oWhist);
    int64_t* const s_goff = reinterpret_cast<int64_t*>(smem + C::oGoff);
    uint64_t* const s_poff = reinterpret_cast<uint64_t*>(smem + C::oPoff);
    int* const s_scan = reinterpret_cast<int*>(smem + C::oMisc);              // [8]
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + C::oMisc + 64);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * kTile;
    const int items = (int)((p.n - base) < kTile ? (p.n - base) : kTile);
    const int wbase = warp * 32 * kIPT;
    uint32_t* const my_hist = s_whist + warp * 256;
    const unsigned dmask = (unsigned)p.bins - 1u;
    const int nbits = 32 - __clz(dmask);
    // the first pass of a squeezed sort stages keys that no longer hold the digit: it is kept beside them
    const bool keep_digit = FIRST && !LAST && p.squeeze;
    const bool keep_src = VB != 0 || (!FIRST && p.squeeze);

    // the warp's private digit counters
#pragma unroll
    for (int i = 0; i < 8; ++i) my_hist[i * 32 + lane] = 0;
    bool tma_vals = false;
    if constexpr (VB != 0) {
        tma_vals = p.vals_aligned && items == kTile;
        if (tid == 0 && tma_vals) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
    }
    SPB_STAMP(0);
    cudaGridDependencySynchronize();               // keys, values and offsets come from the previous launches
    SPB_STAMP(1);
    if constexpr (VB != 0) {
        const VT* vin = reinterpret_cast<const VT*>(p.vals_in) + base;
        if (tma_vals) {
            if (tid == 0) {                        // the whole tile of values in one bulk copy; lands while we rank
                mbar_arrive_expect_tx(bar, (uint32_t)(kTile * VB));
                tma_bulk_g2s(s_val, vin, (uint32_t)(kTile * VB), bar);
            }
        } else {                                   // last (partial) tile, or an unaligned slice
            for (int e = tid; e < items; e += kSNT) s_val[e] = vin[e];
        }
    }

    // ---- load, (re-key + pack), digit ---------------------------------------------------------
    // Elements past the end of the array (last tile only) get the largest digit: they are the last
    // elements of the tile, so they rank behind everything real and are never written out.
    KS key[kIPT];
    uint32_t dr[kIPT];                             // digit, later digit | slot << 8
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        const int e = wbase + k * 32 + lane;
        key[k] = 0;
        dr[k] = dmask;
        if (e < items) {
            if constexpr (FIRST) {
                if constexpr (K32 && !LAST) {
                    if (p.pre_keys) {              // (uniform) pre-keyed by the histogram kernel
                        const uint32_t pkd = __ldcs(p.pre_keys + base + e);
                        key[k] = (KS)pkd;
                        dr[k] = p.pre_dig ? (unsigned)__ldcs(p.pre_dig + base + e) : ((pkd >> p.in_shift) & dmask);
                        continue;
                    }
                }
                const int64_t raw = __ldcs(reinterpret_cast<const int64_t*>(p.keys_in) + base + e);
                if constexpr (K32) {
                    if (p.p32) {                   // (uniform) P == new key < 2^32, nothing squeezed
                        const uint32_t P32 = apply_rekey32(raw, p.rk);
                        dr[k] = (P32 >> p.in_shift) & dmask;
                        key[k] = (KS)P32;
                        continue;
                    }
                }
                const uint64_t P = pack_key(apply_rekey(raw, p.rk), p.pk);
                dr[k] = (unsigned)(P >> p.in_shift) & dmask;
                if (!LAST && p.squeeze) {
                    // squeeze this pass's digit out: [hi | digit | lo] -> [hi | lo]
                    const uint64_t lo = P & ((1ull << p.in_shift) - 1ull);
                    key[k] = (KS)(((P >> (p.in_shift + p.sq_width)) << p.in_shift) | lo);
                } else {
                    key[k] = (KS)P;
                }
            } else {
                const KS in = __ldcs(reinterpret_cast<const KS*>(p.keys_in) + base + e);
                key[k] = in;
                dr[k] = (unsigned)(in >> p.in_shift) & dmask;
            }
        }
    }
    __syncwarp();
    SPB_STAMP(2);

    // ---- tile-local stable ranking: per warp, item rows in order, lanes in order inside a row ----
    // Lanes holding the same digit find each other (eight ballots); the first of them bumps the warp's
    // counter by the group size with a shared-memory atomic and everybody reads the old value back from
    // it.  The kIPT atomics of a warp are independent instructions (the counters see them in program
    // order), so their latencies overlap instead of forming a read-modify-write chain.
    {
        uint32_t old[kIPT];
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            const unsigned d = dr[k];
            const unsigned peers = match_digit(d, nbits);
            const int leader = __ffs(peers) - 1;
            old[k] = 0;
            if (lane == leader) old[k] = atomicAdd(my_hist + d, (unsigned)__popc(peers));
            dr[k] = d | ((unsigned)leader << 8) | ((unsigned)__popc(peers & lt_mask) << 13);
        }
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            const unsigned o = __shfl_sync(0xffffffffu, old[k], (dr[k] >> 8) & 31u);
            dr[k] = (dr[k] & 255u) | ((o + (dr[k] >> 13)) << 8);
        }
    }

    // squeeze: which digit of the previous pass does this tile hold?  (uniform for almost every tile)
    bool prev_inside = false;
    int prev_le = 0;
    if constexpr (!FIRST) {
        if (p.squeeze) {
            const uint64_t o = tid < p.prev_bins ? p.prev_off[tid] : ~0ull;
            s_poff[tid] = o;
            prev_le = o <= (uint64_t)base;
            prev_inside = o > (uint64_t)base && o < (uint64_t)(base + items);
        }
    }
    SPB_STAMP(3);
    const int n_le = __syncthreads_count(prev_le);                   // (1) every warp's counters are final
    SPB_STAMP(4);
    const unsigned dprev_first = (unsigned)(n_le - 1);

    // ---- per digit (thread d): tile count, exclusive offsets across warps and digits -----------------
    unsigned cw[kSWarps];
    unsigned sum = 0;
#pragma unroll
    for (int w = 0; w < kSWarps; ++w) {
        cw[w] = s_whist[w * 256 + tid];
        sum += cw[w];
    }
    // the padding of the last tile must not reach the successors' offsets (there are none) -- but keep the
    // published count exact anyway
    if (tid == (int)dmask) sum -= (unsigned)(kTile - items);
    uint64_t* const st = p.state + tile * p.bins + tid;
    if (tid < p.bins) st_state(st, pack_state(tile == 0 ? kStPrefix : kStAgg, p.epoch, sum));
    if (tid == (int)dmask) sum += (unsigned)(kTile - items);
    // block exclusive scan of `sum` over the digits
    int incl = (int)sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    const int any_inside = __syncthreads_or(prev_inside);            // (2)
    int wofs = 0;
#pragma unroll
    for (int w = 0; w < kSWarps; ++w)
        if (w < warp) wofs += s_scan[w];
    const int dstart = wofs + incl - (int)sum;
    // slot of the first element of (warp w, digit d) in the digit-sorted tile
    {
        unsigned run = (unsigned)dstart;
#pragma unroll
        for (int w = 0; w < kSWarps; ++w) {
            s_whist[w * 256 + tid] = run;
            run += cw[w];
        }
    }
    if (tid == (int)dmask) sum -= (unsigned)(kTile - items);
    SPB_STAMP(5);
    __syncthreads();                                                 // (3) the slot offsets are in place
    SPB_STAMP(6);

    // ---- exchange through shared memory into digit order ------------------------------------------
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        const int e = wbase + k * 32 + lane;
        const unsigned d = dr[k] & 255u;
        const int r = (int)my_hist[d] + (int)(dr[k] >> 8);
        s_key[r] = key[k];
        if (keep_digit) s_dig[r] = (uint8_t)d;
        if (keep_src) s_src[r] = (uint16_t)e;
    }
    SPB_STAMP(7);

    // ---- per-digit look-back -> global offset of every digit's run ----------------------------------
    // SPB_SORT_LBW predecessors are requested per round.  Every round reads bins x SPB_SORT_LBW words per
    // tile through L2: 4 measured best (2: 1 % slower, 8: 3 %, 16: 15 % -- the words of a wide window are as
    // many bytes as the tile's payload).
    if (tid < p.bins) {
        uint64_t excl = 0;
#ifdef SPB_SORT_NOLB
        if (false) {                               // timing experiment only: results are wrong
#else
        if (tile > 0) {
#endif
            constexpr int kLB = SPB_SORT_LBW;
            int64_t t = tile - 1;
            bool done = false;
            while (!done && t >= 0) {
                uint64_t lb[kLB];
#pragma unroll
                for (int j = 0; j < kLB; ++j) {
                    const int64_t tt = t - j;
                    lb[j] = tt >= 0 ? ld_state(p.state + tt * p.bins + tid) : pack_state(kStPrefix, p.epoch, 0);
                }
#pragma unroll
                for (int j = 0; j < kLB; ++j) {
                    if (!done) {
                        const int64_t tt = t - j;
                        while ((lb[j] >> 62) == 0 || (uint32_t)((lb[j] >> 48) & ((1u << kEpochBits) - 1)) != p.epoch) {
                            __nanosleep(20);
                            lb[j] = ld_state(p.state + tt * p.bins + tid);
                        }
                        excl += lb[j] & kCountMask;
                        done = (lb[j] >> 62) == kStPrefix;
                    }
                }
                t -= kLB;
            }
            st_state(st, pack_state(kStPrefix, p.epoch, excl + sum));
        }
        s_goff[tid] = (int64_t)(p.off[tid] + excl) - dstart;
    }
    SPB_STAMP(8);
    __syncthreads();                                                 // (4) sorted tile + offsets complete
    SPB_STAMP(9);
    if constexpr (VB != 0) {
        if (tma_vals) mbar_wait(bar, 0);
    }

    SPB_STAMP(10);
    // ---- run-coalesced scatter: slot q of the sorted tile goes to goff[digit] + q --------------------
    VT* const vout = reinterpret_cast<VT*>(p.vals_out);
#pragma unroll 5
    for (int j = 0; j < kIPT; ++j) {
        const int q = j * kSNT + tid;
        if (q < items) {
            const KS ks = s_key[q];
            const unsigned d = keep_digit ? (unsigned)s_dig[q] : ((unsigned)(ks >> p.in_shift) & dmask);
            const int64_t dst = s_goff[d] + q;
            [[maybe_unused]] unsigned src = 0;
            if (keep_src) src = s_src[q];
            if constexpr (FIRST && !LAST) {
                reinterpret_cast<KS*>(p.keys_out)[dst] = ks;
            } else if constexpr (FIRST && LAST) {
                reinterpret_cast<int64_t*>(p.keys_out)[dst] = unpack_key((uint64_t)ks, p.pk);
            } else {
                uint64_t full = (uint64_t)ks;
                if (p.squeeze) {
                    const unsigned dprev = any_inside ? digit_of_position(s_poff, p.prev_bins, (uint64_t)(base + src))
                                                      : dprev_first;
                    // [hi | this digit | lo] -> re-insert the previous digit below this one
                    const uint64_t lo = full & ((1ull << p.in_shift) - 1ull);
                    const uint64_t hi = full >> p.in_shift;
                    if constexpr (LAST) {
                        full = (hi << (p.in_shift + p.sq_width)) | ((uint64_t)dprev << p.in_shift) | lo;
                    } else {
                        // ... and squeeze this pass's digit out (same width: the field is replaced in place)
                        full = ((hi >> p.sq_width) << (p.in_shift + p.sq_width)) | ((uint64_t)dprev << p.in_shift) | lo;
                    }
                }
                if constexpr (LAST) reinterpret_cast<int64_t*>(p.keys_out)[dst] = unpack_key(full, p.pk);
                else reinterpret_cast<KS*>(p.keys_out)[dst] = (KS)full;
            }
            if constexpr (VB != 0) vout[dst] = s_val[src];
        }
    }
    SPB_STAMP(11);
}

// no digits to sort: just re-key (and copy the values)
template <typename V>
__global__ void __launch_bounds__(256) rekey_copy_kernel(const int64_t* __restrict__ keys_in, const V* __restrict__ vin,
                                                         int64_t* __restrict__ keys_out, V* __restrict__ vout, int64_t n,
                                                         const Rekey rk) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        keys_out[i] = apply_rekey(keys_in[i], rk);
        if (vin) vout[i] = vin[i];
    }
}

// ---- host side: pass plan + launches -------------------------------------------------------------
struct SortPlan {
    int npasses, mode, lowbits, total_bits, sq_width;
    int width[kMaxPasses], shift[kMaxPasses];
};

static SortPlan make_sort_plan(uint64_t sort_divisor, int sort_bits) {
    SortPlan pl{};
    int lowbits = 0;
    while ((1ull << lowbits) < sort_divisor) ++lowbits;
    pl.lowbits = lowbits;
    pl.total_bits = sort_bits + lowbits;
    const int np = (sort_bits + 7) / 8;
    pl.npasses = np;
    if (np == 0) return pl;
    int base_w = sort_bits / np, extra = sort_bits % np;             // equally wide digits
    pl.mode = pl.total_bits <= 32 ? MODE_K32 : MODE_WIDE;
    if (np >= 2 && pl.total_bits > 32) {
        // squeezing the digit just sorted out of the intermediate keys makes them 32 bits wide when
        // every non-last digit has at least total_bits - 32 bits; those digits all get one width
        int w = (sort_bits + np - 1) / np;
        if (w < pl.total_bits - 32) w = pl.total_bits - 32;
        if (w <= 8 && w * (np - 1) < sort_bits) {
            pl.mode = MODE_SQUEEZE;
            pl.sq_width = w;
            int sh = 0;
            for (int ps = 0; ps < np; ++ps) {
                pl.width[ps] = ps < np - 1 ? w : sort_bits - w * (np - 1);
                pl.shift[ps] = sh;
                sh += pl.width[ps];
            }
            return pl;
        }
    }
    int sh = 0;
    for (int ps = 0; ps < np; ++ps) {
        pl.width[ps] = base_w + (ps < extra ? 1 : 0);
        pl.shift[ps] = sh;
        sh += pl.width[ps];
    }
    return pl;
}

template <bool FIRST, bool LAST, bool K32, int VB>
static int launch_pass(const PassParams& p, int64_t ntiles, cudaStream_t stream) {
    using C = PassCfg<K32, VB>;
    auto kern = radix_pass_kernel<FIRST, LAST, K32, VB>;
    static PerDevice configured;
    if (configured.first_use(current_device())) {
        SPB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
        SPB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    SPB_CUDA_TRY(launch_pdl(kern, dim3((unsigned)ntiles), dim3(kSNT), (size_t)C::kTotal, stream, p));
    SPB_LAUNCHED(1);
    return SPB_OK;
}

template <bool K32, int VB>
static int launch_pass_fl(bool first, bool last, const PassParams& p, int64_t ntiles, cudaStream_t stream) {
    if (first && last) return launch_pass<true, true, K32, VB>(p, ntiles, stream);
    if (first) return launch_pass<true, false, K32, VB>(p, ntiles, stream);
    if (last) return launch_pass<false, true, K32, VB>(p, ntiles, stream);
    return launch_pass<false, false, K32, VB>(p, ntiles, stream);
}

template <int VB>
int run_sort(const int64_t* idx, const void* val, int64_t n, const Rekey& rk, uint64_t sort_divisor, int sort_bits,
             int64_t* out_idx, void* out_val, int64_t* tmp_idx, void* tmp_val, spb_workspace* ws, cudaStream_t stream) {
    using VT = typename ValType<VB>::type;
    const SortPlan pl = make_sort_plan(sort_divisor, sort_bits);
    const int np = pl.npasses;
    if (np == 0) {
        int64_t blocks = ceil_div(n, 256 * 4);
        const int64_t cap = (int64_t)device_sm_count() * 16;
        if (blocks > cap) blocks = cap;
        rekey_copy_kernel<VT><<<(unsigned)blocks, 256, 0, stream>>>(idx, VB ? (const VT*)val : nullptr, out_idx,
                                                                    (VT*)out_val, n, rk);
        SPB_LAUNCHED(1);
        return (int)cudaGetLastError();
    }
    if (np > kMaxPasses) return SPB_ERR_BAD_ARG;
    if (np > 1 && (!tmp_idx || (VB && !tmp_val))) return SPB_ERR_BAD_ARG;
    const bool k32 = pl.mode != MODE_WIDE;
    const int tile = k32 ? PassCfg<true, VB>::kTile : PassCfg<false, VB>::kTile;
    const int64_t ntiles = ceil_div(n, tile);
    if (ntiles > 0x7fffffffLL) return SPB_ERR_TOO_LARGE;
    int max_bins = 1;
    for (int ps = 0; ps < np; ++ps) max_bins = (1 << pl.width[ps]) > max_bins ? (1 << pl.width[ps]) : max_bins;
    int rc = workspace_reserve(ws, (size_t)ntiles * max_bins * 8, stream);
    if (rc) return rc;

    Packing pk{};
    pk.sort_div = make_fastdiv(sort_divisor);
    pk.lowbits = pl.lowbits;
    pk.pow2 = pk.sort_div.magic == 0;

    uint64_t* const hist = reinterpret_cast<uint64_t*>(ws_misc(ws) + (512 << 10));     // upper half of the misc slab
    SPB_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint64_t) * kMaxPasses * 256, stream));
    HistParams hp{};
    hp.keys_in = idx; hp.n = n; hp.rk = rk; hp.pk = pk; hp.npasses = np; hp.hist = hist;
    for (int ps = 0; ps < np; ++ps) {
        hp.shift[ps] = pl.lowbits + pl.shift[ps];
        hp.bins[ps] = 1 << pl.width[ps];
    }
    int64_t hblocks = ceil_div(n, kSNT * 16);
    const int64_t hcap = (int64_t)device_sm_count() * 8;
    if (hblocks > hcap) hblocks = hcap;
    // power-of-two shape whose new keys fit 32 bits: the re-key runs in 32-bit registers
    const bool p32 = rk.ndim > 0 && rk.pow2 && rk.out32 && pk.pow2 && pl.total_bits <= 32;
    hp.p32 = p32;
    // Pre-keying (two passes or more, 32-bit intermediate keys): the histogram kernel leaves the packed keys -- and
    // in squeeze mode the first digits -- in the scratch array the first pass does NOT write to (the second pass
    // will overwrite it: by then it has been consumed).  4n (+ n) of its 8n bytes.
    const bool to_final0 = ((np - 1) % 2) == 0;
    uint32_t* pre_keys = nullptr;
    uint8_t* pre_dig = nullptr;
    if (np >= 2 && pl.mode != MODE_WIDE) {
        char* const buf = reinterpret_cast<char*>(to_final0 ? tmp_idx : out_idx);
        pre_keys = reinterpret_cast<uint32_t*>(buf);
        if (pl.mode == MODE_SQUEEZE) pre_dig = reinterpret_cast<uint8_t*>(buf + (size_t)n * 4);
    }
    hp.pre_keys = pre_keys;
    hp.pre_dig = pre_dig;
    hp.pre_width = pl.sq_width;
    {
        const dim3 hg((unsigned)hblocks), hb(kSNT);
        cudaError_t e;
        if (p32 && np == 1) e = launch_pdl(radix_hist_kernel<1, true>, hg, hb, 0, stream, hp);
        else if (p32 && np == 2) e = launch_pdl(radix_hist_kernel<2, true>, hg, hb, 0, stream, hp);
        else if (p32 && np == 3) e = launch_pdl(radix_hist_kernel<3, true>, hg, hb, 0, stream, hp);
        else if (p32 && np == 4) e = launch_pdl(radix_hist_kernel<4, true>, hg, hb, 0, stream, hp);
        else if (np == 1) e = launch_pdl(radix_hist_kernel<1, false>, hg, hb, 0, stream, hp);
        else if (np == 2) e = launch_pdl(radix_hist_kernel<2, false>, hg, hb, 0, stream, hp);
        else if (np == 3) e = launch_pdl(radix_hist_kernel<3, false>, hg, hb, 0, stream, hp);
        else if (np == 4) e = launch_pdl(radix_hist_kernel<4, false>, hg, hb, 0, stream, hp);
        else e = launch_pdl(radix_hist_kernel<0, false>, hg, hb, 0, stream, hp);
        SPB_CUDA_TRY(e);
    }
    SPB_LAUNCHED(1);
    SPB_CUDA_TRY(launch_pdl(radix_scan_kernel, dim3((unsigned)np), dim3(256), 0, stream, hist));
    SPB_LAUNCHED(1);

    // intermediate buffers: packed keys are written into the int64 scratch arrays (the first 4n bytes
    // of them when packed keys are 32 bits wide); ping-pong so that the last pass lands in out_*
    PassParams p{};
    p.n = n; p.rk = rk; p.pk = pk; p.state = ws_state(ws);
    p.squeeze = pl.mode == MODE_SQUEEZE;
    p.p32 = p32 && !p.squeeze;
    p.sq_width = pl.sq_width;
    const void* kin = idx;
    const void* vin = val;
    for (int ps = 0; ps < np; ++ps) {
        const bool first = ps == 0, last = ps == np - 1;
        const bool to_final = ((np - 1 - ps) % 2) == 0;
        p.keys_in = kin; p.vals_in = vin;
        p.keys_out = to_final ? (void*)out_idx : (void*)tmp_idx;
        p.vals_out = to_final ? out_val : tmp_val;
        p.bins = 1 << pl.width[ps];
        p.off = hist + ps * 256;
        p.prev_off = ps > 0 ? hist + (ps - 1) * 256 : nullptr;
        p.prev_bins = ps > 0 ? (1 << pl.width[ps - 1]) : 0;
        // where this pass's digit sits in what the pass reads: the packed key, minus (squeeze) the
        // previous digit, which sat right below it
        p.in_shift = pl.lowbits + pl.shift[ps] - ((p.squeeze && !first) ? pl.sq_width : 0);
        p.vals_aligned = VB ? ((reinterpret_cast<uintptr_t>(vin) & 15) == 0) : 0;
        p.pre_keys = first ? pre_keys : nullptr;
        p.pre_dig = first ? pre_dig : nullptr;
        rc = workspace_next_epoch(ws, stream, &p.epoch);
        if (rc) return rc;
        rc = k32 ? launch_pass_fl<true, VB>(first, last, p, ntiles, stream)
                 : launch_pass_fl<false, VB>(first, last, p, ntiles, stream);
        if (rc) return rc;
        kin = p.keys_out;
        vin = p.vals_out;
    }
    return (int)cudaGetLastError();
}

}  // namespace spb

using namespace spb;

extern "C" {

#ifdef SPB_SORT_TIMING
int spb_debug_sort_timing(unsigned long long* buf) {
    return (int)cudaMemcpyToSymbol(g_sort_timing, &buf, sizeof(buf));
}
#endif

int spb_rekey_sort(int dtype, const int64_t* idx, const void* val, int64_t n, int ndim,
                   const int64_t* in_strides, const int64_t* out_strides, int64_t sort_divisor, int sort_bits,
                   int64_t* out_idx, void* out_val, int64_t* tmp_idx, void* tmp_val, spb_workspace* ws,
                   spb_stream_t stream) {
    if (n < 0 || ndim < 0 || ndim > kMaxRekeyDims || sort_divisor < 1 || sort_bits < 0 || sort_bits > 63)
        return SPB_ERR_BAD_ARG;
    if (n == 0) return SPB_OK;
    if (!idx || !out_idx || (val && !out_val) || (ndim && (!in_strides || !out_strides))) return SPB_ERR_BAD_ARG;
    Rekey rk{};
    const int rc = make_rekey(rk, ndim, in_strides, out_strides);
    if (rc) return rc;
    int vb = 0;
    if (val) {
        switch (dtype) {
            case SPB_F32: case SPB_I32: vb = 4; break;
            case SPB_F64: case SPB_I64: vb = 8; break;
            case SPB_U8: case SPB_BOOL: case SPB_I8: vb = 1; break;
            case SPB_I16: vb = 2; break;
            default: return SPB_ERR_UNSUPPORTED;
        }
    }
    const uint64_t sd = (uint64_t)sort_divisor;
    switch (vb) {
        case 0: return run_sort<0>(idx, val, n, rk, sd, sort_bits, out_idx, out_val, tmp_idx, tmp_val, ws, stream);
        case 1: return run_sort<1>(idx, val, n, rk, sd, sort_bits, out_idx, out_val, tmp_idx, tmp_val, ws, stream);
        case 2: return run_sort<2>(idx, val, n, rk, sd, sort_bits, out_idx, out_val, tmp_idx, tmp_val, ws, stream);
        case 4: return run_sort<4>(idx, val, n, rk, sd, sort_bits, out_idx, out_val, tmp_idx, tmp_val, ws, stream);
        default: return run_sort<8>(idx, val, n, rk, sd, sort_bits, out_idx, out_val, tmp_idx, tmp_val, ws, stream);
    }
}

}  // extern "C"
