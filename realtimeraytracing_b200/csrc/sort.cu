// sort.cu -- Onesweep least-significant-digit radix sort (8-bit digits) for 32/64-bit
// Morton keys with optional 32-bit payload, plus the two pre-pass kernels the reference's
// tests/testsSortGPU pin.
//
// Replaces std::sort over pair<code,index> (bvh.cpp:223-231) -- whose result equals a stable
// sort by code because the indices are an iota -- and the reference's unfinished GPU sort
// (ploc/preprocessing/sort/*.glsl; chainedScanDigitBinning.glsl:31-33 is an empty main).
//
// Kernels
//   radix_histogram_kernel : one read of the keys, all digit places at once (shared atomics)
//   radix_scan_kernel      : exclusive scan of each 256-bin histogram (one warp-scan CTA)
//   onesweep_kernel        : per pass; tile = BLOCK*IPT keys (and payload) staged into shared memory by
//                            TMA bulk copies (cp.async.bulk + one mbarrier), warp-ballot digit ranking
//                            (8 VOTEs per key; MATCH.ANY measured ~10x slower on sm_100), one shared
//                            atomic per digit group for the running warp count, shuffle scans, decoupled
//                            look-back over per-tile digit counts (done before the reorder so the
//                            aggregate->inclusive window stays short), shared-memory reorder,
//                            coalesced scatter
// Algorithmic HBM bytes: n*(s_k + P*2*s_e)  (SURVEY.md 8d): 36 B/key u32 keys, 68 B/pair.
#include "common.cuh"

namespace {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr uint32_t kFlagAgg = 1u << 30;   // tile aggregate available
constexpr uint32_t kFlagIncl = 2u << 30;  // inclusive prefix available
constexpr uint32_t kValueMask = (1u << 30) - 1;
#ifndef RTR_SORT_MINB
#define RTR_SORT_MINB 2   // resident CTAs per SM the register budget is cut for
#endif
#ifndef RTR_SORT_LOOKBATCH
#define RTR_SORT_LOOKBATCH 4
#endif
#ifndef RTR_SORT_RANK_ATOMIC
#define RTR_SORT_RANK_ATOMIC 1   // 1: one shared atomic per digit group and item (items pipelined); 0: read / last peer stores back
#endif
#ifndef RTR_SORT_RANK_GROUP
#define RTR_SORT_RANK_GROUP 4    // items whose atomics are in flight together
#endif
#ifndef RTR_SORT_PREFETCH
#define RTR_SORT_PREFETCH 296   // tiles ahead whose keys are pulled into L2 (two CTAs on each of the 148 SMs)
#endif
constexpr uint32_t kPrefetchTiles = RTR_SORT_PREFETCH;
constexpr int kLookBatch = RTR_SORT_LOOKBATCH;  // predecessors whose status words one look-back round has in flight

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- TMA (bulk async copy) + mbarrier, raw PTX ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, int shift, uint32_t mask) {
    return static_cast<uint32_t>(k >> shift) & mask;
}

// ---------------------------------------------------------------------------------------
// up-front histogram of every digit place (one pass over the keys)
// ---------------------------------------------------------------------------------------
template <typename KeyT, int MAX_PASSES>
__global__ void __launch_bounds__(512)
radix_histogram_kernel(const KeyT* __restrict__ keys, uint32_t n, int begin_bit, int end_bit, int passes,
                       uint32_t* __restrict__ ghist) {
    __shared__ uint32_t s_hist[MAX_PASSES][kRadix];
    for (int i = threadIdx.x; i < MAX_PASSES * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    constexpr int VEC = 16 / sizeof(KeyT);
    const uint32_t nvec = n / VEC;
    const uint4* kv = reinterpret_cast<const uint4*>(keys);
    const bool aligned = (reinterpret_cast<uintptr_t>(keys) & 15u) == 0;
    auto add_key = [&](KeyT k) {
#pragma unroll
        for (int p = 0; p < MAX_PASSES; ++p) {
            if (p < passes) {
                const int shift = begin_bit + p * kRadixBits;
                const int bits = min(kRadixBits, end_bit - shift);
                atomicAdd(&s_hist[p][digit_of<KeyT>(k, shift, (1u << bits) - 1u)], 1u);
            }
        }
    };
    if (aligned) {
        for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
            const uint4 q = __ldg(kv + v);
            if constexpr (sizeof(KeyT) == 4) {
                add_key(q.x); add_key(q.y); add_key(q.z); add_key(q.w);
            } else {
                add_key((KeyT)q.x | ((KeyT)q.y << 32));
                add_key((KeyT)q.z | ((KeyT)q.w << 32));
            }
        }
        for (uint32_t i = nvec * VEC + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            add_key(keys[i]);
    } else {
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            add_key(keys[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// exclusive scan of each pass's 256 bins: gbase[p][d] = #keys with digit < d at place p
__global__ void __launch_bounds__(kRadix)
radix_scan_kernel(const uint32_t* __restrict__ ghist, uint32_t* __restrict__ gbase) {
    __shared__ uint32_t s_warp[kRadix / 32];
    const uint32_t p = blockIdx.x, d = threadIdx.x, l = lane_id(), w = d >> 5;
    const uint32_t v = ghist[p * kRadix + d];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (l >= (uint32_t)o) x += y;
    }
    if (l == 31) s_warp[w] = x;
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t i = 0; i < w; ++i) base += s_warp[i];
    gbase[p * kRadix + d] = base + x - v;
}

// ---------------------------------------------------------------------------------------
// one Onesweep digit pass
// ---------------------------------------------------------------------------------------
// (u32 key, u32 payload) pairs travel through the in-tile reorder as one 64-bit shared-memory slot
template <typename KeyT, bool PAIRS> struct SortPacked { static constexpr bool value = PAIRS && sizeof(KeyT) == 4; };

template <typename KeyT, bool PAIRS, int BLOCK, int IPT>
struct OnesweepSmem {
    static constexpr int TILE = BLOCK * IPT;
    static constexpr int WARPS = BLOCK / 32;
    static constexpr bool PACKED = SortPacked<KeyT, PAIRS>::value;
    // keys as loaded (TMA destination, first TILE entries), later keys -- or (key, payload) slots when
    // PACKED -- in tile-sorted order
    alignas(128) KeyT keys[PACKED ? 2 * TILE : TILE];
    alignas(128) uint32_t vals_in[PAIRS ? TILE : 4];              // payload as loaded (TMA destination)
    alignas(128) uint32_t vals[(PAIRS && !PACKED) ? TILE : 4];    // payload in tile-sorted order
    uint32_t whist[WARPS][kRadix];  // per-warp running digit counts, then first tile-local slot of (warp, digit)
    uint32_t gofs[kRadix];          // global slot of the digit's first key of this tile, minus its tile-local slot
    uint32_t cta_hist[kRadix];      // digit counts of the tile
    uint32_t scan_warp[kRadix / 32];
    alignas(8) uint64_t mbar;
    uint32_t tile;
};

// lanes of the warp holding the same 8-bit digit as the caller: one ballot per digit bit, each narrowing
// the set (R2P + 8 VOTE + 16 predicated LOP3 in SASS; MATCH.ANY is ~10x slower on sm_100)
__device__ __forceinline__ uint32_t digit_peers8(uint32_t d) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int bit = 0; bit < kRadixBits; ++bit) {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t, b, nb;\n\t"
                     "and.b32 t, %1, %2;\n\t"
                     "setp.ne.u32 p, t, 0;\n\t"
                     "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
                     "not.b32 nb, b;\n\t"
                     "@p and.b32 %0, %0, b;\n\t"
                     "@!p and.b32 %0, %0, nb;\n\t}"
                     : "+r"(peers) : "r"(d), "r"(1u << bit));
    }
    return peers;
}
__device__ __forceinline__ uint32_t lanemask_gt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
    return m;
}

#ifdef RTR_SORT_PHASE_CLOCKS
__device__ unsigned long long g_sort_phase_ns[8];   // diagnostics build only (profiles/sort_phases.py)
#define RTR_PHASE(i)                                                              \
    do {                                                                          \
        if (threadIdx.x == 0) {                                                   \
            unsigned long long _t;                                                \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                \
            atomicAdd(&g_sort_phase_ns[i], _t - _phase_t);                        \
            _phase_t = _t;                                                        \
        }                                                                         \
    } while (0)
#else
#define RTR_PHASE(i) do {} while (0)
#endif

template <typename KeyT, bool PAIRS, int BLOCK, int IPT, bool TMA>
__global__ void __launch_bounds__(BLOCK, RTR_SORT_MINB)
onesweep_kernel(const KeyT* __restrict__ keys_in, KeyT* __restrict__ keys_out,
                const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                uint32_t n, uint32_t num_tiles, int shift, uint32_t mask,
                const uint32_t* __restrict__ gbase,   // [256] for this pass
                uint32_t* __restrict__ status,        // [num_tiles][256] zero-initialised
                uint32_t* __restrict__ tile_counter) {
    using Smem = OnesweepSmem<KeyT, PAIRS, BLOCK, IPT>;
    constexpr int TILE = Smem::TILE;
    constexpr int WARPS = Smem::WARPS;
    constexpr bool PACKED = Smem::PACKED;
    static_assert(BLOCK >= kRadix && BLOCK % 32 == 0, "one thread per digit");
    static_assert(TILE < (1 << 16), "tile-local slots are kept in 16 bits");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#ifdef RTR_SORT_PHASE_CLOCKS
    unsigned long long _phase_t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_phase_t));
#endif

    // dynamic tile id: tiles start in issue order, so look-back never waits on an unscheduled CTA
    if (tid == 0) {
        s.tile = atomicAdd(tile_counter, 1u);
        if (TMA) { mbar_init(&s.mbar, 1); mbar_fence_init(); }
    }
    for (uint32_t i = lane; i < kRadix; i += 32) s.whist[warp][i] = 0;
    if (tid < kRadix) s.cta_hist[tid] = 0;
    __syncthreads();
    const uint32_t tile = s.tile;
    RTR_PHASE(0);  // ticket + barrier
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t valid = min((uint32_t)TILE, n - tile_base);
    const bool full = (valid == (uint32_t)TILE);

    // ---- load: warp-striped so that (item, lane) order == memory order (stability).  Full tiles
    //      arrive as TMA bulk copies (keys and payload in flight together, one mbarrier); the tile the
    //      CTA after next on this SM will want is pulled into L2 meanwhile ----
    KeyT key[IPT];
    const uint32_t warp_base = warp * 32u * IPT;
    if (TMA && full) {
        if (tid == 0) {
            mbar_expect_tx(&s.mbar, TILE * (sizeof(KeyT) + (PAIRS ? 4u : 0u)));
            tma_bulk_g2s(s.keys, keys_in + tile_base, TILE * sizeof(KeyT), &s.mbar);
            if (PAIRS) tma_bulk_g2s(s.vals_in, vals_in + tile_base, TILE * 4u, &s.mbar);
            const uint32_t ahead = tile + kPrefetchTiles;
            if (kPrefetchTiles > 0 && (uint64_t)(ahead + 1u) * TILE <= n) {
                l2_prefetch_bulk(keys_in + (size_t)ahead * TILE, TILE * sizeof(KeyT));
                if (PAIRS) l2_prefetch_bulk(vals_in + (size_t)ahead * TILE, TILE * 4u);
            }
        }
        mbar_wait(&s.mbar, 0);
#pragma unroll
        for (int k = 0; k < IPT; ++k) key[k] = s.keys[warp_base + k * 32 + lane];
    } else {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const uint32_t i = warp_base + k * 32 + lane;
            key[k] = (i < valid) ? keys_in[tile_base + i] : ~KeyT(0);  // pad sorts last, never written
            if (PAIRS) s.vals_in[i] = (i < valid) ? vals_in[tile_base + i] : 0u;  // same thread reads it back
        }
    }

    RTR_PHASE(1);  // tile load (TMA arrival, keys to registers)
    // ---- digit counts of the tile first (one shared-memory reduction per key): the aggregate other tiles
    //      look back at is published BEFORE the expensive ranking, and this tile's own look-back runs
    //      interleaved with the ranking instead of after it ----
#pragma unroll
    for (int k = 0; k < IPT; ++k) atomicAdd(&s.cta_hist[digit_of<KeyT>(key[k], shift, mask)], 1u);
    __syncthreads();
    uint32_t total = 0, x = 0;
    if (tid < kRadix) {
        total = s.cta_hist[tid];
        st_relaxed_u32(&status[(size_t)tile * kRadix + tid], (tile == 0 ? kFlagIncl : kFlagAgg) | total);
        x = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (lane == 31) s.scan_warp[warp] = x;
    }
    // decoupled look-back, one thread per digit, kLookBatch predecessors in flight; issued/consumed in steps
    uint32_t before = 0;
    int lb_t = (int)tile - 1;
    bool lb_done = !(tid < kRadix) || tile == 0;
    uint32_t lb_v[kLookBatch];
    auto lb_issue = [&]() {
#pragma unroll
        for (int j = 0; j < kLookBatch; ++j)
            lb_v[j] = (lb_t - j >= 0) ? ld_relaxed_u32(&status[(size_t)(lb_t - j) * kRadix + tid]) : kFlagIncl;
    };
    auto lb_consume = [&]() {
        int adv = 0;
#pragma unroll
        for (int j = 0; j < kLookBatch; ++j) {
            if (!lb_done && adv == j) {
                const uint32_t flag = lb_v[j] & ~kValueMask;
                if (flag != 0) {  // published
                    before += lb_v[j] & kValueMask;
                    ++adv;
                    if (flag == kFlagIncl) lb_done = true;
                }
            }
        }
        lb_t -= adv;
        if (lb_done) st_relaxed_u32(&status[(size_t)tile * kRadix + tid], kFlagIncl | (before + total));
    };

    // ---- rank inside the warp.  peers = lanes with the same digit; the key's slot among the warp's keys of
    //      that digit = running count of the earlier items (shared memory, bumped by the highest peer lane)
    //      + peers below.  rank[k] ends as (digit << 16 | slot) ----
    RTR_PHASE(2);  // tile histogram, aggregate published, first look-back loads
    uint32_t rank[IPT];
    uint32_t* wh = s.whist[warp];
    const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
#if RTR_SORT_RANK_ATOMIC
    // The lowest peer lane adds the group's size to the warp's running count of the digit with ONE shared-memory
    // atomic and hands the old value to its peers by shuffle.  Unlike "read the count, let the last peer store it
    // back" the items do not wait for each other: the atomics of a group of items are all in flight before the first
    // result is consumed (same-address atomics of one warp are applied in issue order, __syncwarp orders the items).
    constexpr int G = RTR_SORT_RANK_GROUP;
    static_assert(IPT % G == 0, "ranking group");
#pragma unroll
    for (int k0 = 0; k0 < IPT; k0 += G) {
        const bool lb_step = !lb_done;
        if (lb_step) lb_issue();
        uint32_t base[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t d = digit_of<KeyT>(key[k0 + j], shift, mask);
            const uint32_t peers = digit_peers8(d);  // bits above a last partial digit are 0 in every lane: their ballots narrow nothing
            const uint32_t below = __popc(peers & lt);
            base[j] = 0u;
            if (below == 0u) base[j] = atomicAdd(&wh[d], (uint32_t)__popc(peers));
            __syncwarp();
            rank[k0 + j] = (d << 16) | (below << 8) | (uint32_t)(__ffs(peers) - 1);
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t r = rank[k0 + j];
            const uint32_t b = __shfl_sync(0xffffffffu, base[j], r & 31u);
            rank[k0 + j] = (r & 0xFFFF0000u) | (b + ((r >> 8) & 0xFFu));
        }
        if (lb_step) lb_consume();
    }
#else
#pragma unroll
    for (int k0 = 0; k0 < IPT; k0 += 4) {
        const bool lb_step = !lb_done;
        if (lb_step) lb_issue();
        uint32_t peers[4], dg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dg[j] = digit_of<KeyT>(key[k0 + j], shift, mask);
            peers[j] = digit_peers8(dg[j]);  // bits above a last partial digit are 0 in every lane: their ballots narrow nothing
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t below = __popc(peers[j] & lt);
            const uint32_t cnt = wh[dg[j]];                              // every peer reads the same word ...
            if ((peers[j] & gt) == 0u) wh[dg[j]] = cnt + below + 1u;     // ... before the last one bumps it
            __syncwarp();                                                // item k's bump is ordered before item k+1's read
            rank[k0 + j] = (dg[j] << 16) | (cnt + below);
        }
        if (lb_step) lb_consume();
    }
#endif
    RTR_PHASE(3);  // ranking (interleaved look-back steps)
    while (!lb_done) { lb_issue(); lb_consume(); }
    RTR_PHASE(4);  // rest of the look-back
    __syncthreads();  // all keys are in registers; s.keys may be overwritten from here on

    // ---- per digit (threads 0..255): first tile-local slot of every (warp, digit) ----
    if (tid < kRadix) {
        uint32_t cta_ofs = x - total;
        for (uint32_t i = 0; i < warp; ++i) cta_ofs += s.scan_warp[i];
        uint32_t run = cta_ofs;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { const uint32_t c = s.whist[w][tid]; s.whist[w][tid] = run; run += c; }
        s.gofs[tid] = gbase[tid] + before - cta_ofs;
    }
    __syncthreads();

    RTR_PHASE(5);  // barrier + slots of (warp, digit)
    // ---- reorder keys (and payload) inside the tile through shared memory ----
    uint2* kv = reinterpret_cast<uint2*>(s.keys);
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        const uint32_t pos = wh[rank[k] >> 16] + (rank[k] & 0xFFFFu);
        if constexpr (PACKED) {
            kv[pos] = make_uint2((uint32_t)key[k], s.vals_in[warp_base + k * 32 + lane]);
        } else {
            s.keys[pos] = key[k];
            if (PAIRS) s.vals[pos] = s.vals_in[warp_base + k * 32 + lane];
        }
    }
    __syncthreads();

    RTR_PHASE(6);  // reorder through shared memory + barrier
    // ---- coalesced scatter: consecutive threads write runs of equal digits ----
    if (full) {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const uint32_t i = k * BLOCK + tid;
            if constexpr (PACKED) {
                const uint2 e = kv[i];
                const uint32_t dst = s.gofs[digit_of<KeyT>((KeyT)e.x, shift, mask)] + i;
                keys_out[dst] = (KeyT)e.x;
                vals_out[dst] = e.y;
            } else {
                const KeyT kk = s.keys[i];
                const uint32_t dst = s.gofs[digit_of<KeyT>(kk, shift, mask)] + i;
                keys_out[dst] = kk;
                if (PAIRS) vals_out[dst] = s.vals[i];
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const uint32_t i = k * BLOCK + tid;
            if (i < valid) {
                if constexpr (PACKED) {
                    const uint2 e = kv[i];
                    const uint32_t dst = s.gofs[digit_of<KeyT>((KeyT)e.x, shift, mask)] + i;
                    keys_out[dst] = (KeyT)e.x;
                    vals_out[dst] = e.y;
                } else {
                    const KeyT kk = s.keys[i];
                    const uint32_t dst = s.gofs[digit_of<KeyT>(kk, shift, mask)] + i;
                    keys_out[dst] = kk;
                    if (PAIRS) vals_out[dst] = s.vals[i];
                }
            }
        }
    }
    RTR_PHASE(7);  // scatter issued
}

// ---------------------------------------------------------------------------------------
// testsSortGPU pre-passes
// ---------------------------------------------------------------------------------------
// histogramOfGlobalDigitCounts.glsl:31-59 : per-bit popcount histogram.  Lane b of every warp
// owns bit b: 32 ballots per 32 keys, no shared atomics in the loop.
__global__ void __launch_bounds__(256)
bit_histogram32_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ out) {
    __shared__ uint32_t s_hist[32];
    if (threadIdx.x < 32) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t lane = lane_id();
    uint32_t acc = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (n + stride - 1) / stride;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t i = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const uint32_t k = (i < n) ? __ldg(keys + i) : 0u;
#pragma unroll
        for (uint32_t b = 0; b < 32; ++b) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (k >> b) & 1u);
            if (lane == b) acc += __popc(bal);
        }
    }
    if (acc) atomicAdd(&s_hist[lane], acc);
    __syncthreads();
    if (threadIdx.x < 32 && s_hist[threadIdx.x]) atomicAdd(&out[threadIdx.x], s_hist[threadIdx.x]);
}

// prefixSumOfGlobalDigitCounts.glsl:24-69 : exclusive scan inside each group of 4 bins (one warp)
__global__ void digitplace_scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
    const uint32_t l = threadIdx.x;
    const uint32_t v = in[l];
    uint32_t x = v;
    uint32_t y = __shfl_up_sync(0xffffffffu, x, 1, 4);
    if ((l & 3u) >= 1u) x += y;
    y = __shfl_up_sync(0xffffffffu, x, 2, 4);
    if ((l & 3u) >= 2u) x += y;
    out[l] = x - v;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <typename KeyT> struct SortCfg;
#ifndef RTR_SORT_BLOCK
#define RTR_SORT_BLOCK 512
#endif
#ifndef RTR_SORT_IPT
#define RTR_SORT_IPT 12
#endif
template <> struct SortCfg<uint32_t> { static constexpr int BLOCK = RTR_SORT_BLOCK, IPT = RTR_SORT_IPT, MAX_PASSES = 4; };
template <> struct SortCfg<uint64_t> { static constexpr int BLOCK = RTR_SORT_BLOCK, IPT = 8, MAX_PASSES = 8; };

struct SortWs {
    void* keys_tmp;
    uint32_t* vals_tmp;
    uint32_t* ghist;     // [MAX_PASSES][256]
    uint32_t* gbase;     // [MAX_PASSES][256]
    uint32_t* counters;  // [MAX_PASSES]
    uint32_t* status;    // [passes][tiles][256]
    size_t zero_bytes;   // ghist..status are contiguous and zeroed together
    void* zero_begin;
};

template <typename KeyT>
size_t sort_ws_layout(void* base, uint32_t n, bool pairs, SortWs* out) {
    using Cfg = SortCfg<KeyT>;
    const uint32_t tile = Cfg::BLOCK * Cfg::IPT;
    const uint32_t tiles = (n + tile - 1) / tile;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = rtr_align_up(off + bytes, 256); return o; };
    const size_t o_keys = take((size_t)n * sizeof(KeyT));
    const size_t o_vals = take(pairs ? (size_t)n * 4 : 0);
    const size_t o_hist = take((size_t)Cfg::MAX_PASSES * kRadix * 4);
    const size_t o_base = take((size_t)Cfg::MAX_PASSES * kRadix * 4);
    const size_t o_cnt = take((size_t)Cfg::MAX_PASSES * 4);
    const size_t o_status = take((size_t)Cfg::MAX_PASSES * tiles * kRadix * 4);
    if (out) {
        char* b = static_cast<char*>(base);
        out->keys_tmp = b + o_keys;
        out->vals_tmp = reinterpret_cast<uint32_t*>(b + o_vals);
        out->ghist = reinterpret_cast<uint32_t*>(b + o_hist);
        out->gbase = reinterpret_cast<uint32_t*>(b + o_base);
        out->counters = reinterpret_cast<uint32_t*>(b + o_cnt);
        out->status = reinterpret_cast<uint32_t*>(b + o_status);
        out->zero_begin = b + o_hist;
        out->zero_bytes = off - o_hist;
    }
    return off;
}

template <typename KeyT, bool PAIRS, bool TMA>
int launch_pass(rtr_ctx* ctx, const KeyT* kin, KeyT* kout, const uint32_t* vin, uint32_t* vout, uint32_t n,
                uint32_t tiles, int shift, uint32_t mask, const uint32_t* gbase, uint32_t* status, uint32_t* counter) {
    using Cfg = SortCfg<KeyT>;
    auto kern = onesweep_kernel<KeyT, PAIRS, Cfg::BLOCK, Cfg::IPT, TMA>;
    const size_t smem = sizeof(OnesweepSmem<KeyT, PAIRS, Cfg::BLOCK, Cfg::IPT>);
    // the opt-in is per device and a process may hold contexts on several: set it every time (a cheap host-side call)
    RTR_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RTR_PROF(ctx, sizeof(KeyT) == 4 ? (PAIRS ? "onesweep_kernel<u32,pairs>" : "onesweep_kernel<u32,keys>")
                                    : (PAIRS ? "onesweep_kernel<u64,pairs>" : "onesweep_kernel<u64,keys>"));
    kern<<<tiles, Cfg::BLOCK, smem, ctx->stream>>>(kin, kout, vin, vout, n, tiles, shift, mask, gbase, status, counter);
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}

template <typename KeyT>
int sort_impl(rtr_ctx* ctx, KeyT* keys, uint32_t* vals, uint32_t n, int begin_bit, int end_bit) {
    using Cfg = SortCfg<KeyT>;
    const int key_bits = (int)sizeof(KeyT) * 8;
    if (begin_bit < 0 || end_bit > key_bits || begin_bit >= end_bit)
        return rtr_set_error(ctx, RTR_E_INVALID, "sort: bad bit range [%d,%d)", begin_bit, end_bit);
    if (n >= (1u << 30)) return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "sort: n=%u >= 2^30", n);
    if (n <= 1) return RTR_OK;
    const bool pairs = vals != nullptr;
    const int passes = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
    const uint32_t tile = Cfg::BLOCK * Cfg::IPT;
    const uint32_t tiles = (n + tile - 1) / tile;

    const size_t need = sort_ws_layout<KeyT>(nullptr, n, pairs, nullptr);
    RTR_CHECK(rtr_ws_reserve(ctx, need));
    SortWs ws;
    sort_ws_layout<KeyT>(ctx->ws, n, pairs, &ws);

    RTR_CUDA(ctx, cudaMemsetAsync(ws.zero_begin, 0, ws.zero_bytes, ctx->stream));
    {
        uint32_t blocks = (n / 4 + 511) / 512;
        const uint32_t cap = (uint32_t)ctx->sm_count * 4;
        blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
        RTR_PROF(ctx, "radix_histogram_kernel");
        radix_histogram_kernel<KeyT, Cfg::MAX_PASSES><<<blocks, 512, 0, ctx->stream>>>(keys, n, begin_bit, end_bit, passes, ws.ghist);
        RTR_LAUNCH_CHECK(ctx);
        radix_scan_kernel<<<passes, kRadix, 0, ctx->stream>>>(ws.ghist, ws.gbase);
        RTR_LAUNCH_CHECK(ctx);
    }
    KeyT* kin = keys;
    KeyT* kout = static_cast<KeyT*>(ws.keys_tmp);
    uint32_t* vin = vals;
    uint32_t* vout = pairs ? ws.vals_tmp : nullptr;
    for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + p * kRadixBits;
        const int bits = (end_bit - shift) < kRadixBits ? (end_bit - shift) : kRadixBits;
        const uint32_t mask = (1u << bits) - 1u;
        // cp.async.bulk needs 16-byte aligned global addresses -- for the payload as well as for the keys
        const bool tma = (reinterpret_cast<uintptr_t>(kin) & 15u) == 0 && (!pairs || (reinterpret_cast<uintptr_t>(vin) & 15u) == 0);
        uint32_t* status = ws.status + (size_t)p * tiles * kRadix;
        int r;
        if (pairs) r = tma ? launch_pass<KeyT, true, true>(ctx, kin, kout, vin, vout, n, tiles, shift, mask, ws.gbase + p * kRadix, status, ws.counters + p)
                           : launch_pass<KeyT, true, false>(ctx, kin, kout, vin, vout, n, tiles, shift, mask, ws.gbase + p * kRadix, status, ws.counters + p);
        else r = tma ? launch_pass<KeyT, false, true>(ctx, kin, kout, vin, vout, n, tiles, shift, mask, ws.gbase + p * kRadix, status, ws.counters + p)
                     : launch_pass<KeyT, false, false>(ctx, kin, kout, vin, vout, n, tiles, shift, mask, ws.gbase + p * kRadix, status, ws.counters + p);
        RTR_CHECK(r);
        KeyT* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {  // odd number of passes: result sits in the workspace
        RTR_CUDA(ctx, cudaMemcpyAsync(keys, kin, (size_t)n * sizeof(KeyT), cudaMemcpyDeviceToDevice, ctx->stream));
        if (pairs) RTR_CUDA(ctx, cudaMemcpyAsync(vals, vin, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return RTR_OK;
}

}  // namespace

size_t rtr_sort_ws_bytes(uint32_t n, int key_bytes, bool pairs) {
    return key_bytes == 8 ? sort_ws_layout<uint64_t>(nullptr, n, pairs, nullptr)
                          : sort_ws_layout<uint32_t>(nullptr, n, pairs, nullptr);
}

int rtr_sort_impl_u32(rtr_ctx* ctx, uint32_t* keys, uint32_t* vals, uint32_t n, int begin_bit, int end_bit) {
    return sort_impl<uint32_t>(ctx, keys, vals, n, begin_bit, end_bit);
}
int rtr_sort_impl_u64(rtr_ctx* ctx, uint64_t* keys, uint32_t* vals, uint32_t n, int begin_bit, int end_bit) {
    return sort_impl<uint64_t>(ctx, keys, vals, n, begin_bit, end_bit);
}

#ifdef RTR_SORT_PHASE_CLOCKS
extern "C" int rtr_debug_sort_phases(unsigned long long out[8], int reset) {
    if (out && cudaMemcpyFromSymbol(out, g_sort_phase_ns, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_sort_phase_ns, z, sizeof(z)); }
    return 0;
}
#endif

int rtr_bit_histogram32_launch(rtr_ctx* ctx, const uint32_t* keys, uint32_t n, uint32_t* out) {
    uint32_t blocks = (n + 2047) / 2048;  // the reference dispatches n/2048 work-groups of 2048 keys
    const uint32_t cap = (uint32_t)ctx->sm_count * 8;
    blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
    bit_histogram32_kernel<<<blocks, 256, 0, ctx->stream>>>(keys, n, out);
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}

int rtr_digitplace_scan_launch(rtr_ctx* ctx, const uint32_t* in, uint32_t* out) {
    digitplace_scan_kernel<<<1, 32, 0, ctx->stream>>>(in, out);
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}
