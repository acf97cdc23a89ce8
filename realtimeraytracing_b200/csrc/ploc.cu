// ploc.cu -- PLOC bottom-up BVH construction and DFS flattening, bit-exact with the serial
// (OMP_NUM_THREADS=1) reference.
//
// Replaces BVH::plocPreprocessing (bvh.cpp:26-46), BVH::ploc (:48-123) with its four phases
// plocNearestNeighborSearch (:193-210), plocMerging (:159-191), plocPrefixScan (:125-148),
// plocCompaction (:150-156), and glr::Scene::getBVH_NodesToGPUData (scene.cpp:189-208).
//
// Layout in HBM (by cluster id, one 32-byte record = one sector, read with a single 256-bit load):
//   node[2c] = (min.x, min.y, min.z, max.x)      node[2c+1] = (max.y, max.z, L, R)
//   leaves c < n:  L = triangle id, R = RTR_NONE;  internal c >= n: L/R = child cluster ids
//   isize[c-n] = nodes in the subtree of internal cluster c (lets flatten place the right child)
//   cin/cout   = active cluster ids in Morton order (the reference's C_In/C_Out)
//
// One kernel per PLOC iteration (ploc_iteration_kernel), NN search + merge + prefix scan + compaction
// of the reference fused into one pass over the active list.  A CTA stages the ids and boxes of its tile
// plus a 32-position halo on each side in shared memory (float4 + float2 per box, interleaved so that no
// access pattern below has bank conflicts); every thread owns 4 consecutive positions whose boxes it keeps
// in registers.  The merged surface area is symmetric, so each pair inside the +-16 window is evaluated
// once, by the thread owning its lower position; the other end receives the column minima by warp shuffle
// (see the kernel).  The result is the reference's argmin in ascending j with strict '<', i.e. lowest j on
// ties (bvh.cpp:199-208), every fp32 op individually rounded in the reference's order.  Mutual pairs are
// ranked with a CTA scan + decoupled look-back across tiles: merged nodes go to id = total + rank (== the
// reference's serial counter, Q3) and survivors are compacted in place order.  The iteration count is data
// dependent, so the loop state lives on the device (PlocState, double buffered by launch parity) and
// launches that find n_active <= kTailN are no-ops; a single-CTA tail kernel finishes the last <= 1024
// clusters entirely in shared memory (PLOC++ 4.4).
#include "bvh.cuh"

namespace {

constexpr int kR = RTR_MAX_SEARCH_RADIUS;       // 16
constexpr int kPP = 4;                           // consecutive positions per thread
constexpr int kHaloLanes = kR / kPP;             // lanes 0..3 of a warp repeat the last 16 positions of the warp before
constexpr int kWarps = RTR_PLOC_WARPS;
#ifndef RTR_PLOC_ROLL
#define RTR_PLOC_ROLL 0   // 1: the three full 4x4 blocks of the pair matrix run as a loop (a third of the code)
#endif
#ifndef RTR_PLOC_MINB
#define RTR_PLOC_MINB (1024 / (32 * RTR_PLOC_WARPS))  // resident CTAs per SM the register budget is cut for (64 registers)
#endif
constexpr int kPT = 32 * kWarps;                 // threads per CTA
constexpr int kWarpSpan = (32 - kHaloLanes) * kPP;  // 112 positions per warp get their nearest neighbour
constexpr int kEval = kWarpSpan * kWarps;        // positions whose nearest neighbour a CTA computes: staged [kR, kR + kEval)
constexpr int kTileT = kPlocTile;                // positions decided per CTA: staged [2 kR, kEval)
constexpr int kStage = kEval + 2 * kR;           // positions staged (tile + 2 kR halo each side)
constexpr int kStageQ = kStage / kPP;            // staged arrays are stored [e % kPP][e / kPP] (conflict free)
constexpr int kStagePT = (kStage + kPT - 1) / kPT;  // staged positions gathered per thread
constexpr uint32_t kTailN = 1024;
static_assert(kR % kPP == 0 && kR == 16 && kPP == 4, "search geometry is written for +-16 and 4 positions per thread");
static_assert(kTileT == kEval - 2 * kR, "tile geometry");
static_assert(kStage % kPP == 0, "staging layout");

// look-back word: [0,27) merges (lo partners), [27,54) removed (hi partners), [54,56) flag, [56,64) tag
constexpr uint64_t kCntMask = (1ull << 27) - 1;
constexpr uint64_t kLbAgg = 1ull, kLbIncl = 2ull;
__device__ __forceinline__ uint64_t lb_pack(uint32_t tag, uint64_t flag, uint32_t lo, uint32_t hi) {
    return (uint64_t)lo | ((uint64_t)hi << 27) | (flag << 54) | ((uint64_t)(tag & 0xFFu) << 56);
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// one 32-byte cluster record
struct Box {
    float4 lo;  // min.xyz, max.x
    float4 hi;  // max.y, max.z, bits(L), bits(R)
};
__device__ __forceinline__ Box load_box(const float4* __restrict__ node, uint32_t c) {
    Box b;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(b.lo.x), "=f"(b.lo.y), "=f"(b.lo.z), "=f"(b.lo.w), "=f"(b.hi.x), "=f"(b.hi.y), "=f"(b.hi.z), "=f"(b.hi.w)
                 : "l"(node + 2 * (size_t)c));
    return b;
}
__device__ __forceinline__ void store_box(float4* __restrict__ node, uint32_t c, const float4 lo, const float4 hi) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(node + 2 * (size_t)c), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w)
                 : "memory");
}

struct BoxSoA {  // views into shared memory (tail kernel)
    float* minx; float* miny; float* minz; float* maxx; float* maxy; float* maxz;
};

// AABB::merge (bvh.cpp:401-413) + AABB::getSurfaceArea (:378-381) without the final "2 *", which
// is exact in binary fp and therefore cannot change any comparison: (dx*dy + dy*dz) + dz*dx
__device__ __forceinline__ float merged_half_area(const BoxSoA& b, int i, int j) {
    const float dx = __fsub_rn(fmaxf(b.maxx[i], b.maxx[j]), fminf(b.minx[i], b.minx[j]));
    const float dy = __fsub_rn(fmaxf(b.maxy[i], b.maxy[j]), fminf(b.miny[i], b.miny[j]));
    const float dz = __fsub_rn(fmaxf(b.maxz[i], b.maxz[j]), fminf(b.minz[i], b.minz[j]));
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dy, dz)), __fmul_rn(dz, dx));
}
// the same on register boxes: lo = (min.xyz, max.x), hi = (max.y, max.z)
__device__ __forceinline__ float pair_half_area(const float4 alo, const float2 ahi, const float4 blo, const float2 bhi) {
    const float dx = __fsub_rn(fmaxf(alo.w, blo.w), fminf(alo.x, blo.x));
    const float dy = __fsub_rn(fmaxf(ahi.x, bhi.x), fminf(alo.y, blo.y));
    const float dz = __fsub_rn(fmaxf(ahi.y, bhi.y), fminf(alo.z, blo.z));
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dy, dz)), __fmul_rn(dz, dx));
}

// plocNearestNeighborSearch (bvh.cpp:193-210): ascending j, strict '<' => lowest j wins ties.
// e_q/e_lo/e_hi are shared-memory slots; returns the slot of the neighbour or -1 (all inf/NaN, Q4).
__device__ __forceinline__ int nearest_neighbour(const BoxSoA& b, int e_q, int e_lo, int e_hi) {
    float best = INFINITY;
    int best_e = -1;
    for (int e = e_lo; e < e_hi; ++e) {
        if (e == e_q) continue;
        const float a = merged_half_area(b, e_q, e);
        if (a < best) { best = a; best_e = e; }
    }
    return best_e;
}

// inclusive scan of a packed (lo | hi << 16) flag pair over the CTA; returns inclusive value,
// *total = CTA sum.  s_warp needs BLOCK/32 + 1 words.
template <int BLOCK>
__device__ __forceinline__ uint32_t block_scan_incl(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (int i = 0; i < BLOCK / 32; ++i) {
        const uint32_t w = s_warp[i];
        if ((uint32_t)i < warp) base += w;
        sum += w;
    }
    *total = sum;
    __syncthreads();  // s_warp may be reused by the caller
    return base + x;
}

// Longest squared triangle edge of the launch (conservative traversal pruning bound, see trace.cu): warp
// shuffle -> CTA -> at most one atomic per CTA, and only while the CTA still raises the global value
// (one same-address atomic per warp kept leaf_init_kernel waiting on a single L2 slice).
__device__ __forceinline__ void publish_edge_bound(float e2, TraceParams* __restrict__ tparams) {
    __shared__ float s_e2[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, o));
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, warps = (blockDim.x + 31u) >> 5;
    if (lane == 0) s_e2[warp] = e2;
    __syncthreads();
    if (warp == 0) {
        e2 = lane < warps ? s_e2[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, o));
        if (lane == 0 && e2 > 0.f) {
            const uint32_t mine = float_to_ordered(e2);
            if (mine > ld_relaxed_u32(&tparams->emax2_ordered)) atomicMax(&tparams->emax2_ordered, mine);
        }
    }
}

// ---------------------------------------------------------------------------------------
// leaf initialisation: plocPreprocessing (bvh.cpp:29-42) + AABB::buildFromTriangle (:383-399)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
leaf_init_kernel(const rtr_triangle* __restrict__ tris, const rtr_mesh* __restrict__ meshes,
                 const uint32_t* __restrict__ tri_idx, uint32_t n,
                 float4* __restrict__ node, uint32_t* __restrict__ cin,
                 float4* __restrict__ wtri, TraceParams* __restrict__ tparams) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float e2 = 0.f;
    if (i < n) {
        const uint32_t ti = tri_idx[i];
        const TriRec t = load_tri(tris, ti);
        const Mat3x4 M = load_model(meshes, t.model_id);
        const float3 a = mat_mul_point(M, t.p0), b = mat_mul_point(M, t.p1), c = mat_mul_point(M, t.p2);
        const float mnx = fminf(a.x, fminf(b.x, c.x)), mny = fminf(a.y, fminf(b.y, c.y)), mnz = fminf(a.z, fminf(b.z, c.z));
        const float mxx = fmaxf(a.x, fmaxf(b.x, c.x)), mxy = fmaxf(a.y, fmaxf(b.y, c.y)), mxz = fmaxf(a.z, fmaxf(b.z, c.z));
        store_box(node, i, make_float4(mnx, mny, mnz, mxx),
                  make_float4(mxy, mxz, __uint_as_float(ti), __uint_as_float(RTR_NONE)));
        cin[i] = i;
        // world-space vertices for the traversal kernels (the M*P of raytracer.glsl:105-107), leaf order
        wtri[3 * (size_t)i + 0] = make_float4(a.x, a.y, a.z, b.x);
        wtri[3 * (size_t)i + 1] = make_float4(b.y, b.z, c.x, c.y);
        wtri[3 * (size_t)i + 2] = make_float4(c.z, 0.f, 0.f, 0.f);
        // longest squared edge (conservative traversal pruning bound, see trace.cu)
        const float abx = b.x - a.x, aby = b.y - a.y, abz = b.z - a.z;
        const float acx = c.x - a.x, acy = c.y - a.y, acz = c.z - a.z;
        const float bcx = c.x - b.x, bcy = c.y - b.y, bcz = c.z - b.z;
        e2 = fmaxf(abx * abx + aby * aby + abz * abz,
                   fmaxf(acx * acx + acy * acy + acz * acz, bcx * bcx + bcy * bcy + bcz * bcz));
    }
    publish_edge_bound(e2, tparams);
}

// adopted node arrays: same cache, indexed by triangle id
__global__ void __launch_bounds__(256)
edge_bound_kernel(const rtr_triangle* __restrict__ tris, const rtr_mesh* __restrict__ meshes, uint32_t n,
                  float4* __restrict__ wtri, TraceParams* __restrict__ tparams) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float e2 = 0.f;
    if (i < n) {
        const TriRec t = load_tri(tris, i);
        const Mat3x4 M = load_model(meshes, t.model_id);
        const float3 a = mat_mul_point(M, t.p0), b = mat_mul_point(M, t.p1), c = mat_mul_point(M, t.p2);
        wtri[3 * (size_t)i + 0] = make_float4(a.x, a.y, a.z, b.x);
        wtri[3 * (size_t)i + 1] = make_float4(b.y, b.z, c.x, c.y);
        wtri[3 * (size_t)i + 2] = make_float4(c.z, 0.f, 0.f, 0.f);
        const float abx = b.x - a.x, aby = b.y - a.y, abz = b.z - a.z;
        const float acx = c.x - a.x, acy = c.y - a.y, acz = c.z - a.z;
        const float bcx = c.x - b.x, bcy = c.y - b.y, bcz = c.z - b.z;
        e2 = fmaxf(abx * abx + aby * aby + abz * abz,
                   fmaxf(acx * acx + acy * acy + acz * acz, bcx * bcx + bcy * bcy + bcz * bcz));
    }
    publish_edge_bound(e2, tparams);
}

// state[0/1]: the loop state, double buffered by iteration parity; state[2]: where the loop kernel leaves off
// and the tail kernel finishes (what the flatten and rtr_bvh_finish read); ctl[0/1]: barrier counters of the
// two persistent kernels
__global__ void ploc_state_init_kernel(PlocState* state, uint32_t n, uint32_t* iter_first_id, uint32_t* ctl) {
    state[0].n_active = n; state[0].total = n; state[0].iter = 0; state[0].tile_counter = 0;
    state[1] = state[0];
    state[2] = state[0];
    iter_first_id[0] = n;
    ctl[0] = ctl[1] = ctl[2] = ctl[3] = 0u;
}

// ---------------------------------------------------------------------------------------
// one PLOC iteration over the whole active list
// ---------------------------------------------------------------------------------------
struct IterSmem {
    float4 lo[kStage];            // [e % kPP][e / kPP]: min.xyz, max.x
    float2 hi[kStage];            //                     max.y, max.z
    uint32_t id[kStage];          //                     cluster id (RTR_NONE outside the active list)
    int nn[kEval];                // neighbour (staged index) of staged position kR + q, or -1
    uint32_t warp[kPT / 32 + 1];
    uint32_t tile, ex_lo, ex_hi;
};
__device__ __forceinline__ int perm_stage(int e) { return (e % kPP) * kStageQ + (e / kPP); }

// One block of the pair matrix: own positions e0+a (a = 0..3) x the positions e0 + 4 dl + x of lane + dl.
// LAST_LANE_BLOCK (dl = 4): only the pairs at distance 16 + x - a <= 16 exist.
template <bool FULL, bool LAST_LANE_BLOCK, typename Smem>
__device__ __forceinline__ void search_block(const Smem& s, int dl, int e0, int radius,
                                             const float4 (&lo)[kPP], const float2 (&hi)[kPP],
                                             float (&fbest)[kPP], int (&fj)[kPP], float (&bbest)[kPP], int (&bj)[kPP]) {
    float4 xlo[kPP]; float2 xhi[kPP];
    const int ex = e0 + kPP * dl;
#pragma unroll
    for (int x = 0; x < kPP; ++x) {
        const int pe = x * kStageQ + ex / kPP;  // perm_stage(ex + x): ex is a multiple of kPP
        xlo[x] = s.lo[pe];
        xhi[x] = s.hi[pe];
    }
    float cv[kPP]; int ca[kPP];
#pragma unroll
    for (int x = kPP - 1; x >= 0; --x) {
        cv[x] = INFINITY; ca[x] = 0;
#pragma unroll
        for (int a = 0; a < kPP; ++a) {
            if (!LAST_LANE_BLOCK || x <= a) {
                float d = pair_half_area(lo[a], hi[a], xlo[x], xhi[x]);
                if (!FULL) d = (kPP * dl + x - a <= radius) ? d : INFINITY;
                if (d <= fbest[a]) { fbest[a] = d; fj[a] = ex + x; }   // candidates above come in descending j
                if (d < cv[x]) { cv[x] = d; ca[x] = a; }               // ascending a: lowest a among equal minima
            }
        }
    }
#pragma unroll
    for (int x = 0; x < kPP; ++x) {  // candidates below come in ascending j (dl descending)
        const float rv = __shfl_up_sync(0xffffffffu, cv[x], dl);
        const int ra = __shfl_up_sync(0xffffffffu, ca[x], dl);
        if (rv < bbest[x]) { bbest[x] = rv; bj[x] = e0 - kPP * dl + ra; }
    }
}

// FULL: radius == 16, the reference's PlocParams::_SEARCH_RADIUS (bvh.hpp:66); otherwise candidates farther
// than `radius` positions away are masked out.
//
// One tile of one iteration: the positions [tile * kTileT, (tile + 1) * kTileT) of the active list are decided.
// `cur` state values (n, total, iter) are passed in; the tile that closes the look-back chain writes `nxt`.
template <bool FULL>
__device__ __forceinline__ void ploc_process_tile(IterSmem& s, const uint32_t tile, const uint32_t tiles,
                                                  const uint32_t n, const uint32_t total, const uint32_t iter,
                                                  const uint32_t n_leaves, const int radius,
                                                  const uint32_t* __restrict__ cin, uint32_t* __restrict__ cout,
                                                  float4* __restrict__ node, uint2* __restrict__ isize,
                                                  PlocState* __restrict__ nxt, uint64_t* __restrict__ tile_status,
                                                  uint32_t* __restrict__ trace_active, uint32_t* __restrict__ trace_merges,
                                                  uint32_t* __restrict__ iter_first_id) {
    const int tid = (int)threadIdx.x;
    const int t0 = (int)(tile * kTileT);
    const int base = t0 - 2 * kR;  // position of staged slot 0

    // ---- stage ids + boxes of the tile and its halo (one 32-byte gather per position; all ids first, then
    //      all boxes, so that a thread has its kStagePT gathers in flight together).  Slots outside the
    //      active list get the box (-inf, +inf): every merged area with it is +inf, which never wins ----
    {
        uint32_t sid[kStagePT];
#pragma unroll
        for (int r = 0; r < kStagePT; ++r) {
            const int e = tid + r * kPT, pos = base + e;
            sid[r] = (e < kStage && pos >= 0 && pos < (int)n) ? __ldcg(cin + pos) : RTR_NONE;
        }
        Box sb[kStagePT];
#pragma unroll
        for (int r = 0; r < kStagePT; ++r) {
            if (sid[r] != RTR_NONE) sb[r] = load_box(node, sid[r]);
            else { sb[r].lo = make_float4(-INFINITY, -INFINITY, -INFINITY, INFINITY); sb[r].hi = make_float4(INFINITY, INFINITY, 0.f, 0.f); }
        }
#pragma unroll
        for (int r = 0; r < kStagePT; ++r) {
            const int e = tid + r * kPT;
            if (e < kStage) {
                const int pe = perm_stage(e);
                s.id[pe] = sid[r];
                s.lo[pe] = sb[r].lo;
                s.hi[pe] = make_float2(sb[r].hi.x, sb[r].hi.y);
            }
        }
    }
    __syncthreads();

    // ---- nearest neighbour (plocNearestNeighborSearch, bvh.cpp:193-210) of the kPP positions this thread
    //      owns, staged e0 .. e0+3.  The merged area is symmetric, so every pair is evaluated ONCE, by the
    //      thread that owns its lower position: block dl = own 4 positions x the 4 positions of lane + dl
    //      (dl = 4..1; dl = 4 only holds the 10 pairs at distance <= 16).  Row minima serve the own
    //      positions (their candidates above), column minima go to lane + dl by shuffle (its candidates
    //      below).  The reference scans j ascending with strict '<', i.e. it returns the LOWEST j among the
    //      minima: candidates above arrive in descending j and replace on '<=', candidates below arrive in
    //      ascending j and replace on '<', and the lower side wins a tie between the two.  Lanes 0..3 of a
    //      warp own the same positions as lanes 28..31 of the warp before (no candidates from below reach
    //      them): they only produce column minima. ----
    const int lane = tid & 31, warp = tid >> 5;
    const int e0 = kWarpSpan * warp + kPP * lane;  // staged index of the first own position
    float4 lo[kPP]; float2 hi[kPP];
#pragma unroll
    for (int i = 0; i < kPP; ++i) {
        const int pe = perm_stage(e0 + i);
        lo[i] = s.lo[pe];
        hi[i] = s.hi[pe];
    }
    float fbest[kPP], bbest[kPP]; int fj[kPP], bj[kPP];  // best candidate above / below
#pragma unroll
    for (int i = 0; i < kPP; ++i) { fbest[i] = bbest[i] = INFINITY; fj[i] = bj[i] = -1; }
    search_block<FULL, true>(s, kHaloLanes, e0, radius, lo, hi, fbest, fj, bbest, bj);
#if RTR_PLOC_ROLL
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int dl = kHaloLanes - 1; dl >= 1; --dl)
        search_block<FULL, false>(s, dl, e0, radius, lo, hi, fbest, fj, bbest, bj);
    {   // the 6 pairs inside the thread
        float dloc[kPP][kPP];
#pragma unroll
        for (int a = 0; a < kPP; ++a)
#pragma unroll
            for (int x = a + 1; x < kPP; ++x) {
                float d = pair_half_area(lo[a], hi[a], lo[x], hi[x]);
                if (!FULL) d = (x - a <= radius) ? d : INFINITY;
                dloc[a][x] = d;
            }
#pragma unroll
        for (int a = 0; a < kPP; ++a)
#pragma unroll
            for (int x = kPP - 1; x > a; --x)
                if (dloc[a][x] <= fbest[a]) { fbest[a] = dloc[a][x]; fj[a] = e0 + x; }
#pragma unroll
        for (int x = 1; x < kPP; ++x)
#pragma unroll
            for (int a = 0; a < x; ++a)
                if (dloc[a][x] < bbest[x]) { bbest[x] = dloc[a][x]; bj[x] = e0 + a; }
    }
    if (lane >= kHaloLanes) {
#pragma unroll
        for (int i = 0; i < kPP; ++i) {
            const bool below = bbest[i] <= fbest[i];
            const float best = below ? bbest[i] : fbest[i];
            s.nn[e0 + i - kR] = (best < INFINITY) ? (below ? bj[i] : fj[i]) : -1;  // all inf/NaN: no neighbour (Q4)
        }
    }
    __syncthreads();

    // ---- mutual pairs (plocMerging, bvh.cpp:159-165): lo = lower partner, hi = removed partner ----
    bool is_lo[kPP], is_hi[kPP]; int partner[kPP];
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < kPP; ++i) {
        const int q = e0 + i;  // staged index
        is_lo[i] = is_hi[i] = false; partner[i] = -1;
        const bool in_tile = lane >= kHaloLanes && q >= 2 * kR && q < 2 * kR + kTileT && (base + q) < (int)n;
        if (in_tile) {
            const int j = s.nn[q - kR];
            if (j >= 0 && s.nn[j - kR] == q) { is_lo[i] = q < j; is_hi[i] = q > j; partner[i] = j; }
        }
        packed += (is_lo[i] ? 1u : 0u) + (is_hi[i] ? 1u << 16 : 0u);
    }
    uint32_t cta_total;
    const uint32_t incl = block_scan_incl<kPT>(packed, s.warp, &cta_total);
    uint32_t ex_lo_local = (incl & 0xFFFFu) - (packed & 0xFFFFu);
    uint32_t ex_hi_local = (incl >> 16) - (packed >> 16);

    // ---- decoupled look-back over tiles (plocPrefixScan, bvh.cpp:125-148, as a single pass).  Warp 0 reads
    //      the status words of 32 predecessors per round: the run of published words nearest to this tile,
    //      up to and including the first inclusive one, is summed with shuffles.  (One predecessor per
    //      round made the walk of the first wave of CTAs -- which all publish their aggregates at the same
    //      moment -- the critical path of the whole launch.) ----
    if (warp == 0) {
        const uint32_t agg_lo = cta_total & 0xFFFFu, agg_hi = cta_total >> 16;
        const uint32_t tag = iter + 1u;
        uint32_t ex_lo = 0, ex_hi = 0;
        if (tile > 0) {
            if (lane == 0) st_relaxed_u64(&tile_status[tile], lb_pack(tag, kLbAgg, agg_lo, agg_hi));
            int t = (int)tile - 1;  // nearest predecessor not yet accounted for
            while (true) {
                const int tt = t - lane;
                uint64_t w = lb_pack(tag, kLbIncl, 0u, 0u);  // before tile 0: an inclusive prefix of nothing
                if (tt >= 0) w = ld_relaxed_u64(&tile_status[tt]);
                const uint32_t flag = ((uint32_t)(w >> 56) == (tag & 0xFFu)) ? (uint32_t)((w >> 54) & 3ull) : 0u;
                const uint32_t m_incl = __ballot_sync(0xffffffffu, flag == (uint32_t)kLbIncl);
                const uint32_t m_none = __ballot_sync(0xffffffffu, flag == 0u);
                const int first_incl = m_incl ? __ffs(m_incl) - 1 : 32;
                const int first_none = m_none ? __ffs(m_none) - 1 : 32;
                const bool closed = first_incl < first_none;          // the run ends with an inclusive prefix
                const int take = closed ? first_incl + 1 : first_none;  // lanes [0, take) are consumed
                uint32_t c_lo = lane < take ? (uint32_t)(w & kCntMask) : 0u;
                uint32_t c_hi = lane < take ? (uint32_t)((w >> 27) & kCntMask) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    c_lo += __shfl_xor_sync(0xffffffffu, c_lo, o);
                    c_hi += __shfl_xor_sync(0xffffffffu, c_hi, o);
                }
                ex_lo += c_lo; ex_hi += c_hi;
                if (closed) break;
                t -= take;
            }
        }
        if (lane == 0) {
            st_relaxed_u64(&tile_status[tile], lb_pack(tag, kLbIncl, ex_lo + agg_lo, ex_hi + agg_hi));
            s.ex_lo = ex_lo; s.ex_hi = ex_hi;
            if (tile == tiles - 1) {  // bvh.cpp:105-113
                const uint32_t merges = ex_lo + agg_lo, removed = ex_hi + agg_hi;
                nxt->n_active = n - removed; nxt->total = total + merges; nxt->iter = iter + 1; nxt->tile_counter = 0;
                if (iter < kMaxPlocIterations) {
                    trace_active[iter] = n; trace_merges[iter] = merges; iter_first_id[iter + 1] = total + merges;
                }
            }
        }
    }
    __syncthreads();

    // ---- merge (bvh.cpp:166-188) + compaction (bvh.cpp:150-156) ----
    const uint32_t lo_base = total + s.ex_lo, hi_base = s.ex_hi;
#pragma unroll
    for (int i = 0; i < kPP; ++i) {
        const int q = e0 + i;
        const bool in_tile = lane >= kHaloLanes && q >= 2 * kR && q < 2 * kR + kTileT && (base + q) < (int)n;
        if (in_tile) {
            uint32_t out_id = s.id[perm_stage(q)];
            if (is_lo[i]) {
                const int pj = perm_stage(partner[i]);
                const float4 plo = s.lo[pj]; const float2 phi = s.hi[pj];
                const uint32_t new_id = lo_base + ex_lo_local;
                const uint32_t cl = out_id, cr = s.id[pj];
                store_box(node, new_id,
                          make_float4(fminf(lo[i].x, plo.x), fminf(lo[i].y, plo.y), fminf(lo[i].z, plo.z), fmaxf(lo[i].w, plo.w)),
                          make_float4(fmaxf(hi[i].x, phi.x), fmaxf(hi[i].y, phi.y), __uint_as_float(cl), __uint_as_float(cr)));
                const uint32_t sl = cl < n_leaves ? 1u : __ldcg(isize + (cl - n_leaves)).x;  // written by other SMs in earlier iterations
                const uint32_t sr = cr < n_leaves ? 1u : __ldcg(isize + (cr - n_leaves)).x;
                isize[new_id - n_leaves] = make_uint2(sl + sr + 1u, sl);  // nodes in the subtree, and in its left part
                out_id = new_id;
                ++ex_lo_local;
            }
            if (!is_hi[i]) cout[(uint32_t)(base + q) - (hi_base + ex_hi_local)] = out_id;
            else ++ex_hi_local;
        }
    }
}


// grid-wide barrier of the persistent kernels below (cooperative launch: every CTA is resident).  `counter` only
// grows: barrier k is passed once it reaches gridDim.x * k.  The fences order this CTA's writes before its arrival
// and invalidate the SM's L1 after the wait (data written by other SMs before the barrier is read after it).
__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t& generation) {
    __syncthreads();
    ++generation;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const uint32_t target = gridDim.x * generation;
        uint32_t pause = 32u;  // hundreds of CTAs poll one word: back off so that the late arrivals' atomics get through
        while (ld_acquire_u32(counter) < target) { __nanosleep(pause); pause = pause < 256u ? pause * 2u : 256u; }
        __threadfence();
    }
    __syncthreads();
}

// The whole PLOC loop down to kTailN clusters as ONE persistent kernel (cooperative launch, one CTA per resident
// slot): per iteration the CTAs draw tiles from a counter, then meet at a grid barrier -- no launch per iteration, no
// host round trip to size a grid or to learn the iteration count (the reference's `while (_Iteration > 1)`,
// bvh.cpp:62, runs on the device).  ctl[0] = barrier counter, ctl[1] = status (see BuildStatus).
template <bool FULL>
__global__ void __launch_bounds__(kPT, RTR_PLOC_MINB)
ploc_loop_kernel(uint32_t n_leaves, int radius, uint32_t* __restrict__ buf0, uint32_t* __restrict__ buf1,
                 float4* __restrict__ node, uint2* __restrict__ isize,
                 PlocState* __restrict__ state, uint64_t* __restrict__ tile_status,
                 uint32_t* __restrict__ trace_active, uint32_t* __restrict__ trace_merges,
                 uint32_t* __restrict__ iter_first_id, uint32_t* __restrict__ barrier_counter,
                 unsigned long long* __restrict__ iter_ns) {
    __shared__ IterSmem s;
    const int tid = (int)threadIdx.x;
    uint32_t generation = 0u, parity = 0u, prev_n = 0xFFFFFFFFu;
    auto stamp = [&](uint32_t slot) {  // %globaltimer at the start of iteration `slot` (rtr_bvh_iteration_times)
        if (blockIdx.x == 0 && tid == 0 && slot <= kMaxPlocIterations) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            iter_ns[slot] = t;
        }
    };
    while (true) {
        PlocState* cur = &state[parity];
        PlocState* nxt = &state[parity ^ 1u];
        const uint32_t n = ld_relaxed_u32(&cur->n_active), total = ld_relaxed_u32(&cur->total), iter = ld_relaxed_u32(&cur->iter);
        // done, handed to the tail kernel, or stuck (non-finite areas never merge, Q4: the reference would spin forever)
        stamp(iter);
        if (n <= kTailN || n >= prev_n || iter >= kMaxPlocIterations) break;
        prev_n = n;
        const uint32_t tiles = (n + kTileT - 1) / kTileT;
        const uint32_t* __restrict__ cin = (iter & 1u) ? buf1 : buf0;
        uint32_t* __restrict__ cout = (iter & 1u) ? buf0 : buf1;
        if (tiles <= gridDim.x) {
            // one round: tile = blockIdx.x.  Every CTA of the (cooperative) grid is resident, so the tiles a look-back
            // waits for are held by running CTAs, and no ticket has to be drawn (one L2 round trip before the first load:
            // measured 1.5 us of every small iteration)
            if (blockIdx.x < tiles) {
                ploc_process_tile<FULL>(s, blockIdx.x, tiles, n, total, iter, n_leaves, radius, cin, cout, node, isize, nxt, tile_status,
                                        trace_active, trace_merges, iter_first_id);
                __syncthreads();
            }
        } else {
            // (Tried: a second staging buffer filled by cp.async -- ids during the first block of the search, boxes during
            //  the rest -- so that the two dependent gathers of the next tile hide behind the current search.  55 KB of
            //  shared memory per CTA, three small LDGSTS per slot instead of one 256-bit load: 473 us instead of 344 us
            //  for the first iteration, the loop 3.28 ms instead of 2.55 ms.  Reverted.)
            while (true) {
                // tiles are handed out in start order (dynamic: a few per cent faster than static rounds on the large
                // iterations), so the look-back only ever waits on tiles of running CTAs
                if (tid == 0) s.tile = atomicAdd(&cur->tile_counter, 1u);
                __syncthreads();
                const uint32_t tile = s.tile;
                if (tile >= tiles) break;
                ploc_process_tile<FULL>(s, tile, tiles, n, total, iter, n_leaves, radius, cin, cout, node, isize, nxt, tile_status,
                                        trace_active, trace_merges, iter_first_id);
                __syncthreads();  // the staging area is reused by the next tile
            }
        }
        grid_barrier(barrier_counter, generation);
        parity ^= 1u;
    }
    if (blockIdx.x == 0 && tid == 0) {  // what the tail kernel and the flatten start from
        const PlocState* cur = &state[parity];
        state[2].n_active = ld_relaxed_u32(&cur->n_active); state[2].total = ld_relaxed_u32(&cur->total);
        state[2].iter = ld_relaxed_u32(&cur->iter); state[2].tile_counter = 0u;
    }
}

// ---------------------------------------------------------------------------------------
// tail: n_active <= 1024, one CTA, everything in shared memory until one cluster is left
// ---------------------------------------------------------------------------------------
struct TailSmem {
    float box[2][6][kTailN];
    uint32_t id[2][kTailN];
    uint32_t size[2][kTailN];
    int nn[kTailN];
    uint32_t warp[kTailN / 32 + 1];
};

__global__ void __launch_bounds__(kTailN)
ploc_tail_kernel(uint32_t n_leaves, int radius,
                 const uint32_t* __restrict__ buf0, const uint32_t* __restrict__ buf1,
                 float4* __restrict__ node, uint2* __restrict__ isize,
                 PlocState* __restrict__ state, uint32_t* __restrict__ trace_active,
                 uint32_t* __restrict__ trace_merges, uint32_t* __restrict__ iter_first_id,
                 unsigned long long* __restrict__ iter_ns) {
    extern __shared__ __align__(16) unsigned char tail_raw[];
    TailSmem& s = *reinterpret_cast<TailSmem*>(tail_raw);
    const uint32_t tid = threadIdx.x;
    PlocState* fin = &state[2];
    uint32_t n = fin->n_active, total = fin->total, iter = fin->iter;
    if (n > kTailN) return;  // the loop kernel gave up (nothing merges any more, Q4): rtr_bvh_finish reports it
    const uint32_t* __restrict__ cin = (iter & 1u) ? buf1 : buf0;
    int cur_buf = 0;
    if (tid < n) {
        const uint32_t id = cin[tid];
        const Box bx_ = load_box(node, id); const float4 lo = bx_.lo, hi = bx_.hi;
        s.id[0][tid] = id;
        s.size[0][tid] = id < n_leaves ? 1u : isize[id - n_leaves].x;
        s.box[0][0][tid] = lo.x; s.box[0][1][tid] = lo.y; s.box[0][2][tid] = lo.z;
        s.box[0][3][tid] = lo.w; s.box[0][4][tid] = hi.x; s.box[0][5][tid] = hi.y;
    }
    __syncthreads();
    while (n > 1) {
        float (*bx)[kTailN] = s.box[cur_buf];
        const BoxSoA B{bx[0], bx[1], bx[2], bx[3], bx[4], bx[5]};
        const int q = (int)tid;
        int nn = -1;
        if (tid < n) nn = nearest_neighbour(B, q, max(0, q - radius), min(q + radius + 1, (int)n));
        s.nn[tid] = nn;
        __syncthreads();
        bool is_lo = false, is_hi = false;
        if (nn >= 0 && s.nn[nn] == q) { is_lo = q < nn; is_hi = q > nn; }
        uint32_t cta_total;
        const uint32_t packed = (is_lo ? 1u : 0u) | (is_hi ? 1u << 16 : 0u);
        const uint32_t incl = block_scan_incl<kTailN>(packed, s.warp, &cta_total);
        const uint32_t merges = cta_total & 0xFFFFu, removed = cta_total >> 16;
        const int nb = cur_buf ^ 1;
        if (tid < n && !is_hi) {
            const uint32_t dst = tid - ((incl >> 16) - (is_hi ? 1u : 0u));
            if (is_lo) {
                const uint32_t new_id = total + (incl & 0xFFFFu) - 1u;
                const uint32_t cl = s.id[cur_buf][q], cr = s.id[cur_buf][nn];
                const float mnx = fminf(B.minx[q], B.minx[nn]), mny = fminf(B.miny[q], B.miny[nn]);
                const float mnz = fminf(B.minz[q], B.minz[nn]), mxx = fmaxf(B.maxx[q], B.maxx[nn]);
                const float mxy = fmaxf(B.maxy[q], B.maxy[nn]), mxz = fmaxf(B.maxz[q], B.maxz[nn]);
                const uint32_t sz = s.size[cur_buf][q] + s.size[cur_buf][nn] + 1u;
                store_box(node, new_id, make_float4(mnx, mny, mnz, mxx),
                          make_float4(mxy, mxz, __uint_as_float(cl), __uint_as_float(cr)));
                isize[new_id - n_leaves] = make_uint2(sz, s.size[cur_buf][q]);
                s.id[nb][dst] = new_id; s.size[nb][dst] = sz;
                s.box[nb][0][dst] = mnx; s.box[nb][1][dst] = mny; s.box[nb][2][dst] = mnz;
                s.box[nb][3][dst] = mxx; s.box[nb][4][dst] = mxy; s.box[nb][5][dst] = mxz;
            } else {
                s.id[nb][dst] = s.id[cur_buf][q]; s.size[nb][dst] = s.size[cur_buf][q];
#pragma unroll
                for (int k = 0; k < 6; ++k) s.box[nb][k][dst] = bx[k][q];
            }
        }
        if (tid == 0 && iter < kMaxPlocIterations) {
            trace_active[iter] = n; trace_merges[iter] = merges; iter_first_id[iter + 1] = total + merges;
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            iter_ns[iter + 1] = t;  // end of iteration `iter` (its start: the previous stamp)
        }
        n -= removed; total += merges; iter += 1; cur_buf = nb;
        __syncthreads();
        if (merges == 0) break;  // non-finite areas (Q4): the reference would spin forever
    }
    if (tid == 0) { fin->n_active = n; fin->total = total; fin->iter = iter; fin->tile_counter = 0; }
}

// ---------------------------------------------------------------------------------------
// flatten: scene.cpp:189-208 as a top-down pass over creation levels (iterations, last first)
// ---------------------------------------------------------------------------------------
// `slot` goes to the third padding word (bytes 44-47, unspecified in the reference layout): leaves keep
// their cluster id = wtri slot there for the traversal kernels
__device__ __forceinline__ void store_node(rtr_node* __restrict__ flat, uint32_t pos, const float4 lo, const float4 hi,
                                           uint32_t tri, uint32_t left, uint32_t right, uint32_t slot = 0u) {
    uint4* dst = reinterpret_cast<uint4*>(flat + pos);
    dst[0] = make_uint4(__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), 0u);
    dst[1] = make_uint4(__float_as_uint(lo.w), __float_as_uint(hi.x), __float_as_uint(hi.y), 0u);
    dst[2] = make_uint4(tri, left, right, slot);
}

// Pass 1 (top-down over creation levels, last iteration first): DFS pre-order position of every cluster.
// One inner cluster c at position p places its children -- left at p + 1, right behind the left subtree
// (scene.cpp:189-199) -- and records which cluster sits at each position.  Only 4-byte arrays are touched
// (order 8 B/triangle, ipos and isize 4 B/triangle), which the 126 MB L2 absorbs.
__device__ __forceinline__ void place_children(uint32_t c, uint32_t n_leaves, const float4* __restrict__ node,
                                               const uint2* __restrict__ isize, uint32_t* __restrict__ ipos,
                                               uint32_t* __restrict__ order) {
    const uint32_t p = __ldcg(ipos + (c - n_leaves));  // written by another SM one level up (same kernel): L2, not L1
    const float4 hi = __ldg(node + 2 * (size_t)c + 1);
    const uint32_t L = __float_as_uint(hi.z), R = __float_as_uint(hi.w);
    const uint32_t size_l = isize[c - n_leaves].y;  // own record (coalesced over a level), not the child's
    const uint32_t pos_l = p + 1u, pos_r = p + 1u + size_l;
    order[pos_l] = L;
    order[pos_r] = R;
    if (L >= n_leaves) ipos[L - n_leaves] = pos_l;
    if (R >= n_leaves) ipos[R - n_leaves] = pos_r;
}

// All creation levels in ONE persistent kernel (cooperative launch), last iteration first: the level table
// (iter_first_id) and the iteration count stay on the device, the host launches this without knowing either.  A run
// of small levels (the top of the tree: the tail's iterations) is walked by CTA 0 alone with __syncthreads between
// levels; a large level is spread over the grid and closed by a grid barrier.
constexpr uint32_t kSmallLevel = 2048;
constexpr int kFlattenBlock = 512;
__global__ void __launch_bounds__(kFlattenBlock)
flatten_positions_kernel(const PlocState* __restrict__ state, const uint32_t* __restrict__ iter_first_id, uint32_t n_leaves,
                         const float4* __restrict__ node, const uint2* __restrict__ isize,
                         uint32_t* __restrict__ ipos, uint32_t* __restrict__ order, uint32_t* __restrict__ barrier_counter) {
    const PlocState fin = state[2];
    if (fin.n_active != 1u || fin.total != 2u * n_leaves - 1u) return;  // the build did not converge: nothing to place
    uint32_t generation = 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // root 2n-2 -> position 0 (scene.cpp:205); its level is CTA 0's
        order[0] = 2u * n_leaves - 2u;
        ipos[n_leaves - 2u] = 0u;
    }
    __syncthreads();
    int it = (int)fin.iter - 1;
    while (it >= 0) {
        const uint32_t first = iter_first_id[it], count = iter_first_id[it + 1] - first;
        if (count > kSmallLevel) {
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
                place_children(first + i, n_leaves, node, isize, ipos, order);
            --it;
            grid_barrier(barrier_counter, generation);
        } else {
            int lo = it;
            while (lo - 1 >= 0 && iter_first_id[lo] - iter_first_id[lo - 1] <= kSmallLevel) --lo;
            if (blockIdx.x == 0) {
                for (int l = it; l >= lo; --l) {
                    const uint32_t f = iter_first_id[l], c = iter_first_id[l + 1] - f;
                    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) place_children(f + i, n_leaves, node, isize, ipos, order);
                    __syncthreads();
                }
            }
            it = lo - 1;
            if (it >= 0) grid_barrier(barrier_counter, generation);
        }
    }
}

// leaf traversal record (bvh.cuh): exact box | P0 P1 P2 | triangle id.  wt = the triangle's wtri slot.
__device__ __forceinline__ void store_leaf_record(uint4* __restrict__ pairs, uint32_t p, const float4 lo, const float4 hi,
                                                  const float4* __restrict__ wt, uint32_t tri) {
    const float4 a = __ldg(wt), b = __ldg(wt + 1), c = __ldg(wt + 2);  // (P0.xyz,P1.x)(P1.yz,P2.xy)(P2.z,-,-,-)
    float* dst = reinterpret_cast<float*>(pairs + (size_t)p * 4);
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(dst), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w), "f"(hi.x), "f"(hi.y), "f"(a.x), "f"(a.y) : "memory");
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(dst + 8), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "f"(c.x),
                   "f"(__uint_as_float(tri)) : "memory");
}
__device__ __forceinline__ void store_inner_record(uint4* __restrict__ pairs, uint32_t p, const uint4 o0, const uint4 o1) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(pairs + (size_t)p * 4), "r"(o0.x), "r"(o0.y), "r"(o0.z), "r"(o0.w), "r"(o1.x), "r"(o1.y), "r"(o1.z), "r"(o1.w)
                 : "memory");
}

// Pass 2, one thread per flat position p: gathers the 32-byte record of the cluster placed there and
// streams out the reference's 48-byte node (scene.cpp:189-208) and the traversal record of the position
// (bvh.cuh: the compressed child pair of an inner node, box + world-space vertices of a leaf).  All stores of
// a warp are contiguous; the left child's record is the next lane's own (position p + 1).
__global__ void __launch_bounds__(256)
flatten_emit_kernel(const PlocState* __restrict__ state, uint32_t nb_nodes, uint32_t n_leaves, const uint32_t* __restrict__ order,
                    const float4* __restrict__ node, const uint2* __restrict__ isize,
                    const float4* __restrict__ wtri, rtr_node* __restrict__ flat, uint4* __restrict__ pairs) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (state[2].n_active != 1u || state[2].total != nb_nodes) {
        // The build did not converge (non-finite triangle data, Q4; rtr_bvh_finish reports it).  Leave a tree that is
        // safe to trace: one leaf with an empty box, which no ray enters.
        if (p == 0) {
            const float4 lo = make_float4(INFINITY, INFINITY, INFINITY, -INFINITY), hi = make_float4(-INFINITY, -INFINITY, 0.f, 0.f);
            store_node(flat, 0, lo, hi, 0u, 0u, 0u, 0u);
            store_leaf_record(pairs, 0, lo, hi, wtri, 0u);
        }
        return;
    }
    const bool live = p < nb_nodes;
    const uint32_t c = live ? order[p] : 0u;
    const Box me = load_box(node, c);
    const uint32_t my_size_l = (live && c >= n_leaves) ? isize[c - n_leaves].y : 1u;  // nodes in the left subtree
    // the records of the next two positions -- the left child of an inner node and that child's own left child --
    // come through the warp instead of two more gathers
    Box nx, nx2;
    nx.lo.x = __shfl_down_sync(0xffffffffu, me.lo.x, 1); nx.lo.y = __shfl_down_sync(0xffffffffu, me.lo.y, 1);
    nx.lo.z = __shfl_down_sync(0xffffffffu, me.lo.z, 1); nx.lo.w = __shfl_down_sync(0xffffffffu, me.lo.w, 1);
    nx.hi.x = __shfl_down_sync(0xffffffffu, me.hi.x, 1); nx.hi.y = __shfl_down_sync(0xffffffffu, me.hi.y, 1);
    nx.hi.z = __shfl_down_sync(0xffffffffu, me.hi.z, 1); nx.hi.w = __shfl_down_sync(0xffffffffu, me.hi.w, 1);
    nx2.lo.x = __shfl_down_sync(0xffffffffu, me.lo.x, 2); nx2.lo.y = __shfl_down_sync(0xffffffffu, me.lo.y, 2);
    nx2.lo.z = __shfl_down_sync(0xffffffffu, me.lo.z, 2); nx2.lo.w = __shfl_down_sync(0xffffffffu, me.lo.w, 2);
    nx2.hi.x = __shfl_down_sync(0xffffffffu, me.hi.x, 2); nx2.hi.y = __shfl_down_sync(0xffffffffu, me.hi.y, 2);
    nx2.hi.z = __shfl_down_sync(0xffffffffu, me.hi.z, 2); nx2.hi.w = __shfl_down_sync(0xffffffffu, me.hi.w, 2);
    const uint32_t nx_size_l = __shfl_down_sync(0xffffffffu, my_size_l, 1);
    if (!live) return;
    const bool leaf = c < n_leaves;
    const uint32_t L = __float_as_uint(me.hi.z), R = __float_as_uint(me.hi.w);
    if (leaf) {  // leaves keep 0/0 and their triangle id; the wtri slot (= cluster id) rides in the padding word
        store_node(flat, p, me.lo, me.hi, L, 0u, 0u, c);
        store_leaf_record(pairs, p, me.lo, me.hi, wtri + 3 * (size_t)c, L);
        return;
    }
    const uint32_t lane = lane_id();
    const bool lleaf = L < n_leaves, rleaf = R < n_leaves;
    const uint32_t pos_l = p + 1u, pos_r = p + 1u + my_size_l;
    const Box br = load_box(node, R);
    const Box bl = (lane < 31u) ? nx : load_box(node, L);
    store_node(flat, p, me.lo, me.hi, 0u, pos_l, pos_r);  // internal _TriangleId stays 0 (bvh.cpp:415-420)
    uint4 o0, o1;
    trav_encode_inner(me.lo, make_float2(me.hi.x, me.hi.y), bl.lo, make_float2(bl.hi.x, bl.hi.y),
                      br.lo, make_float2(br.hi.x, br.hi.y), lleaf, rleaf, pos_r, o0, o1);
    // ---- second half (bvh.cuh): the four grandchild slots on the node's own grid.  Slot 0 / 1 = the children of L
    //      (L itself, twice, when it is a leaf), slot 2 / 3 = those of R; slot 0 sits at p + 2, slot 2 at pos_r + 1. ----
    float4 glo[4]; float2 ghi[4];
    uint32_t idx1, idx3, leaf_bits = 0u;
    if (lleaf) {
        glo[0] = glo[1] = bl.lo; ghi[0] = ghi[1] = make_float2(bl.hi.x, bl.hi.y);
        idx1 = pos_l; leaf_bits |= 3u;
    } else {
        const uint32_t LL = __float_as_uint(bl.hi.z), LR = __float_as_uint(bl.hi.w);
        const Box b0 = (lane < 30u) ? nx2 : load_box(node, LL);
        const Box b1 = load_box(node, LR);
        const uint32_t size_ll = (lane < 31u) ? nx_size_l : isize[L - n_leaves].y;
        glo[0] = b0.lo; ghi[0] = make_float2(b0.hi.x, b0.hi.y);
        glo[1] = b1.lo; ghi[1] = make_float2(b1.hi.x, b1.hi.y);
        idx1 = p + 2u + size_ll;
        leaf_bits |= (LL < n_leaves ? 1u : 0u) | (LR < n_leaves ? 2u : 0u);
    }
    if (rleaf) {
        glo[2] = glo[3] = br.lo; ghi[2] = ghi[3] = make_float2(br.hi.x, br.hi.y);
        idx3 = pos_r; leaf_bits |= 0xCu;
    } else {
        const uint32_t RL = __float_as_uint(br.hi.z), RR = __float_as_uint(br.hi.w);
        const Box b2 = load_box(node, RL), b3 = load_box(node, RR);
        glo[2] = b2.lo; ghi[2] = make_float2(b2.hi.x, b2.hi.y);
        glo[3] = b3.lo; ghi[3] = make_float2(b3.hi.x, b3.hi.y);
        idx3 = pos_r + 1u + isize[R - n_leaves].y;
        leaf_bits |= (RL < n_leaves ? 4u : 0u) | (RR < n_leaves ? 8u : 0u);
    }
    bool ok = true;
    uint4 o2, o3;
    trav_encode_quads(me.lo, make_float2(me.hi.x, me.hi.y), glo, ghi, idx1 | ((leaf_bits & 2u) ? 0x80000000u : 0u),
                      idx3 | ((leaf_bits & 8u) ? 0x80000000u : 0u), o2, o3, ok);
    if (ok && !((o0.w >> 24) & 4u)) o0.w |= (0x80u | (leaf_bits << 3)) << 24;  // second half usable + which slots are leaves
    store_inner_record(pairs, p, o0, o1);
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(pairs + (size_t)p * 4 + 2), "r"(o2.x), "r"(o2.y), "r"(o2.z), "r"(o2.w), "r"(o3.x), "r"(o3.y), "r"(o3.z), "r"(o3.w) : "memory");
}

__global__ void flatten_single_leaf_kernel(const float4* node, const float4* wtri, rtr_node* flat, uint4* pairs) {
    const Box bx_ = load_box(node, 0); const float4 lo = bx_.lo, hi = bx_.hi;
    store_node(flat, 0, lo, hi, __float_as_uint(hi.z), 0u, 0u, 0u);
    store_leaf_record(pairs, 0, lo, hi, wtri, __float_as_uint(hi.z));
}

// traversal records (bvh.cuh) of an adopted / received flat array
__global__ void __launch_bounds__(256)
pack_pairs_kernel(const rtr_node* __restrict__ flat, uint32_t nb_nodes, uint32_t by_rank, const float4* __restrict__ wtri,
                  uint4* __restrict__ pairs, TraceParams* __restrict__ tparams) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb_nodes) return;
    const uint32_t nb_tris = (nb_nodes + 1u) / 2u;
    const uint4* me = reinterpret_cast<const uint4*>(flat + i);
    const uint4 m0 = __ldg(me), m1 = __ldg(me + 1), links = __ldg(me + 2);
    const float4 nlo = make_float4(__uint_as_float(m0.x), __uint_as_float(m0.y), __uint_as_float(m0.z), __uint_as_float(m1.x));
    const float4 nhi = make_float4(__uint_as_float(m1.y), __uint_as_float(m1.z), 0.f, 0.f);
    if (links.y == 0u && links.z == 0u) {
        const uint32_t slot = by_rank ? links.w : links.x;
        if (slot >= nb_tris || links.x >= nb_tris) {  // a foreign array naming a triangle that does not exist: never followed
            atomicAdd(&tparams->bad_layout, 1u);
            store_leaf_record(pairs, i, make_float4(INFINITY, INFINITY, INFINITY, -INFINITY), make_float4(-INFINITY, -INFINITY, 0.f, 0.f), wtri, 0u);
            return;
        }
        store_leaf_record(pairs, i, nlo, nhi, wtri + 3 * (size_t)slot, links.x);
        return;
    }
    // not the DFS pre-order of scene.cpp:189-208, or links that leave the array: counted (the call fails with
    // RTR_E_UNSUPPORTED) and never followed
    if (links.y != i + 1u || links.y >= nb_nodes || links.z >= nb_nodes || links.z <= i) {
        atomicAdd(&tparams->bad_layout, 1u);
        store_inner_record(pairs, i, make_uint4(0u, 0u, 0u, 4u << 24), make_uint4(0u, 0u, 0u, i));
        return;
    }
    const uint4* l = reinterpret_cast<const uint4*>(flat + links.y);
    const uint4* r = reinterpret_cast<const uint4*>(flat + links.z);
    const uint4 l0 = __ldg(l), l1 = __ldg(l + 1), l2 = __ldg(l + 2);
    const uint4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
    const bool lleaf = l2.y == 0u && l2.z == 0u, rleaf = r2.y == 0u && r2.z == 0u;
    uint4 o0, o1;
    trav_encode_inner(nlo, make_float2(nhi.x, nhi.y),
                      make_float4(__uint_as_float(l0.x), __uint_as_float(l0.y), __uint_as_float(l0.z), __uint_as_float(l1.x)),
                      make_float2(__uint_as_float(l1.y), __uint_as_float(l1.z)),
                      make_float4(__uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z), __uint_as_float(r1.x)),
                      make_float2(__uint_as_float(r1.y), __uint_as_float(r1.z)), lleaf, rleaf, links.z, o0, o1);
    store_inner_record(pairs, i, o0, o1);
}

// Second half of every inner traversal record (bvh.cuh) of an ADOPTED / received flat array (a BVH built here gets both
// halves from flatten_emit_kernel): the four grandchild slots on the node's own grid, for the traversal's wide step.
__global__ void __launch_bounds__(256)
pack_quads_kernel(const rtr_node* __restrict__ flat, uint32_t nb_nodes, uint4* __restrict__ pairs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb_nodes) return;
    const uint4* me = reinterpret_cast<const uint4*>(flat + i);
    const uint4 links = __ldg(me + 2);
    if (links.y == 0u && links.z == 0u) return;
    if (links.y >= nb_nodes || links.z >= nb_nodes) return;  // flagged by pack_pairs_kernel
    const uint4 m0 = __ldg(me), m1 = __ldg(me + 1);
    const uint32_t child[2] = {links.y, links.z};
    float4 lo[4];
    float2 hi[4];
    uint32_t idx[4], leaf_bits = 0u;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint4* c = reinterpret_cast<const uint4*>(flat + child[k]);
        const uint4 c2 = __ldg(c + 2);
        if (c2.y == 0u && c2.z == 0u) {  // the child is a leaf: it is its own first slot, the second repeats it (unused)
            const uint4 c0 = __ldg(c), c1 = __ldg(c + 1);
            lo[2 * k] = lo[2 * k + 1] = make_float4(__uint_as_float(c0.x), __uint_as_float(c0.y), __uint_as_float(c0.z), __uint_as_float(c1.x));
            hi[2 * k] = hi[2 * k + 1] = make_float2(__uint_as_float(c1.y), __uint_as_float(c1.z));
            idx[2 * k] = idx[2 * k + 1] = child[k];
            leaf_bits |= 3u << (2 * k);
        } else {
            const uint32_t g[2] = {c2.y, c2.z};
            if (g[0] >= nb_nodes || g[1] >= nb_nodes) return;  // flagged by pack_pairs_kernel at the child
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint4* q = reinterpret_cast<const uint4*>(flat + g[j]);
                const uint4 g0 = __ldg(q), g1 = __ldg(q + 1), g2 = __ldg(q + 2);
                lo[2 * k + j] = make_float4(__uint_as_float(g0.x), __uint_as_float(g0.y), __uint_as_float(g0.z), __uint_as_float(g1.x));
                hi[2 * k + j] = make_float2(__uint_as_float(g1.y), __uint_as_float(g1.z));
                idx[2 * k + j] = g[j];
                if (g2.y == 0u && g2.z == 0u) leaf_bits |= 1u << (2 * k + j);
            }
        }
    }
    bool ok = true;
    uint4 o2, o3;
    trav_encode_quads(make_float4(__uint_as_float(m0.x), __uint_as_float(m0.y), __uint_as_float(m0.z), __uint_as_float(m1.x)),
                      make_float2(__uint_as_float(m1.y), __uint_as_float(m1.z)), lo, hi,
                      idx[1] | ((leaf_bits & 2u) ? 0x80000000u : 0u), idx[3] | ((leaf_bits & 8u) ? 0x80000000u : 0u), o2, o3, ok);
    uint4* rec = pairs + (size_t)i * 4;
    uint32_t w3 = rec[0].w;
    if (ok && !((w3 >> 24) & 4u)) w3 |= (0x80u | (leaf_bits << 3)) << 24;
    rec[0].w = w3;
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(rec + 2), "r"(o2.x), "r"(o2.y), "r"(o2.z), "r"(o2.w), "r"(o3.x), "r"(o3.y), "r"(o3.z), "r"(o3.w) : "memory");
}

// ---------------------------------------------------------------------------------------
// BVH_Params view by cluster id for the accessors / the cr::BVH shim (not on the timed path)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
export_clusters_kernel(uint32_t n_leaves, const float4* __restrict__ node,
                       rtr_node* __restrict__ clusters, uint32_t* __restrict__ parent, uint32_t* __restrict__ left,
                       uint32_t* __restrict__ right, uint8_t* __restrict__ is_leaf) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * n_leaves - 1) return;
    const Box bx_ = load_box(node, c); const float4 lo = bx_.lo, hi = bx_.hi;
    const bool leaf = c < n_leaves;
    const uint32_t L = __float_as_uint(hi.z), R = __float_as_uint(hi.w);
    if (clusters) store_node(clusters, c, lo, hi, leaf ? L : 0u, 0u, 0u);  // links stay 0 (Q10)
    if (left) left[c] = leaf ? RTR_NONE : L;
    if (right) right[c] = leaf ? RTR_NONE : R;
    if (is_leaf) is_leaf[c] = leaf ? 1 : 0;
    if (parent && !leaf) { parent[L] = c; parent[R] = c; }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------
int rtr_bvh_compute_trace_params(rtr_bvh* b) {
    rtr_ctx* ctx = b->ctx;
    RTR_CUDA(ctx, cudaMemsetAsync(b->tparams, 0, sizeof(TraceParams), ctx->stream));
    if (b->wtri_own_cap < b->n) {
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (b->wtri_own) cudaFree(b->wtri_own);
        b->wtri_own = nullptr; b->wtri_own_cap = 0;
        RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->wtri_own), (size_t)b->n * 3 * sizeof(float4)));
        b->wtri_own_cap = b->n;
    }
    if (b->n) {
        edge_bound_kernel<<<(b->n + 255) / 256, 256, 0, ctx->stream>>>(b->tris, b->meshes, b->n, b->wtri_own, b->tparams);
        RTR_LAUNCH_CHECK(ctx);
    }
    b->wtri_view = b->wtri_own;
    b->wtri_by_rank = false;
    RTR_CHECK(rtr_bvh_pack_pairs_own(b));
    // the traversal records rely on the left child sitting at index + 1 (scene.cpp:189-199 always produces that)
    TraceParams* h = static_cast<TraceParams*>(ctx->pinned);
    RTR_CUDA(ctx, cudaMemcpyAsync(h, b->tparams, sizeof(TraceParams), cudaMemcpyDeviceToHost, ctx->stream));
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h->bad_layout)
        return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "adopted node array is not in DFS pre-order: %u inner nodes have Left != index + 1",
                             h->bad_layout);
    return RTR_OK;
}

int rtr_bvh_pack_pairs_own(rtr_bvh* b) {
    rtr_ctx* ctx = b->ctx;
    if (b->pairs_own_cap < b->n) {
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (b->pairs_own) cudaFree(b->pairs_own);
        b->pairs_own = nullptr; b->pairs_own_cap = 0;
        RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->pairs_own), (2 * (size_t)b->n - 1) * 4 * sizeof(uint4)));
        b->pairs_own_cap = b->n;
    }
    const uint32_t nc = 2 * b->n - 1;
    pack_pairs_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(b->flat_view, nc, b->wtri_by_rank ? 1u : 0u, b->wtri_view,
                                                                 b->pairs_own, b->tparams);
    RTR_LAUNCH_CHECK(ctx);
    pack_quads_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(b->flat_view, nc, b->pairs_own);
    RTR_LAUNCH_CHECK(ctx);
    b->pairs_view = b->pairs_own;
    return RTR_OK;
}

static int record(rtr_bvh* b, int i) {
    if (b->timing) RTR_CUDA(b->ctx, cudaEventRecord(b->ev[i], b->ctx->stream));
    return RTR_OK;
}

int rtr_bvh_run_build(rtr_bvh* b) {
    rtr_ctx* ctx = b->ctx;
    const uint32_t n = b->n;
    const int radius = (int)b->radius;
    b->built = false;
    b->iterations = 0;
    RTR_CHECK(record(b, 0));

    // 1. scene box + Morton codes (+ iota indices)
    if (b->key_bits == 64 && b->codes64_cap < n) {
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (b->codes64) cudaFree(b->codes64);
        b->codes64 = nullptr; b->codes64_cap = 0;
        RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->codes64), (size_t)b->capacity * sizeof(uint64_t)));
        b->codes64_cap = b->capacity;
    }
    RTR_CHECK(rtr_morton_launch(ctx, b->tris, n, b->array_len, b->meshes, b->codes, b->tri_idx,
                                b->key_bits == 64 ? b->codes64 : nullptr, b->bounds12, b->ordered6));
    RTR_CHECK(record(b, 1));
    // 2. stable sort of (code, index): 30-bit codes in a u32 (the reference), or 63-bit codes in a u64 (21 bits
    //    per axis, 8 passes; the codes array then holds the unsorted 30-bit codes and is not exported)
    if (b->key_bits == 64) RTR_CHECK(rtr_sort_impl_u64(ctx, b->codes64, b->tri_idx, n, 0, 63));
    else RTR_CHECK(rtr_sort_impl_u32(ctx, b->codes, b->tri_idx, n, 0, 32));
    RTR_CHECK(record(b, 2));
    // 3. leaves
    RTR_CUDA(ctx, cudaMemsetAsync(b->tparams, 0, sizeof(TraceParams), ctx->stream));
    RTR_PROF(ctx, "leaf_init_kernel");
    leaf_init_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(b->tris, b->meshes, b->tri_idx, n, b->node,
                                                               b->cin, b->wtri, b->tparams);
    RTR_LAUNCH_CHECK(ctx);
    ploc_state_init_kernel<<<1, 1, 0, ctx->stream>>>(b->state, n, b->iter_first_id, b->ctl);
    RTR_LAUNCH_CHECK(ctx);
    RTR_CHECK(record(b, 3));

    // 4. PLOC loop: one persistent kernel down to kTailN clusters, one CTA for the rest.  Nothing here waits for
    //    the device: iteration count, level table and the verdict stay there (rtr_bvh_finish reads them on demand).
    RTR_CUDA(ctx, cudaFuncSetAttribute(ploc_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem)));
    if (n > kTailN) {
        const bool full = radius == kR;
        const void* kern = full ? (const void*)ploc_loop_kernel<true> : (const void*)ploc_loop_kernel<false>;
        int& per_sm = full ? ctx->ploc_ctas_per_sm[0] : ctx->ploc_ctas_per_sm[1];
        if (per_sm == 0) {
            int c = 0;
            RTR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kern, kPT, 0));
            per_sm = c < 1 ? 1 : c;
        }
        const uint32_t tiles = (n + kTileT - 1) / kTileT;
        // Every CTA of a cooperative grid must be resident at once.  SMs the caller keeps for the kernels of a concurrent
        // collective (rtr_ctx_reserve_sms) are left out of the count: sized for the whole device, the grid could not
        // become resident before the collective's kernel has left, and the rebuild would queue up behind the broadcast.
        const int sms = ctx->sm_count - ctx->reserved_sms > 8 ? ctx->sm_count - ctx->reserved_sms : 8;
        uint32_t grid = (uint32_t)(sms * per_sm);
        if (grid > tiles) grid = tiles;
        uint32_t n_arg = n; int radius_arg = radius;
        uint32_t* ctl0 = b->ctl;
        void* args[] = {&n_arg, &radius_arg, &b->cin, &b->cout, &b->node, &b->isize, &b->state, &b->tile_status,
                        &b->trace_active, &b->trace_merges, &b->iter_first_id, &ctl0, &b->iter_ns};
        RTR_PROF(ctx, "ploc_loop_kernel");
        RTR_CUDA(ctx, cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kPT), args, 0, ctx->stream));
        RTR_LAUNCH_CHECK(ctx);
    }
    RTR_PROF(ctx, "ploc_tail_kernel");
    ploc_tail_kernel<<<1, kTailN, sizeof(TailSmem), ctx->stream>>>(n, radius, b->cin, b->cout, b->node, b->isize, b->state,
                                                                   b->trace_active, b->trace_merges, b->iter_first_id, b->iter_ns);
    RTR_LAUNCH_CHECK(ctx);
    RTR_CHECK(record(b, 4));

    // 5. flatten, last creation level first
    if (n == 1) {
        flatten_single_leaf_kernel<<<1, 1, 0, ctx->stream>>>(b->node, b->wtri, b->flat, b->pairs);
        RTR_LAUNCH_CHECK(ctx);
    } else {
        if (ctx->flatten_ctas_per_sm == 0) {
            int c = 0;
            RTR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, flatten_positions_kernel, kFlattenBlock, 0));
            ctx->flatten_ctas_per_sm = c < 1 ? 1 : c;
        }
        const int sms = ctx->sm_count - ctx->reserved_sms > 8 ? ctx->sm_count - ctx->reserved_sms : 8;  // see the PLOC loop
        uint32_t grid = (uint32_t)(sms * ctx->flatten_ctas_per_sm);
        const uint32_t need = (n + kFlattenBlock - 1) / kFlattenBlock;
        if (grid > need) grid = need;
        uint32_t n_arg = n;
        uint32_t* ctl1 = b->ctl + 1;
        void* args[] = {&b->state, &b->iter_first_id, &n_arg, &b->node, &b->isize, &b->ipos, &b->order, &ctl1};
        RTR_PROF(ctx, "flatten_positions_kernel");
        RTR_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)flatten_positions_kernel, dim3(grid), dim3(kFlattenBlock), args, 0, ctx->stream));
        RTR_LAUNCH_CHECK(ctx);
        const uint32_t nc = 2 * n - 1;
        RTR_PROF(ctx, "flatten_emit_kernel");
        flatten_emit_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(b->state, nc, n, b->order, b->node, b->isize, b->wtri, b->flat, b->pairs);
        RTR_LAUNCH_CHECK(ctx);
    }
    b->pairs_view = b->pairs;  // written by the flatten kernels
    RTR_CHECK(record(b, 5));
    b->flat_view = b->flat;
    b->wtri_view = b->wtri;
    b->wtri_by_rank = true;
    b->built = true;            // optimistic: the verdict is still on the device
    b->finish_pending = true;
    return RTR_OK;
}

// The part of a build the host has to wait for: did PLOC converge, and in how many iterations.  Called by everything
// that hands results to the host (accessors, host-pointer trace calls, the synchronous rtr_bvh_build).
int rtr_bvh_finish(rtr_bvh* b) {
    if (!b->finish_pending) return RTR_OK;
    rtr_ctx* ctx = b->ctx;
    b->finish_pending = false;
    PlocState* h = static_cast<PlocState*>(ctx->pinned);
    RTR_CUDA(ctx, cudaMemcpyAsync(h, &b->state[2], sizeof(PlocState), cudaMemcpyDeviceToHost, ctx->stream));
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    b->iterations = h->iter;
    if (h->n_active != 1 || h->total != 2 * b->n - 1) {
        b->built = false;
        if (h->iter >= kMaxPlocIterations)
            return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "more than %u PLOC iterations (degenerate input)", kMaxPlocIterations);
        return rtr_set_error(ctx, RTR_E_INVALID, "PLOC did not converge: %u clusters left, %u created after %u iterations (non-finite triangle data?)",
                             h->n_active, h->total, h->iter);
    }
    return RTR_OK;
}

int rtr_bvh_export_clusters(rtr_bvh* b, rtr_node* clusters, uint32_t* parent, uint32_t* left, uint32_t* right,
                            uint8_t* is_leaf) {
    rtr_ctx* ctx = b->ctx;
    const uint32_t nc = 2 * b->n - 1;
    if (parent) RTR_CUDA(ctx, cudaMemsetAsync(parent, 0xFF, sizeof(uint32_t) * nc, ctx->stream));
    export_clusters_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(b->n, b->node, clusters, parent, left,
                                                                     right, is_leaf);
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}
