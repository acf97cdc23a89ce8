// ploc.cu -- PLOC bottom-up BVH construction and DFS flattening, bit-exact with the serial
// (OMP_NUM_THREADS=1) reference.
//
// Replaces BVH::plocPreprocessing (bvh.cpp:26-46), BVH::ploc (:48-123) with its four phases
// plocNearestNeighborSearch (:193-210), plocMerging (:159-191), plocPrefixScan (:125-148),
// plocCompaction (:150-156), and glr::Scene::getBVH_NodesToGPUData (scene.cpp:189-208).
//
// Layout in HBM (by cluster id, SoA of two float4 so a box is two 16-byte loads):
//   node_lo[c] = (min.x, min.y, min.z, max.x)      node_hi[c] = (max.y, max.z, L, R)
//   leaves c < n:  L = triangle id, R = RTR_NONE;  internal c >= n: L/R = child cluster ids
//   isize[c-n] = nodes in the subtree of internal cluster c (lets flatten place the right child)
//   cin/cout   = active cluster ids in Morton order (the reference's C_In/C_Out)
//
// One kernel per PLOC iteration (ploc_iteration_kernel): each 512-thread CTA takes a tile of 480
// positions plus a 32-position halo on each side, gathers ids + boxes into shared memory, finds
// every nearest neighbour in the +-16 window (argmin of merged surface area, lowest j on ties,
// fp32 ops in the reference's order), decides mutual pairs, ranks merges and survivors with a
// CTA scan + decoupled look-back across tiles, writes the merged nodes at id = total + rank
// (== the reference's serial counter, Q3) and compacts survivors -- NN search, merge, prefix scan
// and compaction of the reference fused into one pass over the active list.  The iteration
// count is data dependent, so the loop state lives on the device (PlocState, double buffered by
// launch parity) and launches that find n_active <= kTailN are no-ops; a single-CTA tail kernel
// finishes the last <= 1024 clusters entirely in shared memory (PLOC++ 4.4).
#include "bvh.cuh"

namespace {

constexpr int kR = RTR_MAX_SEARCH_RADIUS;       // 16
constexpr int kPBlock = 512;
constexpr int kTileT = kPBlock - 2 * kR;        // 480 positions decided per CTA
constexpr int kExt = kTileT + 4 * kR;           // 544 positions staged per CTA
constexpr uint32_t kTailN = 1024;
constexpr int kChunk = 8;                        // iteration launches between host checks

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

struct BoxSoA {  // views into shared memory
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

// ---------------------------------------------------------------------------------------
// leaf initialisation: plocPreprocessing (bvh.cpp:29-42) + AABB::buildFromTriangle (:383-399)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
leaf_init_kernel(const rtr_triangle* __restrict__ tris, const rtr_mesh* __restrict__ meshes,
                 const uint32_t* __restrict__ tri_idx, uint32_t n,
                 float4* __restrict__ node_lo, float4* __restrict__ node_hi, uint32_t* __restrict__ cin,
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
        node_lo[i] = make_float4(mnx, mny, mnz, mxx);
        node_hi[i] = make_float4(mxy, mxz, __uint_as_float(ti), __uint_as_float(RTR_NONE));
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, o));
    if (lane_id() == 0 && e2 > 0.f) atomicMax(&tparams->emax2_ordered, float_to_ordered(e2));
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, o));
    if (lane_id() == 0 && e2 > 0.f) atomicMax(&tparams->emax2_ordered, float_to_ordered(e2));
}

__global__ void ploc_state_init_kernel(PlocState* state, uint32_t n, uint32_t* iter_first_id) {
    state[0].n_active = n; state[0].total = n; state[0].iter = 0; state[0].tile_counter = 0;
    state[1] = state[0];
    iter_first_id[0] = n;
}

// ---------------------------------------------------------------------------------------
// one PLOC iteration over the whole active list
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPBlock)
ploc_iteration_kernel(uint32_t launch_idx, uint32_t n_leaves, int radius,
                      uint32_t* __restrict__ buf0, uint32_t* __restrict__ buf1,
                      float4* __restrict__ node_lo, float4* __restrict__ node_hi, uint32_t* __restrict__ isize,
                      PlocState* __restrict__ state, uint64_t* __restrict__ tile_status,
                      uint32_t* __restrict__ trace_active, uint32_t* __restrict__ trace_merges,
                      uint32_t* __restrict__ iter_first_id) {
    __shared__ float s_box[6][kExt];
    __shared__ uint32_t s_id[kExt];
    __shared__ int s_nn[kPBlock];        // neighbour position of q = t0 - kR + x, or -1
    __shared__ uint32_t s_warp[kPBlock / 32 + 1];
    __shared__ uint32_t s_tile, s_ex_lo, s_ex_hi;

    const uint32_t tid = threadIdx.x;
    PlocState* cur = &state[launch_idx & 1u];
    PlocState* nxt = &state[(launch_idx + 1u) & 1u];
    const uint32_t n = cur->n_active, total = cur->total, iter = cur->iter;
    if (n <= kTailN) {  // finished (or the tail kernel's job): carry the state to the next launch
        if (blockIdx.x == 0 && tid == 0) {
            nxt->n_active = n; nxt->total = total; nxt->iter = iter; nxt->tile_counter = 0;
        }
        return;
    }
    const uint32_t tiles = (n + kTileT - 1) / kTileT;
    if (tid == 0) s_tile = atomicAdd(&cur->tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= tiles) return;

    const uint32_t* __restrict__ cin = (iter & 1u) ? buf1 : buf0;
    uint32_t* __restrict__ cout = (iter & 1u) ? buf0 : buf1;
    const int t0 = (int)(tile * kTileT);
    const int base = t0 - 2 * kR;  // position of shared slot 0

    // ---- stage ids + boxes of the tile and its halo ----
    for (int e = tid; e < kExt; e += kPBlock) {
        const int pos = base + e;
        if (pos >= 0 && pos < (int)n) {
            const uint32_t id = cin[pos];
            const float4 lo = node_lo[id], hi = node_hi[id];
            s_id[e] = id;
            s_box[0][e] = lo.x; s_box[1][e] = lo.y; s_box[2][e] = lo.z;
            s_box[3][e] = lo.w; s_box[4][e] = hi.x; s_box[5][e] = hi.y;
        } else {
            s_id[e] = RTR_NONE;
        }
    }
    __syncthreads();

    // ---- nearest neighbour of every position in [t0 - kR, t0 + kTileT + kR) ----
    const BoxSoA B{s_box[0], s_box[1], s_box[2], s_box[3], s_box[4], s_box[5]};
    const int q = t0 - kR + (int)tid;  // this thread's position
    {
        int nn = -1;
        if (q >= 0 && q < (int)n) {
            const int jlo = max(0, q - radius), jhi = min(q + radius + 1, (int)n);
            const int e = nearest_neighbour(B, q - base, jlo - base, jhi - base);
            nn = (e < 0) ? -1 : e + base;
        }
        s_nn[tid] = nn;
    }
    __syncthreads();

    // ---- mutual pairs (plocMerging, bvh.cpp:159-165): lo = lower partner, hi = removed partner ----
    const bool in_tile = (tid >= (uint32_t)kR) && (tid < (uint32_t)(kR + kTileT)) && (q < (int)n);
    const int j = in_tile ? s_nn[tid] : -1;
    bool is_lo = false, is_hi = false;
    if (j >= 0) {
        const int nnj = s_nn[j - (t0 - kR)];
        if (nnj == q) { is_lo = q < j; is_hi = q > j; }
    }
    uint32_t cta_total;
    const uint32_t packed = (is_lo ? 1u : 0u) | (is_hi ? 1u << 16 : 0u);
    const uint32_t incl = block_scan_incl<kPBlock>(packed, s_warp, &cta_total);
    const uint32_t ex_lo_local = (incl & 0xFFFFu) - (is_lo ? 1u : 0u);
    const uint32_t ex_hi_local = (incl >> 16) - (is_hi ? 1u : 0u);

    // ---- decoupled look-back over tiles (plocPrefixScan, bvh.cpp:125-148, as a single pass) ----
    if (tid == 0) {
        const uint32_t agg_lo = cta_total & 0xFFFFu, agg_hi = cta_total >> 16;
        const uint32_t tag = iter + 1u;
        uint32_t ex_lo = 0, ex_hi = 0;
        if (tile > 0) {
            st_relaxed_u64(&tile_status[tile], lb_pack(tag, kLbAgg, agg_lo, agg_hi));
            int t = (int)tile - 1;
            while (true) {
                const uint64_t w = ld_relaxed_u64(&tile_status[t]);
                const uint64_t flag = (w >> 54) & 3ull;
                if ((uint32_t)(w >> 56) != (tag & 0xFFu) || flag == 0) continue;
                ex_lo += (uint32_t)(w & kCntMask);
                ex_hi += (uint32_t)((w >> 27) & kCntMask);
                if (flag == kLbIncl) break;
                --t;
            }
        }
        st_relaxed_u64(&tile_status[tile], lb_pack(tag, kLbIncl, ex_lo + agg_lo, ex_hi + agg_hi));
        s_ex_lo = ex_lo; s_ex_hi = ex_hi;
        if (tile == tiles - 1) {  // bvh.cpp:105-113
            const uint32_t merges = ex_lo + agg_lo, removed = ex_hi + agg_hi;
            nxt->n_active = n - removed; nxt->total = total + merges; nxt->iter = iter + 1; nxt->tile_counter = 0;
            if (iter < kMaxPlocIterations) {
                trace_active[iter] = n; trace_merges[iter] = merges; iter_first_id[iter + 1] = total + merges;
            }
        }
    }
    __syncthreads();

    // ---- merge (bvh.cpp:166-188) + compaction (bvh.cpp:150-156) ----
    if (in_tile) {
        const int e_q = q - base;
        uint32_t out_id = s_id[e_q];
        if (is_lo) {
            const int e_j = j - base;
            const uint32_t new_id = total + s_ex_lo + ex_lo_local;
            const uint32_t cl = s_id[e_q], cr = s_id[e_j];
            node_lo[new_id] = make_float4(fminf(B.minx[e_q], B.minx[e_j]), fminf(B.miny[e_q], B.miny[e_j]),
                                          fminf(B.minz[e_q], B.minz[e_j]), fmaxf(B.maxx[e_q], B.maxx[e_j]));
            node_hi[new_id] = make_float4(fmaxf(B.maxy[e_q], B.maxy[e_j]), fmaxf(B.maxz[e_q], B.maxz[e_j]),
                                          __uint_as_float(cl), __uint_as_float(cr));
            const uint32_t sl = cl < n_leaves ? 1u : isize[cl - n_leaves];
            const uint32_t sr = cr < n_leaves ? 1u : isize[cr - n_leaves];
            isize[new_id - n_leaves] = sl + sr + 1u;
            out_id = new_id;
        }
        if (!is_hi) cout[(uint32_t)q - (s_ex_hi + ex_hi_local)] = out_id;
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
ploc_tail_kernel(uint32_t launch_idx, uint32_t n_leaves, int radius,
                 const uint32_t* __restrict__ buf0, const uint32_t* __restrict__ buf1,
                 float4* __restrict__ node_lo, float4* __restrict__ node_hi, uint32_t* __restrict__ isize,
                 PlocState* __restrict__ state, uint32_t* __restrict__ trace_active,
                 uint32_t* __restrict__ trace_merges, uint32_t* __restrict__ iter_first_id) {
    extern __shared__ __align__(16) unsigned char tail_raw[];
    TailSmem& s = *reinterpret_cast<TailSmem*>(tail_raw);
    const uint32_t tid = threadIdx.x;
    PlocState* cur = &state[launch_idx & 1u];
    PlocState* nxt = &state[(launch_idx + 1u) & 1u];
    uint32_t n = cur->n_active, total = cur->total, iter = cur->iter;
    if (n > kTailN) {  // not ours yet: carry the state
        if (tid == 0) { nxt->n_active = n; nxt->total = total; nxt->iter = iter; nxt->tile_counter = 0; }
        return;
    }
    const uint32_t* __restrict__ cin = (iter & 1u) ? buf1 : buf0;
    int cur_buf = 0;
    if (tid < n) {
        const uint32_t id = cin[tid];
        const float4 lo = node_lo[id], hi = node_hi[id];
        s.id[0][tid] = id;
        s.size[0][tid] = id < n_leaves ? 1u : isize[id - n_leaves];
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
                node_lo[new_id] = make_float4(mnx, mny, mnz, mxx);
                node_hi[new_id] = make_float4(mxy, mxz, __uint_as_float(cl), __uint_as_float(cr));
                isize[new_id - n_leaves] = sz;
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
        }
        n -= removed; total += merges; iter += 1; cur_buf = nb;
        __syncthreads();
        if (merges == 0) break;  // non-finite areas (Q4): the reference would spin forever
    }
    if (tid == 0) { nxt->n_active = n; nxt->total = total; nxt->iter = iter; nxt->tile_counter = 0; }
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

__device__ __forceinline__ void flatten_one(uint32_t c, uint32_t n_leaves,
                                            const float4* __restrict__ node_lo, const float4* __restrict__ node_hi,
                                            const uint32_t* __restrict__ isize, uint32_t* __restrict__ ipos,
                                            rtr_node* __restrict__ flat) {
    const uint32_t p = ipos[c - n_leaves];
    const float4 lo = node_lo[c], hi = node_hi[c];
    const uint32_t L = __float_as_uint(hi.z), R = __float_as_uint(hi.w);
    const uint32_t size_l = L < n_leaves ? 1u : isize[L - n_leaves];
    const uint32_t pos_l = p + 1u, pos_r = p + 1u + size_l;
    store_node(flat, p, lo, hi, 0u, pos_l, pos_r);  // internal _TriangleId stays 0 (bvh.cpp:415-420)
    if (L < n_leaves) {
        const float4 llo = node_lo[L], lhi = node_hi[L];
        store_node(flat, pos_l, llo, lhi, __float_as_uint(lhi.z), 0u, 0u, L);
    } else {
        ipos[L - n_leaves] = pos_l;
    }
    if (R < n_leaves) {
        const float4 rlo = node_lo[R], rhi = node_hi[R];
        store_node(flat, pos_r, rlo, rhi, __float_as_uint(rhi.z), 0u, 0u, R);
    } else {
        ipos[R - n_leaves] = pos_r;
    }
}

__global__ void __launch_bounds__(256)
flatten_level_kernel(uint32_t first, uint32_t count, uint32_t n_leaves,
                     const float4* __restrict__ node_lo, const float4* __restrict__ node_hi,
                     const uint32_t* __restrict__ isize, uint32_t* __restrict__ ipos, rtr_node* __restrict__ flat) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) flatten_one(first + i, n_leaves, node_lo, node_hi, isize, ipos, flat);
}

// levels [it_lo, it_hi] processed last-first by one CTA (the top of the tree: many tiny levels)
__global__ void __launch_bounds__(1024)
flatten_small_levels_kernel(const uint32_t* __restrict__ iter_first_id, int it_hi, int it_lo, uint32_t n_leaves,
                            const float4* __restrict__ node_lo, const float4* __restrict__ node_hi,
                            const uint32_t* __restrict__ isize, uint32_t* __restrict__ ipos, rtr_node* __restrict__ flat) {
    for (int it = it_hi; it >= it_lo; --it) {
        const uint32_t first = iter_first_id[it], count = iter_first_id[it + 1] - first;
        for (uint32_t i = threadIdx.x; i < count; i += blockDim.x)
            flatten_one(first + i, n_leaves, node_lo, node_hi, isize, ipos, flat);
        __syncthreads();
    }
}

__global__ void flatten_single_leaf_kernel(const float4* node_lo, const float4* node_hi, rtr_node* flat) {
    const float4 lo = node_lo[0], hi = node_hi[0];
    store_node(flat, 0, lo, hi, __float_as_uint(hi.z), 0u, 0u, 0u);
}

// child-pair records for the traversal kernels (layout in bvh.cuh): pure repacking of the flat array
__global__ void __launch_bounds__(256)
pack_pairs_kernel(const rtr_node* __restrict__ flat, uint32_t nb_nodes, uint32_t by_rank, uint4* __restrict__ pairs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb_nodes) return;
    const uint4 links = __ldg(reinterpret_cast<const uint4*>(flat + i) + 2);
    if (links.y == 0u && links.z == 0u) return;  // leaf: no record
    const uint4* l = reinterpret_cast<const uint4*>(flat + links.y);
    const uint4* r = reinterpret_cast<const uint4*>(flat + links.z);
    const uint4 l0 = __ldg(l), l1 = __ldg(l + 1), l2 = __ldg(l + 2);
    const uint4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
    const bool lleaf = l2.y == 0u && l2.z == 0u, rleaf = r2.y == 0u && r2.z == 0u;
    uint4* dst = pairs + (size_t)i * 4;
    dst[0] = make_uint4(l0.x, l0.y, l0.z, l1.x);
    dst[1] = make_uint4(l1.y, l1.z, r0.x, r0.y);
    dst[2] = make_uint4(r0.z, r1.x, r1.y, r1.z);
    dst[3] = make_uint4(lleaf ? (0x80000000u | links.y) : links.y, rleaf ? (0x80000000u | links.z) : links.z,
                        lleaf ? (by_rank ? l2.w : l2.x) : 0u, rleaf ? (by_rank ? r2.w : r2.x) : 0u);
}

// ---------------------------------------------------------------------------------------
// BVH_Params view by cluster id for the accessors / the cr::BVH shim (not on the timed path)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
export_clusters_kernel(uint32_t n_leaves, const float4* __restrict__ node_lo, const float4* __restrict__ node_hi,
                       rtr_node* __restrict__ clusters, uint32_t* __restrict__ parent, uint32_t* __restrict__ left,
                       uint32_t* __restrict__ right, uint8_t* __restrict__ is_leaf) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * n_leaves - 1) return;
    const float4 lo = node_lo[c], hi = node_hi[c];
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
    return rtr_bvh_pack_pairs_own(b);
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
    pack_pairs_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(b->flat_view, nc, b->wtri_by_rank ? 1u : 0u, b->pairs_own);
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
    RTR_CHECK(rtr_morton_launch(ctx, b->tris, n, b->array_len, b->meshes, b->codes, b->tri_idx, nullptr,
                                b->bounds12, b->ordered6));
    RTR_CHECK(record(b, 1));
    // 2. stable sort of (code, index); Morton codes use bits [0,30)
    RTR_CHECK(rtr_sort_impl_u32(ctx, b->codes, b->tri_idx, n, 0, 32));
    RTR_CHECK(record(b, 2));
    // 3. leaves
    RTR_CUDA(ctx, cudaMemsetAsync(b->tparams, 0, sizeof(TraceParams), ctx->stream));
    RTR_PROF(ctx, "leaf_init_kernel");
    leaf_init_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(b->tris, b->meshes, b->tri_idx, n, b->node_lo,
                                                               b->node_hi, b->cin, b->wtri, b->tparams);
    RTR_LAUNCH_CHECK(ctx);
    ploc_state_init_kernel<<<1, 1, 0, ctx->stream>>>(b->state, n, b->iter_first_id);
    RTR_LAUNCH_CHECK(ctx);
    RTR_CHECK(record(b, 3));

    // 4. PLOC loop
    PlocState* h_state = static_cast<PlocState*>(ctx->pinned);
    uint32_t launch_idx = 0;
    uint32_t bound_n = n;  // upper bound of n_active, refreshed from the device every kChunk launches
    static bool tail_configured = false;
    if (!tail_configured) {
        RTR_CUDA(ctx, cudaFuncSetAttribute(ploc_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(TailSmem)));
        tail_configured = true;
    }
    uint32_t last_iter = 0;
    while (bound_n > kTailN) {
        const uint32_t tiles = (bound_n + kTileT - 1) / kTileT;
        for (int k = 0; k < kChunk; ++k) {
            RTR_PROF(ctx, "ploc_iteration_kernel");
            ploc_iteration_kernel<<<tiles, kPBlock, 0, ctx->stream>>>(
                launch_idx, n, radius, b->cin, b->cout, b->node_lo, b->node_hi, b->isize, b->state, b->tile_status,
                b->trace_active, b->trace_merges, b->iter_first_id);
            RTR_LAUNCH_CHECK(ctx);
            ++launch_idx;
        }
        RTR_CUDA(ctx, cudaMemcpyAsync(h_state, &b->state[launch_idx & 1u], sizeof(PlocState), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h_state->n_active > kTailN && h_state->iter == last_iter)
            return rtr_set_error(ctx, RTR_E_INVALID, "PLOC made no progress (non-finite triangle data?)");
        if (h_state->n_active >= bound_n && h_state->n_active > kTailN)
            return rtr_set_error(ctx, RTR_E_INVALID, "PLOC iteration merged nothing (non-finite triangle data?)");
        if (h_state->iter >= kMaxPlocIterations)
            return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "more than %u PLOC iterations (degenerate input)", kMaxPlocIterations);
        last_iter = h_state->iter;
        bound_n = h_state->n_active;
    }
    RTR_PROF(ctx, "ploc_tail_kernel");
    ploc_tail_kernel<<<1, kTailN, sizeof(TailSmem), ctx->stream>>>(launch_idx, n, radius, b->cin, b->cout, b->node_lo,
                                                                   b->node_hi, b->isize, b->state, b->trace_active,
                                                                   b->trace_merges, b->iter_first_id);
    RTR_LAUNCH_CHECK(ctx);
    ++launch_idx;
    RTR_CUDA(ctx, cudaMemcpyAsync(h_state, &b->state[launch_idx & 1u], sizeof(PlocState), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_state->n_active != 1 || h_state->total != 2 * n - 1)
        return rtr_set_error(ctx, RTR_E_INVALID, "PLOC did not converge: %u clusters left, %u created (non-finite input?)",
                             h_state->n_active, h_state->total);
    if (h_state->iter > kMaxPlocIterations)
        return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "more than %u PLOC iterations (degenerate input)", kMaxPlocIterations);
    b->iterations = h_state->iter;
    b->h_first_id.resize(b->iterations + 1);
    if (b->iterations)
        RTR_CUDA(ctx, cudaMemcpyAsync(b->h_first_id.data(), b->iter_first_id, sizeof(uint32_t) * (b->iterations + 1),
                                      cudaMemcpyDeviceToHost, ctx->stream));
    RTR_CHECK(record(b, 4));
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

    // 5. flatten, last creation level first
    if (n == 1) {
        flatten_single_leaf_kernel<<<1, 1, 0, ctx->stream>>>(b->node_lo, b->node_hi, b->flat);
        RTR_LAUNCH_CHECK(ctx);
    } else {
        RTR_CUDA(ctx, cudaMemsetAsync(b->ipos + (n - 2), 0, sizeof(uint32_t), ctx->stream));  // root 2n-2 -> position 0
        const uint32_t kSmall = 4096;
        int it = (int)b->iterations - 1;
        while (it >= 0) {
            const uint32_t count = b->h_first_id[it + 1] - b->h_first_id[it];
            if (count > kSmall) {
                RTR_PROF(ctx, "flatten_level_kernel");
                flatten_level_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(b->h_first_id[it], count, n, b->node_lo,
                                                                                   b->node_hi, b->isize, b->ipos, b->flat);
                RTR_LAUNCH_CHECK(ctx);
                --it;
            } else {
                int lo = it;
                while (lo - 1 >= 0 && b->h_first_id[lo] - b->h_first_id[lo - 1] <= kSmall) --lo;
                RTR_PROF(ctx, "flatten_small_levels_kernel");
                flatten_small_levels_kernel<<<1, 1024, 0, ctx->stream>>>(b->iter_first_id, it, lo, n, b->node_lo,
                                                                         b->node_hi, b->isize, b->ipos, b->flat);
                RTR_LAUNCH_CHECK(ctx);
                it = lo - 1;
            }
        }
    }
    {   // child-pair records of the default traversal (part of the flatten stage time)
        const uint32_t nc = 2 * n - 1;
        RTR_PROF(ctx, "pack_pairs_kernel");
        pack_pairs_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(b->flat, nc, 1u, b->pairs);
        RTR_LAUNCH_CHECK(ctx);
        b->pairs_view = b->pairs;
    }
    RTR_CHECK(record(b, 5));
    b->flat_view = b->flat;
    b->wtri_view = b->wtri;
    b->wtri_by_rank = true;
    b->built = true;
    return RTR_OK;
}

int rtr_bvh_export_clusters(rtr_bvh* b, rtr_node* clusters, uint32_t* parent, uint32_t* left, uint32_t* right,
                            uint8_t* is_leaf) {
    rtr_ctx* ctx = b->ctx;
    const uint32_t nc = 2 * b->n - 1;
    if (parent) RTR_CUDA(ctx, cudaMemsetAsync(parent, 0xFF, sizeof(uint32_t) * nc, ctx->stream));
    export_clusters_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(b->n, b->node_lo, b->node_hi, clusters, parent, left,
                                                                     right, is_leaf);
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}
