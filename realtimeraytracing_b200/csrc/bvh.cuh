// bvh.cuh -- device-side layout of a BVH under construction and of a built BVH.
#pragma once
#include "common.cuh"

// PLOC loop state, double buffered by launch parity (see ploc.cu)
struct PlocState {
    uint32_t n_active;      // clusters alive at the start of the iteration
    uint32_t total;         // clusters created so far (next new id)
    uint32_t iter;          // iterations completed
    uint32_t tile_counter;  // dynamic tile ids of the launch that reads this slot
};

// scene constants the traversal kernels need (device resident so an adopted BVH can carry them)
struct TraceParams {
    uint32_t emax2_ordered;  // max squared edge length over all triangles, ordered-uint encoded
    uint32_t stack_overflows;
    uint32_t job_counter;    // next 32-job tile of the persistent traversal launch in flight
    uint32_t bad_layout;     // pack_pairs_kernel: inner nodes whose left child is not at index + 1 (not DFS pre-order)
};

constexpr uint32_t kMaxPlocIterations = 1u << 16;
#ifndef RTR_PLOC_WARPS
#define RTR_PLOC_WARPS 8
#endif
// positions decided per CTA of ploc_iteration_kernel (ploc.cu): every warp finds the nearest neighbour of
// 112 positions, the first and last 16 of the CTA's range only serve the mutual-pair test of the others
constexpr int kPlocTile = 112 * RTR_PLOC_WARPS - 32;

struct rtr_bvh {
    rtr_ctx* ctx = nullptr;
    uint32_t n = 0;          // triangles
    uint32_t capacity = 0;   // triangles the arrays below were sized for
    uint32_t array_len = 0;
    uint32_t nb_meshes = 0;
    uint32_t radius = RTR_DEFAULT_SEARCH_RADIUS;
    bool adopted = false;    // flat/tris/meshes are borrowed, no build arrays
    bool trav_only = false;  // received by rtr_bvh_broadcast_traversal: traversal records + node 0 only
    bool built = false;

    // inputs (device)
    const rtr_triangle* tris = nullptr;
    const rtr_mesh* meshes = nullptr;
    rtr_triangle* tris_own = nullptr;  size_t tris_own_cap = 0;   // triangles
    rtr_mesh* meshes_own = nullptr;    size_t meshes_own_cap = 0; // meshes

    // build arrays (device), sized by capacity
    uint32_t key_bits = 32;        // Morton key width of the last build: 32 (30-bit codes, the reference) or 64 (63-bit)
    uint64_t* codes64 = nullptr;   // [cap]  sorted 63-bit Morton codes (64-bit builds only, allocated on first use)
    size_t codes64_cap = 0;
    uint32_t* codes = nullptr;     // [cap]  sorted Morton codes
    uint32_t* tri_idx = nullptr;   // [cap]  BVH_Params::_TriangleIndices
    float4* node = nullptr;        // [2*(2cap-1)] by cluster id, 32 B records: (min.xyz, max.x)(max.y, max.z, bits(left | triangle id), bits(right | NONE))
    uint2* isize = nullptr;        // [cap] internal cluster (id - n): (nodes in its subtree, nodes in its left subtree)
    uint32_t* ipos = nullptr;      // [cap] DFS pre-order position of internal cluster (id - n)
    uint32_t* order = nullptr;     // [2cap-1] cluster id at every DFS pre-order position (flatten pass 1)
    uint32_t* cin = nullptr;       // [cap] active list, ping
    uint32_t* cout = nullptr;      // [cap] active list, pong
    uint64_t* tile_status = nullptr;  // [ceil(cap/tile)] decoupled look-back words
    PlocState* state = nullptr;    // [3]: loop state by iteration parity, [2] = final (see ploc.cu)
    uint32_t* ctl = nullptr;       // [4] barrier counters of the persistent build kernels
    uint32_t* trace_active = nullptr;   // [kMaxPlocIterations]
    uint32_t* trace_merges = nullptr;   // [kMaxPlocIterations]
    uint32_t* iter_first_id = nullptr;  // [kMaxPlocIterations + 1] first cluster id created by iteration i
    unsigned long long* iter_ns = nullptr;  // [kMaxPlocIterations + 2] %globaltimer at the start of iteration i (diagnostics)
    float* bounds12 = nullptr;     // scene box + cube
    uint32_t* ordered6 = nullptr;
    rtr_node* flat = nullptr;      // [2cap-1] DFS pre-order, the reference's SSBO 5
    const rtr_node* flat_view = nullptr;  // what traversal reads (== flat, flat_recv, or an adopted pointer)
    rtr_node* flat_recv = nullptr; // receive buffer of rtr_bvh_broadcast on non-root ranks
    size_t recv_cap = 0;           // triangles flat_recv/tris_own were sized for by a broadcast
    TraceParams* tparams = nullptr;
    // world-space triangle cache read by the traversal kernels: 3 float4 per triangle
    // (P0.xyz,P1.x)(P1.yz,P2.xy)(P2.z,-,-,-).  Built BVH: slot = leaf cluster id (Morton rank), which the
    // flatten also stores in the leaf's third padding word; adopted nodes: slot = triangle id.
    float4* wtri = nullptr;             // [3*cap] written by leaf_init_kernel
    float4* wtri_own = nullptr;         // adopted / received BVHs
    size_t wtri_own_cap = 0;            // triangles wtri_own was sized for
    const float4* wtri_view = nullptr;  // what traversal reads
    bool wtri_by_rank = true;
    // traversal records read by the default traversal, 64 bytes per flat index (layout: TravRec below)
    uint4* pairs = nullptr;        // [4*(2cap-1)]
    uint4* pairs_own = nullptr;    // adopted / received BVHs
    size_t pairs_own_cap = 0;      // triangles pairs_own was sized for
    const uint4* pairs_view = nullptr;

    // host mirrors of the last build
    uint32_t iterations = 0;       // valid after rtr_bvh_finish
    bool finish_pending = false;   // the last build's verdict has not been read back yet
    bool timing = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float stage_ms[6] = {0, 0, 0, 0, 0, 0};
};

// ---------------------------------------------------------------------------------------
// Traversal records: the default traversal's private copy of the flat array, 64 bytes per flat index.
//
//   inner node p (first 32 bytes, one 256-bit load):
//     w0..w2  origin = min corner of the node's own box (fp32)
//     w3      bytes (Ex, Ey, Ez, flags): grid step of axis k = 2^(Ek - 127);
//             flags bit0 / bit1: left / right child is a leaf, bit2: no culling at this node (box outside
//             the supported range)
//     w4..w6  per axis k the byte quadruple (Llo, Lhi, Rlo, Rhi): child plane = origin_k + byte * step_k,
//             min planes rounded down, max planes rounded up -- a SUPERSET of the child's exact box
//     w7      flat index of the right child (the left child is p + 1: DFS pre-order, scene.cpp:189-199)
//   inner node p, second 32 bytes (the "wide step", one more 256-bit load): four slots on the SAME grid --
//     slot 0 / 1 = the children of L (slot 0 = L itself and slot 1 unused when L is a leaf), slot 2 / 3 = those of R
//     w8..w10   per axis the byte quadruple (S0lo, S0hi, S1lo, S1hi);  w11  flat index of slot 1 (bit 31: it is a leaf)
//     w12..w14  per axis the byte quadruple (S2lo, S2hi, S3lo, S3hi);  w15  flat index of slot 3 (bit 31: it is a leaf)
//     (slot 0 is at p + 2, or p + 1 when L is a leaf; slot 2 at right + 1, or right when R is a leaf)
//     flags bit3..bit6: slot 0..3 is a leaf; bit7: the second half is usable
//   leaf p (64 bytes, two 256-bit loads):
//     exact box min.xyz, max.xyz | world-space vertices P0, P1, P2 (host naming) | triangle id
//
// Why a superset is enough: the reference tests a leaf's triangle iff the slab test passes on the leaf's own
// box (merged boxes are exact unions and the slab test is monotone in the box).  The leaf
// record keeps that box exactly and the leaf's test is the reference's; inner records only decide which
// leaves get there, so they may err on the side of visiting -- never on the side of skipping.
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
// The encoder and the decoder of the traversal records are __host__ __device__: the kernels use the intrinsics, and
// tests/host/trav_records_check.cu runs the very same functions on the CPU (directed rounding through <cfenv>, the
// fused multiply-add through fmaf) to check their contract -- "the reference's slab test passes on the exact box ==>
// the compressed test passes, with an entry distance that is not larger" -- on millions of random boxes and rays.
#ifndef __CUDA_ARCH__
#include <cfenv>
#include <cmath>
#include <cstring>
#endif
#define TRAV_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
TRAV_HD long long trav_d2ll(double x) { return __double_as_longlong(x); }
TRAV_HD double trav_ll2d(long long x) { return __longlong_as_double(x); }
TRAV_HD uint32_t trav_f2u(float x) { return __float_as_uint(x); }
TRAV_HD float trav_u2f(uint32_t x) { return __uint_as_float(x); }
TRAV_HD float trav_sub_rd(float a, float b) { return __fsub_rd(a, b); }
TRAV_HD float trav_sub_ru(float a, float b) { return __fsub_ru(a, b); }
TRAV_HD float trav_mul_rd(float a, float b) { return __fmul_rd(a, b); }
TRAV_HD float trav_mul_ru(float a, float b) { return __fmul_ru(a, b); }
TRAV_HD int trav_f2i_rd(float x) { return __float2int_rd(x); }
TRAV_HD int trav_f2i_ru(float x) { return __float2int_ru(x); }
TRAV_HD float trav_sub(float a, float b) { return __fsub_rn(a, b); }
TRAV_HD float trav_add(float a, float b) { return __fadd_rn(a, b); }
TRAV_HD float trav_mul(float a, float b) { return __fmul_rn(a, b); }
TRAV_HD float trav_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
TRAV_HD uint32_t trav_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
#else
inline long long trav_d2ll(double x) { long long r; std::memcpy(&r, &x, 8); return r; }
inline double trav_ll2d(long long x) { double r; std::memcpy(&r, &x, 8); return r; }
inline uint32_t trav_f2u(float x) { uint32_t r; std::memcpy(&r, &x, 4); return r; }
inline float trav_u2f(uint32_t x) { float r; std::memcpy(&r, &x, 4); return r; }
template <class F> inline float trav_rounded(int mode, F op) {  // one fp32 operation in the given rounding mode
    const int old = std::fegetround();
    std::fesetround(mode);
    volatile float r = op();
    std::fesetround(old);
    return r;
}
inline float trav_sub_rd(float a, float b) { volatile float x = a, y = b; return trav_rounded(FE_DOWNWARD, [&] { return x - y; }); }
inline float trav_sub_ru(float a, float b) { volatile float x = a, y = b; return trav_rounded(FE_UPWARD, [&] { return x - y; }); }
inline float trav_mul_rd(float a, float b) { volatile float x = a, y = b; return trav_rounded(FE_DOWNWARD, [&] { return x * y; }); }
inline float trav_mul_ru(float a, float b) { volatile float x = a, y = b; return trav_rounded(FE_UPWARD, [&] { return x * y; }); }
inline int trav_f2i_rd(float x) { return (int)std::floor(x); }
inline int trav_f2i_ru(float x) { return (int)std::ceil(x); }
inline float trav_sub(float a, float b) { volatile float r = a - b; return r; }
inline float trav_add(float a, float b) { volatile float r = a + b; return r; }
inline float trav_mul(float a, float b) { volatile float r = a * b; return r; }
inline float trav_fma(float a, float b, float c) { return std::fmaf(a, b, c); }
inline uint32_t trav_prmt(uint32_t a, uint32_t b, uint32_t sel) {  // PRMT, default mode: result byte i = byte sel[4i+2:4i] of {b, a}
    const unsigned long long both = ((unsigned long long)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((both >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
#endif
TRAV_HD int trav_ilogb(double x) {  // floor(log2 x) of a positive normal double
    return (int)((trav_d2ll(x) >> 52) & 0x7FF) - 1023;
}
TRAV_HD double trav_pow2(int e) { return trav_ll2d((long long)(e + 1023) << 52); }

// One axis of a node's grid: step 2^e with origin + 255 * step >= hi (the extent is taken in double, where hi - lo of
// two fp32 values is exact); the step is kept within 2^-40 of the origin's magnitude.
struct TravAxis {
    double lo, step, inv_step;
    uint32_t ebyte;
};
TRAV_HD TravAxis trav_axis_frame(float nlo_f, float nhi_f, bool& ok) {
    const double lo = nlo_f, hi = nhi_f;
    const double ext = hi - lo;
    int e = -100;
    if (ext > 0.0 && ext < 1e300) {
        e = trav_ilogb(ext) - 7;              // 256 * 2^e > ext
        if (255.0 * trav_pow2(e) < ext) ++e;  // smallest e with 255 * 2^e >= ext
    }
    if (lo != 0.0) {
        const int eo = trav_ilogb(fabs(lo)) + 1;
        if (e < eo - 40) e = eo - 40;
    }
    if (e < -100) e = -100;
    if (!(ext >= 0.0) || e > 30 || !(fabs(lo) < 1e15)) { ok = false; e = 0; }
    TravAxis f;
    f.lo = lo; f.step = trav_pow2(e); f.inv_step = trav_pow2(-e); f.ebyte = (uint32_t)(e + 127);
    return f;
}
// The four plane bytes (Alo, Ahi, Blo, Bhi) of two boxes inside the node: min planes rounded down, max planes up.
// fp32 with directed rounding: x - lo is rounded towards the safe side and the division by the power-of-two step is
// exact (rounded the same way where it leaves the normal range), so byte <= (x - lo) / step for a min plane and >= for
// a max plane hold in exact arithmetic -- no check in higher precision is needed, and at worst the byte is one step
// looser than the exactly rounded one.  (An earlier version did this in double with a verification pass: +0.2 ms per
// 10 M-triangle rebuild for the same traversal time.)
TRAV_HD uint32_t trav_axis_quad_f32(const TravAxis& f, float alo, float ahi, float blo, float bhi, bool& ok) {
    const float lo = (float)f.lo, inv_step = (float)f.inv_step;  // both exact: the origin is an fp32 value, the step 2^e, |e| <= 100
    auto down = [&](float x) -> uint32_t {
        const float t = trav_mul_rd(trav_sub_rd(x, lo), inv_step);
        return (uint32_t)trav_f2i_rd(fminf(fmaxf(t, 0.f), 255.f));
    };
    auto up = [&](float x) -> uint32_t {
        const float t = trav_mul_ru(trav_sub_ru(x, lo), inv_step);
        if (!(t <= 255.5f)) ok = false;  // beyond the grid (never for a box inside the node) or NaN
        return (uint32_t)trav_f2i_ru(fminf(fmaxf(t, 0.f), 255.f));
    };
    return down(alo) | (up(ahi) << 8) | (down(blo) << 16) | (up(bhi) << 24);
}
TRAV_HD void trav_encode_axis(float nlo_f, float nhi_f, float llo, float lhi, float rlo, float rhi,
                                                 uint32_t& ebyte, uint32_t& quad, bool& ok) {
    const TravAxis f = trav_axis_frame(nlo_f, nhi_f, ok);
    ebyte = f.ebyte;
    quad = trav_axis_quad_f32(f, llo, lhi, rlo, rhi, ok);
}

// node box n, child boxes l and r as (min.xyz, max.x)(max.y, max.z) float4 + float2
TRAV_HD void trav_encode_inner(const float4 nlo, const float2 nhi, const float4 llo, const float2 lhi,
                                                  const float4 rlo, const float2 rhi, bool lleaf, bool rleaf,
                                                  uint32_t right, uint4& o0, uint4& o1) {
    uint32_t ex, ey, ez, qx, qy, qz;
    bool ok = true;
    trav_encode_axis(nlo.x, nlo.w, llo.x, llo.w, rlo.x, rlo.w, ex, qx, ok);
    trav_encode_axis(nlo.y, nhi.x, llo.y, lhi.x, rlo.y, rhi.x, ey, qy, ok);
    trav_encode_axis(nlo.z, nhi.y, llo.z, lhi.y, rlo.z, rhi.y, ez, qz, ok);
    const uint32_t flags = (lleaf ? 1u : 0u) | (rleaf ? 2u : 0u) | (ok ? 0u : 4u);
    o0 = make_uint4(trav_f2u(nlo.x), trav_f2u(nlo.y), trav_f2u(nlo.z),
                    ex | (ey << 8) | (ez << 16) | (flags << 24));
    o1 = make_uint4(qx, qy, qz, right);
}
// Second half of an inner record: the four GRANDCHILD slots on the node's own grid (see the layout above).
// box k = (min.xyz, max.x)(max.y, max.z); an unused slot repeats its sibling.  ok = false: the node keeps to pair steps.
TRAV_HD void trav_encode_quads(const float4 nlo, const float2 nhi, const float4 lo[4], const float2 hi[4],
                                                  uint32_t idx1, uint32_t idx3, uint4& o2, uint4& o3, bool& ok) {
    const TravAxis fx = trav_axis_frame(nlo.x, nlo.w, ok);
    const TravAxis fy = trav_axis_frame(nlo.y, nhi.x, ok);
    const TravAxis fz = trav_axis_frame(nlo.z, nhi.y, ok);
    o2 = make_uint4(trav_axis_quad_f32(fx, lo[0].x, lo[0].w, lo[1].x, lo[1].w, ok), trav_axis_quad_f32(fy, lo[0].y, hi[0].x, lo[1].y, hi[1].x, ok),
                    trav_axis_quad_f32(fz, lo[0].z, hi[0].y, lo[1].z, hi[1].y, ok), idx1);
    o3 = make_uint4(trav_axis_quad_f32(fx, lo[2].x, lo[2].w, lo[3].x, lo[3].w, ok), trav_axis_quad_f32(fy, lo[2].y, hi[2].x, lo[3].y, hi[3].x, ok),
                    trav_axis_quad_f32(fz, lo[2].z, hi[2].y, lo[3].z, hi[3].y, ok), idx3);
}

// ---- decoder side (trace.cu) ----
// One axis of the compressed test.  With the node origin c, grid step s = 2^e, plane byte q and the ray (o, j = 1/d):
//     a = s*j (exact), b = fl(fl(c - o)*j), x = 2^23 + q (exact, built by PRMT), t = fl(x*a + fl(fl(b -+ m) - 2^23*a))
// with the margin m = 0.52|a| + 2e-6|b| + 1e-30 (derivation in trace.cu).  The bytes (lo, hi) of a slot become
// (near, far) by a byte swap when j < 0.
// pair step: quad = (Llo, Lhi, Rlo, Rhi); updates (near, far) of the left and of the right child
TRAV_HD void trav_axis_planes2(float c, uint32_t ebyte, uint32_t quad, float o, float j,
                               float& tnl, float& tfl, float& tnr, float& tfr) {
    const float step = trav_u2f(ebyte << 23);
    const float sa = trav_mul(step, j);
    const float sb = trav_mul(trav_sub(c, o), j);
    const float m = trav_fma(fabsf(sa), 0.52f, trav_fma(fabsf(sb), 2e-6f, 1e-30f));
    const float bn = trav_fma(-8388608.f, sa, trav_sub(sb, m));
    const float bf = trav_fma(-8388608.f, sa, trav_add(sb, m));
    // bytes (Llo, Lhi, Rlo, Rhi) -> (Lnear, Lfar, Rnear, Rfar)
    const uint32_t w = trav_prmt(quad, quad, j < 0.f ? 0x2301u : 0x3210u);
    const float x0 = trav_u2f(trav_prmt(w, 0x4B000000u, 0x7650u));
    const float x1 = trav_u2f(trav_prmt(w, 0x4B000000u, 0x7651u));
    const float x2 = trav_u2f(trav_prmt(w, 0x4B000000u, 0x7652u));
    const float x3 = trav_u2f(trav_prmt(w, 0x4B000000u, 0x7653u));
    tnl = fmaxf(tnl, trav_fma(x0, sa, bn)); tfl = fminf(tfl, trav_fma(x1, sa, bf));
    tnr = fmaxf(tnr, trav_fma(x2, sa, bn)); tfr = fminf(tfr, trav_fma(x3, sa, bf));
}
// wide step: quads (S0lo, S0hi, S1lo, S1hi) and (S2lo, S2hi, S3lo, S3hi); returns this axis' entry / exit distances of
// the four slots, the caller combines the axes
TRAV_HD void trav_axis_vals4(float c, uint32_t ebyte, uint32_t qa, uint32_t qb, float o, float j, float (&vn)[4], float (&vf)[4]) {
    const float step = trav_u2f(ebyte << 23);
    const float sa = trav_mul(step, j);
    const float sb = trav_mul(trav_sub(c, o), j);
    const float m = trav_fma(fabsf(sa), 0.52f, trav_fma(fabsf(sb), 2e-6f, 1e-30f));
    const float bn = trav_fma(-8388608.f, sa, trav_sub(sb, m));
    const float bf = trav_fma(-8388608.f, sa, trav_add(sb, m));
    const uint32_t sel = j < 0.f ? 0x2301u : 0x3210u;
    const uint32_t wa = trav_prmt(qa, qa, sel), wb = trav_prmt(qb, qb, sel);
    vn[0] = trav_fma(trav_u2f(trav_prmt(wa, 0x4B000000u, 0x7650u)), sa, bn);
    vf[0] = trav_fma(trav_u2f(trav_prmt(wa, 0x4B000000u, 0x7651u)), sa, bf);
    vn[1] = trav_fma(trav_u2f(trav_prmt(wa, 0x4B000000u, 0x7652u)), sa, bn);
    vf[1] = trav_fma(trav_u2f(trav_prmt(wa, 0x4B000000u, 0x7653u)), sa, bf);
    vn[2] = trav_fma(trav_u2f(trav_prmt(wb, 0x4B000000u, 0x7650u)), sa, bn);
    vf[2] = trav_fma(trav_u2f(trav_prmt(wb, 0x4B000000u, 0x7651u)), sa, bf);
    vn[3] = trav_fma(trav_u2f(trav_prmt(wb, 0x4B000000u, 0x7652u)), sa, bn);
    vf[3] = trav_fma(trav_u2f(trav_prmt(wb, 0x4B000000u, 0x7653u)), sa, bf);
}
#endif

// ploc.cu
int rtr_bvh_run_build(rtr_bvh* b);
int rtr_bvh_finish(rtr_bvh* b);  // waits for the last build's verdict (converged? iterations) if it is still pending
int rtr_bvh_export_clusters(rtr_bvh* b, rtr_node* clusters_dev, uint32_t* parent_dev, uint32_t* left_dev,
                            uint32_t* right_dev, uint8_t* is_leaf_dev);
int rtr_bvh_compute_trace_params(rtr_bvh* b);
// (re)derive the child-pair records of an adopted / received flat array (flat_view, wtri_by_rank must be set)
int rtr_bvh_pack_pairs_own(rtr_bvh* b);
// trace.cu
int rtr_trace_primary_launch(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera& cam, uint32_t width, uint32_t height,
                             uint32_t denom_w, uint32_t denom_h, uint32_t row0, uint32_t row1, uint32_t flags,
                             rtr_hit* hits_dev);
int rtr_trace_rays_launch(rtr_ctx* ctx, const rtr_bvh* b, const rtr_ray* rays_dev, uint64_t n_rays, int any_hit,
                          const float* t_max_dev, uint32_t flags, rtr_hit* hits_dev);
int rtr_render_launch(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera& cam, uint32_t width, uint32_t height,
                      uint32_t denom_w, uint32_t denom_h, uint32_t row0, uint32_t row1, uint32_t bounces, int shadow,
                      const float light[3], uint32_t flags, float* rgba_dev, rtr_hit* hits_dev, uint64_t* rays_dev,
                      uint32_t rows_per_block = 0, uint32_t total_stripes = 1, uint32_t nb_stripes = 1,
                      const uint8_t* stripe_offsets = nullptr);
int rtr_shade_launch(rtr_ctx* ctx, const rtr_hit* hits_dev, uint64_t n, const rtr_triangle* tris_dev, const rtr_mesh* meshes_dev,
                     const rtr_material* materials_dev, uint32_t flags, const float* bvh_rgba_dev, float* rgba_dev);
int rtr_depth_overlay_launch(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera& cam, uint32_t width, uint32_t height,
                             uint32_t denom_w, uint32_t denom_h, int display_depth, float* bvh_rgba_dev);
