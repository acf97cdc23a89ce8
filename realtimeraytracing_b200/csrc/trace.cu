// trace.cu -- stack-based BVH traversal + ray/triangle intersection over the reference's
// 48-byte DFS pre-order node array.
//
// Replaces the compute shader srcCommon/shaders/raytracer.glsl: getRay (:92-100), pixel mapping
// (:303-305), intersectBVH (:182-237), isLeafBVH (:239-244), rayTriangleIntersection (:102-147),
// getClosestHitBVH (:246-295) with the guarded-miss semantics of getAllHits (:149-157) (Q6).
// Every fp32 expression is evaluated in the order of oracle/rtr_oracle.c (library built with
// -fmad=false), so hit records are bit-identical to the CPU restatement.
//
// Two visiting orders produce the same records:
//   RTR_TRACE_REFERENCE_ORDER  exactly the shader's loop: pop, box test, leaf -> triangle test,
//                              else push Left then Right; no pruning.
//   default                    front-to-back: at an inner node both child boxes are tested (the same
//                              slab tests the shader performs when it pops them), the nearer child is
//                              entered first and the farther one is stacked with its entry distance;
//                              a node is skipped when that distance is beyond the closest hit so far
//                              by more than a conservative margin derived from the shader's own
//                              |a| >= 1e-4 acceptance test (DESIGN.md "pruning bound"), so skipped
//                              subtrees cannot contain a hit the shader would have preferred.  The
//                              shader keeps the FIRST of several equal-t hits it meets and meets leaves
//                              in descending flat index (right child first), so here an equal-t hit
//                              replaces the current one iff its leaf index is larger.
//
// Triangles are read from the world-space cache the build writes beside the leaves (wtri, 48 B per
// triangle in leaf = Morton order; the leaf's slot is kept in the node's third padding word), i.e. the
// M*P products the shader recomputes at every test (:105-107), evaluated once with the same fp32 ops.
#include "bvh.cuh"

namespace {

constexpr int kStack = 128;       // lane stack of the fast kernels
constexpr int kStackDeep = 1024;  // the shader's own: `const uint STACK_SIZE = 1024` (raytracer.glsl:251), RTR_TRACE_DEEP_STACK
constexpr int kTraceBlock = 128;

struct Ray {
    float ox, oy, oz;
    float dx, dy, dz;
    float ix, iy, iz;  // 1/d, the shader recomputes these per node (:187,:200,:213): same value
};

struct Hit {
    float b0, b1, b2, t;
    uint32_t did_hit, tri;
    uint32_t slot;  // wtri slot of the triangle (not part of the record)
};
__device__ __forceinline__ Hit no_hit() {
    Hit h;
    h.b0 = h.b1 = h.b2 = h.t = 0.f; h.did_hit = 0u; h.tri = 0u; h.slot = 0u;
    return h;
}

// where the traversal finds its data
struct Accel {
    const rtr_node* __restrict__ nodes;
    const float4* __restrict__ wtri;
    const uint4* __restrict__ pairs;  // traversal records (bvh.cuh: TravRec), default traversal only
    uint32_t by_rank;  // 1: wtri slot = leaf's third padding word (built here); 0: slot = triangle id (adopted nodes)
};

__device__ __forceinline__ Ray make_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
    r.ix = __fdiv_rn(1.0f, dx); r.iy = __fdiv_rn(1.0f, dy); r.iz = __fdiv_rn(1.0f, dz);
    return r;
}

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}
// GLSL cross(x, y) = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
__device__ __forceinline__ void cross3(float xx, float xy, float xz, float yx, float yy, float yz,
                                       float& ox, float& oy, float& oz) {
    ox = __fsub_rn(__fmul_rn(xy, yz), __fmul_rn(yy, xz));
    oy = __fsub_rn(__fmul_rn(xz, yx), __fmul_rn(yz, xx));
    oz = __fsub_rn(__fmul_rn(xx, yy), __fmul_rn(yx, xy));
}
// normalize(v) := v / sqrt(dot(v, v)) (pinned definition, see oracle/rtr_oracle.c normalize3)
__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {
    const float len = __fsqrt_rn(dot3(x, y, z, x, y, z));
    x = __fdiv_rn(x, len); y = __fdiv_rn(y, len); z = __fdiv_rn(z, len);
}

// raytracer.glsl:92-100 + :303-305
__device__ __forceinline__ Ray camera_ray(const rtr_camera& cam, uint32_t x, uint32_t y, uint32_t denom_w,
                                          uint32_t denom_h) {
    const float px = __fdiv_rn((float)x, (float)denom_w);
    const float py = __fdiv_rn((float)y, (float)denom_h);
    const float v0 = __fmul_rn(__fsub_rn(px, 0.5f), cam.plane_width);
    const float v1 = __fmul_rn(__fsub_rn(py, 0.5f), cam.plane_height);
    const float v2 = __fmul_rn(1.f, cam.plane_near);
    const float v3 = 1.f;
    const float* m = cam.inv_view;
    float pw[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        pw[r] = __fadd_rn(__fadd_rn(__fmul_rn(m[0 + r], v0), __fmul_rn(m[4 + r], v1)),
                          __fadd_rn(__fmul_rn(m[8 + r], v2), __fmul_rn(m[12 + r], v3)));
    const float d0 = __fsub_rn(pw[0], cam.eye[0]), d1 = __fsub_rn(pw[1], cam.eye[1]);
    const float d2 = __fsub_rn(pw[2], cam.eye[2]), d3 = __fsub_rn(pw[3], cam.eye[3]);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)),
                                           __fmul_rn(d3, d3)));
    return make_ray(cam.eye[0], cam.eye[1], cam.eye[2], __fdiv_rn(d0, len), __fdiv_rn(d1, len), __fdiv_rn(d2, len));
}

// world-space vertices in SHADER naming (Q8): shader _P1 = host _P2, shader _P2 = host _P1.
// wtri slot = 3 float4: (P0.xyz, P1.x) (P1.yz, P2.xy) (P2.z, -, -, -), host naming, world space.
struct TriWorld { float3 p0, p1, p2; };
__device__ __forceinline__ TriWorld shader_vertices(const float4* __restrict__ wtri, uint32_t slot) {
    const float4* p = wtri + (size_t)slot * 3;
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    TriWorld w;
    w.p0 = make_float3(a.x, a.y, a.z);
    w.p1 = make_float3(b.z, b.w, c.x);
    w.p2 = make_float3(a.w, b.x, b.y);
    return w;
}

// raytracer.glsl:102-147 on world-space vertices in shader naming
__device__ __forceinline__ bool ray_triangle_w(const Ray& r, const TriWorld& w, uint32_t slot, uint32_t tri, Hit& hit) {
    const float e0x = __fsub_rn(w.p1.x, w.p0.x), e0y = __fsub_rn(w.p1.y, w.p0.y), e0z = __fsub_rn(w.p1.z, w.p0.z);
    const float e1x = __fsub_rn(w.p2.x, w.p0.x), e1y = __fsub_rn(w.p2.y, w.p0.y), e1z = __fsub_rn(w.p2.z, w.p0.z);
    float nx, ny, nz;
    cross3(e1x, e1y, e1z, e0x, e0y, e0z, nx, ny, nz);
    float qx, qy, qz;
    cross3(r.dx, r.dy, r.dz, e1x, e1y, e1z, qx, qy, qz);
    const float a = dot3(e0x, e0y, e0z, qx, qy, qz);
    if (fabsf(a) < 1e-4f) return false;  // the shader's "facing away || |a| < 1e-4" has no side effects: order is free
    // Facing test dot(normalize(n), d) >= 0 (raytracer.glsl:112-119).  Only its sign matters, and with
    // D = n.d the normalised value differs from D/|n| by less than 4.3u|d| while the plain fp32
    // dot3(n, d) differs from D by less than 3.1u|n||d|: beyond |dot3(n, d)| > 2e-6 |n||d| both have D's
    // sign, and the square root and three divisions are skipped.  Closer to zero (or out of fp32's
    // normal range) the shader's own expression decides.
    {
        const float sd = dot3(nx, ny, nz, r.dx, r.dy, r.dz);
        const float n2 = dot3(nx, ny, nz, nx, ny, nz);
        const float d2 = dot3(r.dx, r.dy, r.dz, r.dx, r.dy, r.dz);
        const float bound = 4.1e-12f * n2 * d2;
        if (n2 > 1e-30f && n2 < 1e30f && d2 > 1e-30f && d2 < 1e30f && sd * sd > bound) {
            if (sd > 0.f) return false;
        } else {
            normalize3(nx, ny, nz);
            if (dot3(nx, ny, nz, r.dx, r.dy, r.dz) >= 0.f) return false;
        }
    }
    const float sx = __fdiv_rn(__fsub_rn(r.ox, w.p0.x), a);
    const float sy = __fdiv_rn(__fsub_rn(r.oy, w.p0.y), a);
    const float sz = __fdiv_rn(__fsub_rn(r.oz, w.p0.z), a);
    float rx, ry, rz;
    cross3(sx, sy, sz, e0x, e0y, e0z, rx, ry, rz);
    const float bx = dot3(sx, sy, sz, qx, qy, qz);
    const float by = dot3(rx, ry, rz, r.dx, r.dy, r.dz);
    const float bz = __fsub_rn(__fsub_rn(1.f, bx), by);
    if (bx < 0.f || by < 0.f || bz < 0.f) return false;
    const float t = dot3(e1x, e1y, e1z, rx, ry, rz);
    if (t < 0.f) return false;
    hit.b0 = bx; hit.b1 = by; hit.b2 = bz; hit.t = t; hit.did_hit = 1u; hit.tri = tri; hit.slot = slot;
    return true;
}
__device__ __forceinline__ bool ray_triangle(const Ray& r, const float4* __restrict__ wtri, uint32_t slot,
                                             uint32_t tri, Hit& hit) {
    return ray_triangle_w(r, shader_vertices(wtri, slot), slot, tri, hit);
}

// raytracer.glsl:182-237 (edge code 2 folded into "hit"); IEEE fminf/fmaxf (Q11)
__device__ __forceinline__ bool intersect_box(const Ray& r, const float4 lo, const float4 hi, float& t_entry) {
    const float tx1 = __fmul_rn(__fsub_rn(lo.x, r.ox), r.ix), tx2 = __fmul_rn(__fsub_rn(hi.x, r.ox), r.ix);
    float tMin = fminf(tx1, tx2), tMax = fmaxf(tx1, tx2);
    if (tMax < 0.f || tMin > tMax) return false;
    const float ty1 = __fmul_rn(__fsub_rn(lo.y, r.oy), r.iy), ty2 = __fmul_rn(__fsub_rn(hi.y, r.oy), r.iy);
    tMin = fmaxf(tMin, fminf(ty1, ty2));
    tMax = fminf(tMax, fmaxf(ty1, ty2));
    if (tMax < 0.f || tMin > tMax) return false;
    const float tz1 = __fmul_rn(__fsub_rn(lo.z, r.oz), r.iz), tz2 = __fmul_rn(__fsub_rn(hi.z, r.oz), r.iz);
    tMin = fmaxf(tMin, fminf(tz1, tz2));
    tMax = fminf(tMax, fmaxf(tz1, tz2));
    t_entry = tMin;
    return tMax >= 0.f && tMin <= tMax;
}

struct NodeRec {
    float4 lo;  // min.xyz, pad
    float4 hi;  // max.xyz, pad
    uint4 links;  // triangle id, left, right, wtri slot (third padding word)
};
__device__ __forceinline__ NodeRec load_node(const rtr_node* __restrict__ nodes, uint32_t idx) {
    const uint4* p = reinterpret_cast<const uint4*>(nodes + idx);
    const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    NodeRec n;
    n.lo = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), 0.f);
    n.hi = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), 0.f);
    n.links = c;
    return n;
}
// 256-bit read-only load; p must be 32-byte aligned
__device__ __forceinline__ void ld_nc_256(const uint4* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}
__device__ __forceinline__ uint4 load_links(const rtr_node* __restrict__ nodes, uint32_t idx) {
    return __ldg(reinterpret_cast<const uint4*>(nodes + idx) + 2);
}
__device__ __forceinline__ bool is_leaf(const uint4 links) { return links.y == 0u && links.z == 0u; }
__device__ __forceinline__ uint32_t slot_of(const Accel& A, const uint4 links) { return A.by_rank ? links.w : links.x; }

// Conservative pruning bound.  A triangle accepted by ray_triangle has |a| >= 1e-4, which bounds the
// rounding error of its computed t by 40*u*(|e0||e1|/1e-4)*(t_true + Emax); with
// K = 64*u*Emax^2/1e-4 no triangle inside a box whose computed entry distance exceeds
// (t_best + K*Emax)*(1 + 2K) can produce a computed t at or below t_best.  K >= 0.25 disables pruning.
struct PruneBound {
    float k, kemax;
    bool enabled;
};
__device__ __forceinline__ PruneBound make_prune_bound(const TraceParams* tp, bool want) {
    PruneBound pb;
    pb.enabled = false; pb.k = 0.f; pb.kemax = 0.f;
    if (!want) return pb;
    const float emax2 = ordered_to_float(tp->emax2_ordered);
    if (!(emax2 >= 0.f) || tp->emax2_ordered == 0u) { pb.enabled = true; return pb; }  // no triangles with extent
    const float k = 64.f * 5.9604645e-8f * emax2 * 1.0e4f * 1.01f;
    if (k < 0.25f) { pb.enabled = true; pb.k = k; pb.kemax = k * sqrtf(emax2) * 1.01f; }
    return pb;
}
__device__ __forceinline__ float prune_limit(const PruneBound& pb, float t_best) {
    return (t_best + pb.kemax) * (1.f + 2.f * pb.k);
}

// ---- the shader's own loop (raytracer.glsl:246-295): RTR_TRACE_REFERENCE_ORDER ----
template <int STACK = kStack>
__device__ __forceinline__ Hit closest_hit_reference(const Ray& r, const Accel& A, uint32_t* overflow) {
    Hit best = no_hit();
    uint32_t stack[STACK];
    int sp = 0;
    stack[sp++] = 0u;
    while (sp > 0) {
        const uint32_t idx = stack[--sp];
        const NodeRec n = load_node(A.nodes, idx);
        float t_entry;
        if (!intersect_box(r, n.lo, n.hi, t_entry)) continue;
        if (is_leaf(n.links)) {
            Hit h;
            if (ray_triangle(r, A.wtri, slot_of(A, n.links), n.links.x, h)) {
                if (best.did_hit == 0u || h.t < best.t) best = h;
            }
        } else if (sp + 2 <= STACK) {
            stack[sp++] = n.links.y;
            stack[sp++] = n.links.z;
        } else {
            *overflow = 1u;
        }
    }
    return best;
}
template <int STACK = kStack>
__device__ __forceinline__ bool any_hit_reference(const Ray& r, float t_max, const Accel& A, uint32_t* overflow) {
    uint32_t stack[STACK];
    int sp = 0;
    stack[sp++] = 0u;
    while (sp > 0) {
        const uint32_t idx = stack[--sp];
        const NodeRec n = load_node(A.nodes, idx);
        float t_entry;
        if (!intersect_box(r, n.lo, n.hi, t_entry)) continue;
        if (is_leaf(n.links)) {
            Hit h;
            if (ray_triangle(r, A.wtri, slot_of(A, n.links), n.links.x, h) && h.t < t_max) return true;
        } else if (sp + 2 <= STACK) {
            stack[sp++] = n.links.y;
            stack[sp++] = n.links.z;
        } else {
            *overflow = 1u;  // a subtree is dropped: "no occluder" would not be trustworthy
        }
    }
    return false;
}

// ---- front-to-back traversal with conservative pruning; same records as the loop above ----
// ANY: return as soon as a triangle passes with t < t_max (did_hit only).
template <bool ANY>
__device__ __forceinline__ Hit trace_ordered(const Ray& r, float t_max, const Accel& A, const PruneBound& pb,
                                             uint32_t* overflow) {
    Hit best = no_hit();
    uint32_t best_node = 0u;
    float limit = (ANY && pb.enabled) ? prune_limit(pb, t_max) : INFINITY;
    uint32_t cur = 0u;
    uint4 links;
    {
        const NodeRec root = load_node(A.nodes, 0u);
        float te;
        if (!intersect_box(r, root.lo, root.hi, te) || te > limit) return best;
        links = root.links;
    }
    uint2 stack[kStack];  // (node, bits of its entry distance); both boxes of a pair were already tested
    int sp = 0;
    while (true) {
        if (is_leaf(links)) {
            Hit h;
            if (ray_triangle(r, A.wtri, slot_of(A, links), links.x, h)) {
                if (ANY) {
                    if (h.t < t_max) return h;
                } else if (best.did_hit == 0u || h.t < best.t || (h.t == best.t && cur > best_node)) {
                    best = h; best_node = cur;
                    if (pb.enabled) limit = prune_limit(pb, h.t);
                }
            }
        } else {
            const uint32_t li = links.y, ri = links.z;
            const NodeRec L = load_node(A.nodes, li), R = load_node(A.nodes, ri);
            float tl, tr;
            const bool hl = intersect_box(r, L.lo, L.hi, tl) && !(tl > limit);
            const bool hr = intersect_box(r, R.lo, R.hi, tr) && !(tr > limit);
            if (hl && hr) {
                const bool left_first = tl <= tr;
                if (sp < kStack) stack[sp++] = left_first ? make_uint2(ri, __float_as_uint(tr)) : make_uint2(li, __float_as_uint(tl));
                else *overflow = 1u;
                cur = left_first ? li : ri;
                links = left_first ? L.links : R.links;
                continue;
            }
            if (hl) { cur = li; links = L.links; continue; }
            if (hr) { cur = ri; links = R.links; continue; }
        }
        bool found = false;
        while (sp > 0) {
            const uint2 e = stack[--sp];
            if (__uint_as_float(e.y) > limit) continue;
            cur = e.x;
            links = load_links(A.nodes, cur);
            found = true;
            break;
        }
        if (!found) break;
    }
    return best;
}

// MODE 0: the shader's loop with the 128-entry stack; 1: front-to-back with pruning (one thread per ray; the default
// order itself runs in trace_persistent_kernel); 2: the shader's loop with the shader's own 1024-entry stack
template <int MODE>
__device__ __forceinline__ Hit closest_hit(const Ray& r, const Accel& A, const PruneBound& pb, uint32_t* overflow) {
    if (MODE == 1) return trace_ordered<false>(r, INFINITY, A, pb, overflow);
    if (MODE == 2) return closest_hit_reference<kStackDeep>(r, A, overflow);
    return closest_hit_reference<kStack>(r, A, overflow);
}
template <int MODE>
__device__ __forceinline__ bool any_hit(const Ray& r, float t_max, const Accel& A, const PruneBound& pb, uint32_t* overflow) {
    if (MODE == 1) return trace_ordered<true>(r, t_max, A, pb, overflow).did_hit != 0u;
    if (MODE == 2) return any_hit_reference<kStackDeep>(r, t_max, A, overflow);
    return any_hit_reference<kStack>(r, t_max, A, overflow);
}

__device__ __forceinline__ void store_hit(rtr_hit* __restrict__ out, size_t i, const Hit& h) {
    // 24-byte records: three 8-byte stores
    uint2* p = reinterpret_cast<uint2*>(out + i);
    p[0] = make_uint2(__float_as_uint(h.b0), __float_as_uint(h.b1));
    p[1] = make_uint2(__float_as_uint(h.b2), __float_as_uint(h.t));
    p[2] = make_uint2(h.did_hit, h.tri);
}

// hit frame shared by both secondary rays: unit front normal and offset origin h + n*1e-3
__device__ __forceinline__ void hit_frame_w(const Ray& r, float t, const TriWorld& w, float& nx, float& ny, float& nz,
                                            float& ox, float& oy, float& oz) {
    const float e0x = __fsub_rn(w.p1.x, w.p0.x), e0y = __fsub_rn(w.p1.y, w.p0.y), e0z = __fsub_rn(w.p1.z, w.p0.z);
    const float e1x = __fsub_rn(w.p2.x, w.p0.x), e1y = __fsub_rn(w.p2.y, w.p0.y), e1z = __fsub_rn(w.p2.z, w.p0.z);
    cross3(e1x, e1y, e1z, e0x, e0y, e0z, nx, ny, nz);
    normalize3(nx, ny, nz);
    ox = __fadd_rn(__fadd_rn(r.ox, __fmul_rn(r.dx, t)), __fmul_rn(nx, 1e-3f));
    oy = __fadd_rn(__fadd_rn(r.oy, __fmul_rn(r.dy, t)), __fmul_rn(ny, 1e-3f));
    oz = __fadd_rn(__fadd_rn(r.oz, __fmul_rn(r.dz, t)), __fmul_rn(nz, 1e-3f));
}
__device__ __forceinline__ void hit_frame(const Ray& r, const Hit& h, const Accel& A, float& nx, float& ny, float& nz,
                                          float& ox, float& oy, float& oz) {
    hit_frame_w(r, h.t, shader_vertices(A.wtri, h.slot), nx, ny, nz, ox, oy, oz);
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// Which image rows a launch covers: a contiguous range [row0, row0+rows) written at local row
// indices, or (count > 1) the row blocks dealt to one rank -- block b of rows_per_block rows belongs to
// stripe b % count, the rank owns the `span` stripes off[0..span) -- written at their global rows of the
// full image.  (span == 1, off[0] = rank: block b -> rank b % count.)
struct RowMap {
    uint32_t row0, rows, height;
    uint32_t rpb, count, span;
    uint8_t off[RTR_MAX_STRIPES_PER_RANK];
};
__device__ __forceinline__ bool map_row(const RowMap& m, uint32_t y_local, uint32_t& y, uint32_t& out_row) {
    if (m.count <= 1) { y = m.row0 + y_local; out_row = y_local; return true; }
    const uint32_t blk = y_local / m.rpb;  // local block: cycle blk / span, stripe off[blk % span]
    y = ((blk / m.span) * m.count + m.off[blk % m.span]) * m.rpb + (y_local % m.rpb);
    out_row = y;
    return y < m.height;
}

// Pixels are mapped to threads in 8x4 blocks per warp (2-D locality => coherent node fetches).
__device__ __forceinline__ bool pixel_of_thread(uint32_t width, uint32_t rows, uint32_t& x, uint32_t& y_local) {
    const uint32_t tiles_x = (width + 7) / 8;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tx = warp_global % tiles_x, ty = warp_global / tiles_x;
    x = tx * 8 + (lane & 7u);
    y_local = ty * 4 + (lane >> 3);
    return x < width && y_local < rows;
}

template <int PRUNE>
__global__ void __launch_bounds__(kTraceBlock)
trace_primary_kernel(const Accel A, TraceParams* __restrict__ tp, const rtr_camera cam,
                     uint32_t width, uint32_t denom_w, uint32_t denom_h, uint32_t row0, uint32_t rows,
                     rtr_hit* __restrict__ hits) {
    uint32_t x, yl;
    if (!pixel_of_thread(width, rows, x, yl)) return;
    const uint32_t y = row0 + yl;
    Hit h = no_hit();
    if (x < denom_w && y < denom_h) {
        const PruneBound pb = make_prune_bound(tp, PRUNE == 1);
        const Ray r = camera_ray(cam, x, y, denom_w, denom_h);
        uint32_t ovf = 0;
        h = closest_hit<PRUNE>(r, A, pb, &ovf);
        if (ovf) atomicAdd(&tp->stack_overflows, 1u);
    }
    store_hit(hits, (size_t)yl * width + x, h);
}

template <int PRUNE>
__global__ void __launch_bounds__(kTraceBlock)
trace_rays_kernel(const Accel A, TraceParams* __restrict__ tp,
                  const rtr_ray* __restrict__ rays, uint64_t n_rays, int want_any, const float* __restrict__ t_max,
                  rtr_hit* __restrict__ hits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const float4 o = __ldg(reinterpret_cast<const float4*>(rays) + 2 * i);
    const float4 d = __ldg(reinterpret_cast<const float4*>(rays) + 2 * i + 1);
    const Ray r = make_ray(o.x, o.y, o.z, d.x, d.y, d.z);
    const PruneBound pb = make_prune_bound(tp, PRUNE == 1);
    Hit h = no_hit();
    uint32_t ovf = 0;
    if (want_any) {
        const float tm = t_max ? t_max[i] : INFINITY;
        h.did_hit = any_hit<PRUNE>(r, tm, A, pb, &ovf) ? 1u : 0u;
    } else {
        h = closest_hit<PRUNE>(r, A, pb, &ovf);
    }
    if (ovf) atomicAdd(&tp->stack_overflows, 1u);
    store_hit(hits, i, h);
}

// Multi-bounce frame, one thread per pixel (definition: oracle/rtr_oracle.c orc_render).
template <int PRUNE>
__global__ void __launch_bounds__(kTraceBlock)
render_kernel(const Accel A, TraceParams* __restrict__ tp, const rtr_camera cam,
              uint32_t width, uint32_t denom_w, uint32_t denom_h, const RowMap rm,
              uint32_t bounces, int shadow, float lx, float ly, float lz,
              float4* __restrict__ rgba, rtr_hit* __restrict__ hits, unsigned long long* __restrict__ rays_traced) {
    uint32_t x, yl, y = 0, out_row = 0;
    const bool live = pixel_of_thread(width, rm.rows, x, yl) && map_row(rm, yl, y, out_row);
    uint32_t traced = 0;
    if (live) {
        float L = 0.f, wgt = 1.f;
        Hit first = no_hit();
        if (x < denom_w && y < denom_h) {
            const PruneBound pb = make_prune_bound(tp, PRUNE == 1);
            Ray r = camera_ray(cam, x, y, denom_w, denom_h);
            uint32_t ovf = 0;
            for (uint32_t k = 0; k <= bounces; ++k) {
                const Hit h = closest_hit<PRUNE>(r, A, pb, &ovf);
                ++traced;
                if (k == 0) first = h;
                if (!h.did_hit) break;
                float nx, ny, nz, ox, oy, oz, c;
                hit_frame(r, h, A, nx, ny, nz, ox, oy, oz);
                if (shadow) {
                    const float sx = __fsub_rn(lx, ox), sy = __fsub_rn(ly, oy), sz = __fsub_rn(lz, oz);
                    const float len = __fsqrt_rn(dot3(sx, sy, sz, sx, sy, sz));
                    const Ray sr = make_ray(ox, oy, oz, __fdiv_rn(sx, len), __fdiv_rn(sy, len), __fdiv_rn(sz, len));
                    const bool occ = any_hit<PRUNE>(sr, len, A, pb, &ovf);
                    ++traced;
                    const float ndl = dot3(nx, ny, nz, sr.dx, sr.dy, sr.dz);
                    c = occ ? 0.f : fmaxf(0.f, ndl);
                } else {
                    c = -dot3(nx, ny, nz, r.dx, r.dy, r.dz);
                }
                L = __fadd_rn(L, __fmul_rn(wgt, c));
                wgt = __fmul_rn(wgt, 0.5f);
                if (k < bounces) {
                    const float kk = __fmul_rn(2.f, dot3(r.dx, r.dy, r.dz, nx, ny, nz));
                    float rx = __fsub_rn(r.dx, __fmul_rn(nx, kk));
                    float ry = __fsub_rn(r.dy, __fmul_rn(ny, kk));
                    float rz = __fsub_rn(r.dz, __fmul_rn(nz, kk));
                    normalize3(rx, ry, rz);
                    r = make_ray(ox, oy, oz, rx, ry, rz);
                }
            }
            if (ovf) atomicAdd(&tp->stack_overflows, 1u);
        }
        const size_t o = (size_t)out_row * width + x;
        if (rgba) rgba[o] = make_float4(L, L, L, 1.f);
        if (hits) store_hit(hits, o, first);
    }
    if (rays_traced) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) traced += __shfl_xor_sync(0xffffffffu, traced, o);
        if ((threadIdx.x & 31u) == 0 && traced) atomicAdd(rays_traced, (unsigned long long)traced);
    }
}

inline uint32_t pixel_grid(uint32_t width, uint32_t rows) {
    const uint64_t warps = (uint64_t)((width + 7) / 8) * ((rows + 3) / 4);
    return (uint32_t)((warps * 32 + kTraceBlock - 1) / kTraceBlock);
}

// ---------------------------------------------------------------------------------------
// Persistent traversal (default order).  Rays of a soup diverge badly under one-thread-one-pixel:
// lanes wait for the longest path of their warp and almost every loop iteration executes both the
// box branch and the triangle branch for a handful of lanes (ncu: 6.5 of 32 threads active per
// instruction, profiles/r01_ncu_v2_render.txt).  Here
//   * a launch is one CTA per resident slot; warps pull 32-job tiles (an 8x4 pixel block or 32 rays)
//     from a global counter and hand them to lanes as they fall idle, a lane keeps its pixel through
//     all bounces;
//   * a lane that reaches a leaf parks it (one pending leaf) and keeps walking; the triangle tests of
//     a warp run together once a lane is blocked on a second leaf or enough leaves are parked;
//   * finished rays are shaded / replaced in batches.
// The per-ray arithmetic and the set of leaves a ray meets before its stack runs dry are those of
// trace_ordered; only the interleaving between lanes changes, so records stay bit-identical.
// ---------------------------------------------------------------------------------------
#ifndef RTR_LEAF_BATCH
#define RTR_LEAF_BATCH 10
#endif
#ifndef RTR_FIN_BATCH
#define RTR_FIN_BATCH 4
#endif
#ifndef RTR_TRACE_MIN_CTAS
#define RTR_TRACE_MIN_CTAS 8
#endif
constexpr uint32_t kLeafBatch = RTR_LEAF_BATCH;  // run the triangle step once this many lanes hold a parked leaf
constexpr uint32_t kFinBatch = RTR_FIN_BATCH;    // shade/replace finished rays once this many lanes wait
#ifndef RTR_PREFETCH
#define RTR_PREFETCH 2   // bit 0: record of a stacked child (pair step), bit 1: record of a parked leaf
#endif
constexpr int kPrefetch = RTR_PREFETCH;
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// three-input min / max (FMNMX3, sm_100+); NaN operands are ignored like in fminf / fmaxf
__device__ __forceinline__ float max3f(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min3f(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
// Wide step: test the four grandchild slots of the second record half (bvh.cuh) instead of the child pair -- half the
// dependent fetches per ray.
#ifndef RTR_WIDE
#define RTR_WIDE 1
#endif
// A leaf met while the parking slot is free is parked in the same step (and the walk pops on) instead of costing a
// walk step of its own.
#ifndef RTR_EAGER_PARK
#define RTR_EAGER_PARK 1
#endif
constexpr bool kEagerPark = RTR_EAGER_PARK != 0;
constexpr bool kWide = RTR_WIDE != 0;
#ifndef RTR_WALK_STEPS
#define RTR_WALK_STEPS 4
#endif
constexpr int kWalkSteps = RTR_WALK_STEPS;
#ifndef RTR_BLOCK_BATCH
#define RTR_BLOCK_BATCH 4
#endif
constexpr uint32_t kBlockBatch = RTR_BLOCK_BATCH;  // ... or once this many lanes wait on a second leaf
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kDry = 0xFFFFFFFFu;  // stack ran dry (also the state of an idle lane)
// Drain: once the job pool is empty, idle lanes of a warp take stack entries off the lanes still walking (below).
#ifndef RTR_STEAL
#define RTR_STEAL 1
#endif
constexpr bool kSteal = RTR_STEAL != 0;
#ifdef RTR_STEAL_STATS   // diagnostics build: lane occupancy of the walk before / after the pool ran empty (profiles/steal_stats.py)
__device__ unsigned long long g_steal_stats[8];
#define RTR_STAT(i, v) do { if (lane == 0u) atomicAdd(&g_steal_stats[i], (unsigned long long)(v)); } while (0)
#else
#define RTR_STAT(i, v) do { } while (0)
#endif
constexpr uint32_t kHelperBit = 0x80000u;      // st bit 19: this lane walks part of another lane's ray
constexpr uint32_t kMasterShift = 20u;         // st bits 20-24: the lane that owns the ray
#ifndef RTR_STEAL_MIN_IDLE
#define RTR_STEAL_MIN_IDLE 4
#endif
constexpr uint32_t kStealMinIdle = RTR_STEAL_MIN_IDLE;   // stacks are split once this many lanes of the warp are idle
// best hit of a ray as one word whose maximum is the closest hit: smaller t first (t >= 0: its bits order like the
// value), then the larger leaf index -- the rule of the leaf step.  0: no hit.
__device__ __forceinline__ unsigned long long pack_best(float t, uint32_t node) {
    return ((unsigned long long)(~__float_as_uint(t)) << 32) | node;
}
#ifndef RTR_SMEM_STACK
#define RTR_SMEM_STACK 20
#endif
constexpr int kSmemStack = RTR_SMEM_STACK;

struct JobDesc {
    uint32_t kind;   // 0: pixels (render / primary), 1: explicit rays
    uint32_t total;  // jobs: 32 per 8x4 pixel block, or rays
    // pixels
    rtr_camera cam;
    uint32_t width, denom_w, denom_h;
    RowMap rm;
    uint32_t bounces;
    int shadow;
    float lx, ly, lz;
    float4* rgba;
    rtr_hit* hits;
    // rays
    const rtr_ray* rays;
    const float* t_max;
    int want_any;
};

// Lane state is kept small (occupancy hides the node-fetch latency): the node being expanded is one word --
// an inner flat index, or kLeafBit | index for a leaf -- stack entries carry that word plus the entry
// distance, and the best hit is (t, leaf index) only: its barycentrics and triangle id are regenerated by one
// more ray_triangle when the ray finishes.
//
// Inner nodes are expanded from the compressed record (bvh.cuh) -- 256-bit loads where two 48-byte nodes took six
// 128-bit loads; the kernel was bound by L1 wavefronts of exactly these divergent loads, then by the number of
// dependent fetches per ray (hence the wide step below).  The compressed child boxes are supersets, evaluated here
// with an error margin that also covers the rounding of the reference's own slab test, so that
//     reference slab test passes on the exact box  ==>  compressed test passes            (*)
// and the entry distance computed here is a lower bound of the reference's.  With the node origin c, grid
// step s = 2^e, plane byte q and the ray (o, inv = 1/d), per axis
//     a = s*inv (exact), b = fl(fl(c - o)*inv), x = 2^23 + q (exact, built by PRMT),
//     t = fl(x*a + fl(fl(b -+ m) - 2^23*a))
// differs from the real (c + q*s - o)*inv by at most |a|/2 + 5u(|b| + 256|a|), the reference's
// fl(fl(X - o)*inv) from its real value by at most 2.01u(|b| + 256|a|) (u = 2^-24): the margin
// m = 0.52|a| + 2e-6|b| + 1e-30 covers both with room.  A ray with an |inv| not below 2^64 (d = 0, where
// the reference's inf/NaN arithmetic decides, or so small that a product could overflow) takes the
// reference's slab test on the decoded superset boxes instead (monotone in the box, hence (*) again).  A
// leaf is then tested exactly as in the reference: slab test on its own exact box, ray/triangle on its vertices, both
// from its 64-byte record.
//
// Wide step (RTR_WIDE): the second half of an inner record holds the boxes of the node's four GRANDCHILD slots on
// the same grid (same origin, same steps, bytes 0..255), so the derivation above applies to them word for word; a
// slot's box is a superset of the grandchild's exact box, which contains every leaf box below it, hence (*) holds
// for every leaf under a slot that is skipped.  One step tests the four slots, orders the hit ones by entry distance
// and stacks up to three: half the dependent fetches per ray.  The step clamps the entry distance at 0 and the exit
// distance at the pruning limit before comparing them -- "exit >= 0 and entry <= exit and entry <= limit" in one
// comparison; the limit is >= 0 whenever a hit is still possible (it bounds a hit distance, and hits have t > 0),
// and a negative limit (an any-hit ray with t_max < 0) can only reject.  Rays on the exact-slab path and nodes whose
// second half could not be encoded keep to pair steps; the two kinds of step can alternate freely along a path,
// every inner record serves both.
__global__ void __launch_bounds__(kTraceBlock, RTR_TRACE_MIN_CTAS)
trace_persistent_kernel(const Accel A, TraceParams* __restrict__ tp, const JobDesc jd,
                        unsigned long long* __restrict__ rays_traced, uint32_t* __restrict__ sm_table, uint32_t reserve) {
    // rtr_ctx_reserve_sms: the launch covers every resident slot of the device and the CTAs that land on the first
    // `reserve` SMs to show up leave at once, so that WHOLE SMs stay empty for the kernels of a concurrent
    // collective (a smaller grid would still spread over all SMs and leave no room for a 512-thread NCCL CTA).
    // The first CTA of an SM decides for it; sm_table[0] counts the SMs given away, sm_table[1 + smid] is
    // 0 unseen / 1 deciding / 2 given away / 3 kept.
    if (reserve != 0u) {
        __shared__ uint32_t s_leave;
        if (threadIdx.x == 0) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            uint32_t* slot = sm_table + 1u + (smid & 1023u);
            uint32_t state = atomicCAS(slot, 0u, 1u);
            if (state == 0u) {
                state = atomicAdd(sm_table, 1u) < reserve ? 2u : 3u;
                atomicExch(slot, state);
            } else {
                while (state == 1u) state = ld_relaxed_u32(slot);
            }
            s_leave = state == 2u ? 1u : 0u;
        }
        __syncthreads();
        if (s_leave) return;
    }
    // Traversal stack: the first kSmemStack entries of every thread live in shared memory, laid out
    // [word][depth][thread] so that any mix of depths across a warp is bank-conflict free (bank = lane);
    // a divergent local-memory access would cost one L1 wavefront per lane instead.
    __shared__ uint32_t s_stack[2][kSmemStack][kTraceBlock];
    __shared__ float s_pb[2];                   // pruning bound constants k, k*Emax (negative k: disabled)
    // Six words per lane with two uses that never overlap in time: while the lane's SHADOW ray is out, the normal and the
    // incoming direction of the bounce (words 0-5); while its CLOSEST-HIT ray is shared with helper lanes (drain, below),
    // the best hit they have published (words 0-1 as one 64-bit word) and the number of helpers still out (word 2).
    // Any-hit rays are never shared and a ray only ends once its helpers are back, so the second use finds its
    // words 0-2 zero: set here and again when the stash is read.  (A separate array would push the CTA's shared memory
    // past the 196 KB carve-out for 8 CTAs per SM and shrink the L1: measured, the frame takes 27.8 ms instead of 24.2.)
    __shared__ __align__(8) uint32_t s_lane[kTraceBlock][6];
    auto lane_best = [&](uint32_t t) -> unsigned long long* { return reinterpret_cast<unsigned long long*>(&s_lane[t][0]); };
    s_lane[threadIdx.x][0] = 0u; s_lane[threadIdx.x][1] = 0u; s_lane[threadIdx.x][2] = 0u;
    if (threadIdx.x == 0) {
        const PruneBound pb = make_prune_bound(tp, true);
        s_pb[0] = pb.enabled ? pb.k : -1.f;
        s_pb[1] = pb.kemax;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;

    // warp-uniform
    uint32_t pool_next = 0u, pool_end = 0u, traced = 0u;
    bool exhausted = false;

    // lane state
    uint32_t job = RTR_NONE;
    Ray r = make_ray(0.f, 0.f, 0.f, 1.f, 1.f, 1.f);
    float best_t = INFINITY;        // closest hit so far; for an any-hit ray: its t_max
    uint32_t best_node = RTR_NONE;  // leaf of the best hit
    float limit = INFINITY;
    uint32_t a = kDry;              // node being expanded: inner flat index, or kLeafBit | index
    uint32_t pend_node = RTR_NONE;  // parked leaf
    int sp = 0;
    uint32_t st = 0u;               // bits 0-15 bounce index, bit 16 any-hit ray, bit 17 stack overflow seen,
                                    // bit 18 ray with a (nearly) zero direction component, bits 19-24 drain (kHelperBit ...)
    float L = 0.f;
    float stack_t[kStack - kSmemStack];   // entries the shared-memory part has no room for
    uint32_t stack_a[kStack - kSmemStack];

    auto limit_of = [&](float t) -> float {
        const float k = s_pb[0];
        return (k < 0.f) ? INFINITY : (t + s_pb[1]) * (1.f + 2.f * k);
    };

    // begin walking ray r: root test (raytracer.glsl:255-262 pops node 0 first)
    auto start_ray = [&](bool any, float tm) {
        st = (st & 0xFFFFu) | (any ? 0x10000u : 0u) | (st & 0x20000u);  // clears bit 18
        best_t = tm; best_node = RTR_NONE; sp = 0; pend_node = RTR_NONE;
        limit = any ? limit_of(tm) : INFINITY;
        // the compressed test needs every |1/d| below 2^64 (false for inf and NaN too)
        if (!(fabsf(r.ix) < 1.8e19f && fabsf(r.iy) < 1.8e19f && fabsf(r.iz) < 1.8e19f)) st |= 0x40000u;
        const NodeRec root = load_node(A.nodes, 0u);
        float te;
        a = kDry;
        if (intersect_box(r, root.lo, root.hi, te) && !(te > limit)) a = is_leaf(root.links) ? kLeafBit : 0u;
    };
    // Stack: entries below kSmemStack in shared memory, deeper ones (rare) in local memory.
    auto push_far = [&](float tf, uint32_t fa) {
        if (sp < kSmemStack) {
            s_stack[0][sp][threadIdx.x] = __float_as_uint(tf); s_stack[1][sp][threadIdx.x] = fa;
            ++sp;
        } else if (sp < kStack) {
            stack_t[sp - kSmemStack] = tf; stack_a[sp - kSmemStack] = fa;
            ++sp;
        } else st |= 0x20000u;
    };
    auto pop_next = [&]() {
        a = kDry;
        while (sp > 0) {
            --sp;
            float te; uint32_t w;
            if (sp < kSmemStack) { te = __uint_as_float(s_stack[0][sp][threadIdx.x]); w = s_stack[1][sp][threadIdx.x]; }
            else { te = stack_t[sp - kSmemStack]; w = stack_a[sp - kSmemStack]; }
            if (te > limit) continue;
            if (kEagerPark && (w & kLeafBit) != 0u && pend_node == RTR_NONE) {  // a leaf and the parking slot is free: park it, pop on
                pend_node = w & ~kLeafBit;
                if (kPrefetch & 2) prefetch_l2(A.pairs + (size_t)pend_node * 4);
                continue;
            }
            a = w;
            break;
        }
    };
    // one axis of the compressed tests (bvh.cuh: the same functions the host-side contract check runs)
    auto axis_planes = [&](float c, uint32_t ebyte, uint32_t quad, float o, float j,
                           float& tnl, float& tfl, float& tnr, float& tfr) {
        trav_axis_planes2(c, ebyte, quad, o, j, tnl, tfl, tnr, tfr);
    };
    auto axis_vals4 = [&](float c, uint32_t ebyte, uint32_t qa, uint32_t qb, float o, float j, float (&vn)[4], float (&vf)[4]) {
        trav_axis_vals4(c, ebyte, qa, qb, o, j, vn, vf);
    };
    // decoded planes of one axis, rounded outwards: min planes down, max planes up
    auto decode_axis = [&](float c, uint32_t ebyte, uint32_t quad, float& llo, float& lhi, float& rlo, float& rhi) {
        const float step = __uint_as_float(ebyte << 23);
        llo = __fadd_rd(c, __fmul_rn(__uint2float_rn(quad & 0xFFu), step));
        lhi = __fadd_ru(c, __fmul_rn(__uint2float_rn((quad >> 8) & 0xFFu), step));
        rlo = __fadd_rd(c, __fmul_rn(__uint2float_rn((quad >> 16) & 0xFFu), step));
        rhi = __fadd_ru(c, __fmul_rn(__uint2float_rn(quad >> 24), step));
    };
    auto job_pixel = [&](uint32_t j, uint32_t& x, uint32_t& y, uint32_t& out_row) -> bool {
        const uint32_t tiles_x = (jd.width + 7u) / 8u;
        const uint32_t wg = j >> 5, ln = j & 31u;
        // (dealing the tiles in compact super-blocks of 16..128 pixels instead of row-major order was measured: +1..2 %
        //  slower at 4K -- the rows of tiles in flight already share the upper tree, profiles/exp_r02_1.sh)
        x = (wg % tiles_x) * 8u + (ln & 7u);
        const uint32_t yl = (wg / tiles_x) * 4u + (ln >> 3);
        return x < jd.width && yl < jd.rm.rows && map_row(jd.rm, yl, y, out_row);
    };

    while (true) {
        // ---- hand jobs to idle lanes ----
        uint32_t idle = __ballot_sync(0xffffffffu, job == RTR_NONE);
        while (idle != 0u && !exhausted) {
            if (pool_next == pool_end) {
                uint32_t base = 0u;
                if (lane == 0u) base = atomicAdd(&tp->job_counter, 32u);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= jd.total) { exhausted = true; break; }
                pool_next = base;
                pool_end = min(base + 32u, jd.total);
            }
            const uint32_t avail = pool_end - pool_next;
            const uint32_t rank = __popc(idle & lanemask_lt());
            if (job == RTR_NONE && rank < avail) {
                job = pool_next + rank;
                st = 0u; L = 0.f;
                if (jd.kind == 0u) {
                    uint32_t x, y, out_row;
                    if (!job_pixel(job, x, y, out_row)) {
                        job = RTR_NONE;  // padding of the 8x4 block
                    } else if (x < jd.denom_w && y < jd.denom_h) {
                        r = camera_ray(jd.cam, x, y, jd.denom_w, jd.denom_h);
                        start_ray(false, INFINITY);
                    } else {  // not traced (Q5): zero record, black pixel
                        const size_t o = (size_t)out_row * jd.width + x;
                        if (jd.rgba) jd.rgba[o] = make_float4(0.f, 0.f, 0.f, 1.f);
                        if (jd.hits) store_hit(jd.hits, o, no_hit());
                        job = RTR_NONE;
                    }
                } else {
                    const float4 o = __ldg(reinterpret_cast<const float4*>(jd.rays) + 2 * (size_t)job);
                    const float4 d = __ldg(reinterpret_cast<const float4*>(jd.rays) + 2 * (size_t)job + 1);
                    r = make_ray(o.x, o.y, o.z, d.x, d.y, d.z);
                    const bool any = jd.want_any != 0;
                    start_ray(any, (any && jd.t_max) ? jd.t_max[job] : INFINITY);
                }
            }
            pool_next += min(avail, (uint32_t)__popc(idle));
            idle = __ballot_sync(0xffffffffu, job == RTR_NONE);
        }
        // ---- drain: the pool is empty and lanes fall idle while the last rays are still being walked -- the end of every
        //      launch, about 2 ms of a 4-bounce frame however few pixels it has (a 10 M-triangle soup: hundreds of
        //      dependent steps per ray, up to five rays per path).  Idle lanes pair up with walking lanes and take EVERY
        //      OTHER entry of their stacks (read straight from the donor's column of the shared-memory stack), together
        //      with a copy of the ray and of its best hit so far, and walk them as helpers; all pairs of a warp split
        //      at once.  The lanes of a ray publish their best hit in the owner's shared-memory word -- the closest hit is the
        //      maximum of a total order (smaller t, then larger leaf index: the rule of the leaf step) -- prune with
        //      it, and the owner takes it when its own stack is dry and no helper is out.  The leaves met are those of
        //      the one-lane walk or a superset, so the record is the same bits.  Any-hit rays stop at their first hit
        //      and are not shared. ----
        if (kSteal && exhausted && idle != 0xffffffffu && (uint32_t)__popc(idle) >= kStealMinIdle) {
            const bool can_give = job != RTR_NONE && (st & 0x10000u) == 0u && sp >= 2 && sp <= kSmemStack;
            const uint32_t m_give = __ballot_sync(0xffffffffu, can_give);
            const uint32_t pairs = min((uint32_t)__popc(m_give), (uint32_t)__popc(idle));
            RTR_STAT(0, 1); RTR_STAT(1, __popc(idle)); RTR_STAT(2, __popc(m_give)); RTR_STAT(3, pairs);
            if (pairs != 0u) {
                const uint32_t below = lanemask_lt();
                const bool thief = job == RTR_NONE && (uint32_t)__popc(idle & below) < pairs;
                const bool donor = can_give && (uint32_t)__popc(m_give & below) < pairs;
                uint32_t partner = lane;   // i-th idle lane <-> i-th donor
                if (thief) partner = __fns(m_give, 0u, __popc(idle & below) + 1);
                // the thief's lane state is replaced value by value (no block of temporaries: the walk's registers are
                // all live here); every lane takes part in the shuffles, the others read themselves
                {
                    const uint32_t g_st = __shfl_sync(0xffffffffu, st, partner);
                    const uint32_t g_job = __shfl_sync(0xffffffffu, job, partner);
                    if (thief) {
                        const uint32_t owner = (g_st & kHelperBit) ? ((g_st >> kMasterShift) & 31u) : partner;
                        atomicAdd(&s_lane[(threadIdx.x & ~31u) + owner][2], 1u);
                        job = g_job;   // not RTR_NONE: the lane is busy; the owner's lane writes the results
                        st = (g_st & 0x4FFFFu) | kHelperBit | (owner << kMasterShift);   // bounce index and bit 18 carry over
                        pend_node = RTR_NONE; L = 0.f;
                    }
                }
                float v;
                v = __shfl_sync(0xffffffffu, r.ox, partner); if (thief) r.ox = v;
                v = __shfl_sync(0xffffffffu, r.oy, partner); if (thief) r.oy = v;
                v = __shfl_sync(0xffffffffu, r.oz, partner); if (thief) r.oz = v;
                v = __shfl_sync(0xffffffffu, r.dx, partner); if (thief) r.dx = v;
                v = __shfl_sync(0xffffffffu, r.dy, partner); if (thief) r.dy = v;
                v = __shfl_sync(0xffffffffu, r.dz, partner); if (thief) r.dz = v;
                v = __shfl_sync(0xffffffffu, r.ix, partner); if (thief) r.ix = v;   // (the owner's 1/d, as computed by make_ray)
                v = __shfl_sync(0xffffffffu, r.iy, partner); if (thief) r.iy = v;
                v = __shfl_sync(0xffffffffu, r.iz, partner); if (thief) r.iz = v;
                v = __shfl_sync(0xffffffffu, best_t, partner); if (thief) best_t = v;
                v = __shfl_sync(0xffffffffu, limit, partner); if (thief) limit = v;
                {
                    const uint32_t g_bn = __shfl_sync(0xffffffffu, best_node, partner);
                    if (thief) best_node = g_bn;
                }
                const int g_sp = __shfl_sync(0xffffffffu, sp, partner);
                __syncwarp();   // the donors' stack entries, written lanes apart, are read by other lanes below
                if (thief) {
                    const uint32_t col = (threadIdx.x & ~31u) + partner;
                    int m = 0;
#pragma unroll 1
                    for (int i = 0; i < g_sp; i += 2) {   // entries 0, 2, 4 ... in order; dead ones are dropped on the way
                        const uint32_t tb = s_stack[0][i][col], w = s_stack[1][i][col];
                        if (!(__uint_as_float(tb) > limit)) { s_stack[0][m][threadIdx.x] = tb; s_stack[1][m][threadIdx.x] = w; ++m; }
                    }
                    sp = m;
                }
                __syncwarp();
                if (donor) {
                    int m = 0;
#pragma unroll 1
                    for (int i = 1; i < sp; i += 2) {     // entries 1, 3, 5 ... stay
                        const uint32_t tb = s_stack[0][i][threadIdx.x], w = s_stack[1][i][threadIdx.x];
                        if (!(__uint_as_float(tb) > limit)) { s_stack[0][m][threadIdx.x] = tb; s_stack[1][m][threadIdx.x] = w; ++m; }
                    }
                    sp = m;
                }
                __syncwarp();
                if (thief) pop_next();   // (a lane whose entries were all dead is dry: it reports back at once)
                idle = __ballot_sync(0xffffffffu, job == RTR_NONE);
            }
        }
        if (idle == 0xffffffffu) break;

        // ---- walk: one step per iteration for every lane that can move (idle lanes keep a == 0) ----
        while (true) {
#pragma unroll 1
          for (int rep = 0; rep < kWalkSteps; ++rep) {  // the warp-wide bookkeeping below runs once per kWalkSteps steps
            bool need_pop = false;
            if (a != kDry) {
                if (a & kLeafBit) {
                    if (pend_node == RTR_NONE) {  // park the leaf and keep walking
                        pend_node = a & ~kLeafBit;
                        need_pop = true;
                        if (kPrefetch & 2) prefetch_l2(A.pairs + (size_t)pend_node * 4);  // its record will be wanted at the leaf step
                    }
                } else {
                    // compressed child pair: one 256-bit load (LDG.E.256, sm_100+)
                    uint4 c0, c1;
                    ld_nc_256(A.pairs + (size_t)a * 4, c0, c1);
                    uint4 c2 = make_uint4(0u, 0u, 0u, 0u), c3 = c2;
                    const bool wide_ray = kWide && !(st & 0x40000u);
                    if (wide_ray) ld_nc_256(A.pairs + (size_t)a * 4 + 2, c2, c3);  // both halves in flight together
                    float tl = -INFINITY, fl = INFINITY, tr = -INFINITY, fr = INFINITY;
                    const uint32_t flags = c0.w >> 24;
                    bool hl, hr;
                    if (wide_ray && (flags & 0x80u)) {
                        // ---- wide step: the four grandchild slots, nearest first ----
                        // entry distance clamped at 0 and exit distance clamped at the pruning limit (>= 0 whenever a hit is
                        // still possible): "exit >= 0, entry <= exit, entry <= limit" becomes one comparison per slot
                        float tn[4], tf[4], vn[4], vf[4], un[4], uf[4];
                        axis_vals4(__uint_as_float(c0.x), c0.w & 0xFFu, c2.x, c3.x, r.ox, r.ix, vn, vf);
#pragma unroll
                        for (int k = 0; k < 4; ++k) { tn[k] = fmaxf(vn[k], 0.f); tf[k] = fminf(vf[k], limit); }
                        axis_vals4(__uint_as_float(c0.y), (c0.w >> 8) & 0xFFu, c2.y, c3.y, r.oy, r.iy, vn, vf);
                        axis_vals4(__uint_as_float(c0.z), (c0.w >> 16) & 0xFFu, c2.z, c3.z, r.oz, r.iz, un, uf);
#pragma unroll
                        for (int k = 0; k < 4; ++k) { tn[k] = max3f(tn[k], vn[k], un[k]); tf[k] = min3f(tf[k], vf[k], uf[k]); }
                        const uint32_t used = 0x5u | ((flags & 1u) ? 0u : 2u) | ((flags & 2u) ? 0u : 8u);  // slots 1 / 3 exist iff L / R is inner
                        float key[4];
                        uint32_t wd[4];
                        wd[0] = (a + 2u - (flags & 1u)) | ((flags & 0x08u) ? kLeafBit : 0u);
                        wd[1] = c2.w;  // slots 1 and 3 carry their leaf bit in the stored index
                        wd[2] = (c1.w + 1u - ((flags >> 1) & 1u)) | ((flags & 0x20u) ? kLeafBit : 0u);
                        wd[3] = c3.w;
#pragma unroll
                        for (int k = 0; k < 4; ++k) key[k] = (((used >> k) & 1u) && tn[k] <= tf[k]) ? tn[k] : INFINITY;
                        auto cswap = [&](int i, int j) {
                            const bool sw = key[j] < key[i];
                            const float ki = key[i], kj = key[j];
                            const uint32_t wi = wd[i], wj = wd[j];
                            key[i] = sw ? kj : ki; key[j] = sw ? ki : kj;
                            wd[i] = sw ? wj : wi; wd[j] = sw ? wi : wj;
                        };
                        cswap(0, 1); cswap(2, 3); cswap(0, 2); cswap(1, 3); cswap(1, 2);  // a partial order is slower (+2..3 %)
                        if (sp + 3 <= kSmemStack) {  // the usual case: no overflow checks per entry
#pragma unroll
                            for (int k = 3; k >= 1; --k)
                                if (key[k] < INFINITY) {
                                    s_stack[0][sp][threadIdx.x] = __float_as_uint(key[k]); s_stack[1][sp][threadIdx.x] = wd[k];
                                    ++sp;
                                }
                        } else {
                            if (key[3] < INFINITY) push_far(key[3], wd[3]);
                            if (key[2] < INFINITY) push_far(key[2], wd[2]);
                            if (key[1] < INFINITY) push_far(key[1], wd[1]);
                        }
                        if (key[0] < INFINITY) {
                            if (kEagerPark && (wd[0] & kLeafBit) != 0u && pend_node == RTR_NONE) {  // park the nearest slot right away
                                pend_node = wd[0] & ~kLeafBit;
                                if (kPrefetch & 2) prefetch_l2(A.pairs + (size_t)pend_node * 4);
                                need_pop = true;
                            } else a = wd[0];
                        } else need_pop = true;
                        hl = hr = false;
                    } else {
                    if (st & 0x40000u) {
                        // a direction component is 0 (or nearly): the reference's own slab test -- it is what defines
                        // the outcome for a ray parallel to a slab (Q11) -- on the decoded superset boxes
                        float4 llo, lhi, rlo, rhi;
                        decode_axis(__uint_as_float(c0.x), c0.w & 0xFFu, c1.x, llo.x, lhi.x, rlo.x, rhi.x);
                        decode_axis(__uint_as_float(c0.y), (c0.w >> 8) & 0xFFu, c1.y, llo.y, lhi.y, rlo.y, rhi.y);
                        decode_axis(__uint_as_float(c0.z), (c0.w >> 16) & 0xFFu, c1.z, llo.z, lhi.z, rlo.z, rhi.z);
                        hl = intersect_box(r, llo, lhi, tl);
                        hr = intersect_box(r, rlo, rhi, tr);
                        if (flags & 4u) { hl = hr = true; tl = tr = -INFINITY; }
                        hl = hl && !(tl > limit);
                        hr = hr && !(tr > limit);
                    } else {
                        axis_planes(__uint_as_float(c0.x), c0.w & 0xFFu, c1.x, r.ox, r.ix, tl, fl, tr, fr);
                        axis_planes(__uint_as_float(c0.y), (c0.w >> 8) & 0xFFu, c1.y, r.oy, r.iy, tl, fl, tr, fr);
                        axis_planes(__uint_as_float(c0.z), (c0.w >> 16) & 0xFFu, c1.z, r.oz, r.iz, tl, fl, tr, fr);
                        if (flags & 4u) { tl = tr = -INFINITY; fl = fr = INFINITY; }  // box outside the encodable range
                        hl = fl >= 0.f && tl <= fl && !(tl > limit);
                        hr = fr >= 0.f && tr <= fr && !(tr > limit);
                    }
                    const uint32_t lw = (a + 1u) | ((flags & 1u) ? kLeafBit : 0u);
                    const uint32_t rw = c1.w | ((flags & 2u) ? kLeafBit : 0u);
                    if (hl && hr) {
                        const bool left_first = tl <= tr;
                        push_far(left_first ? tr : tl, left_first ? rw : lw);
                        // the stacked child is visited later, if at all: start moving its record towards the SM now
                        if (kPrefetch & 1) prefetch_l2(A.pairs + (size_t)((left_first ? rw : lw) & ~kLeafBit) * 4);
                        a = left_first ? lw : rw;
                    } else if (hl) a = lw;
                    else if (hr) a = rw;
                    else need_pop = true;
                    }
                }
            }
            if (need_pop) pop_next();  // one site: lanes coming from a parked leaf and from a double miss pop together
          }
            const bool busy = job != RTR_NONE;
            const bool blocked = a != kDry && (a & kLeafBit) != 0u && pend_node != RTR_NONE;
            const uint32_t m_blocked = __ballot_sync(0xffffffffu, blocked);
            const uint32_t m_pend = __ballot_sync(0xffffffffu, pend_node != RTR_NONE);
            const uint32_t m_walk = __ballot_sync(0xffffffffu, a != kDry);
            const uint32_t m_dry = __ballot_sync(0xffffffffu, busy && a == kDry);
            const uint32_t movable = m_walk & ~m_blocked;
            RTR_STAT(exhausted ? 4 : 6, 1); RTR_STAT(exhausted ? 5 : 7, __popc(m_walk));
            const bool want_fin = m_dry != 0u && ((uint32_t)__popc(m_dry) >= kFinBatch || (uint32_t)__popc(m_walk) < 16u);
            if ((uint32_t)__popc(m_blocked) >= kBlockBatch || (uint32_t)__popc(m_pend) >= kLeafBatch || movable == 0u || (want_fin && (m_dry & m_pend) != 0u)) {
                // ---- leaf step: every parked leaf of the warp.  The reference's own two tests, on the exact
                //      box and the vertices of the 64-byte leaf record ----
                if (pend_node != RTR_NONE) {
                    uint4 c0, c1, c2, c3;
                    ld_nc_256(A.pairs + (size_t)pend_node * 4, c0, c1);
                    ld_nc_256(A.pairs + (size_t)pend_node * 4 + 2, c2, c3);
                    // Both tests are pure: "box passes, then triangle hits" == "triangle hits, and box passes".  The
                    // compressed parent already let this leaf through, so the exact box test rejects few and the
                    // triangle test most: the triangle goes first, the box is only consulted for an actual hit.
                    TriWorld w;  // shader naming (Q8): _P1 = host P2, _P2 = host P1
                    w.p0 = make_float3(__uint_as_float(c1.z), __uint_as_float(c1.w), __uint_as_float(c2.x));
                    w.p2 = make_float3(__uint_as_float(c2.y), __uint_as_float(c2.z), __uint_as_float(c2.w));
                    w.p1 = make_float3(__uint_as_float(c3.x), __uint_as_float(c3.y), __uint_as_float(c3.z));
                    Hit h;
                    if (ray_triangle_w(r, w, 0u, 0u, h)) {
                        const bool any_mode = (st & 0x10000u) != 0u;
                        const bool better = h.t < best_t ||
                                            (!any_mode && h.t == best_t && best_node != RTR_NONE && pend_node > best_node);
                        if (better) {
                            const float4 blo = make_float4(__uint_as_float(c0.x), __uint_as_float(c0.y), __uint_as_float(c0.z), 0.f);
                            const float4 bhi = make_float4(__uint_as_float(c0.w), __uint_as_float(c1.x), __uint_as_float(c1.y), 0.f);
                            float te;
                            if (intersect_box(r, blo, bhi, te)) {
                                best_node = pend_node;
                                if (any_mode) { a = kDry; sp = 0; }
                                else { best_t = h.t; limit = limit_of(h.t); }
                            }
                        }
                    }
                    pend_node = RTR_NONE;
                }
            }
            if (want_fin || (movable == 0u && m_blocked == 0u)) break;
            if (kSteal && exhausted && m_walk != 0xffffffffu) break;   // drain: lanes are free to take over stacked entries
        }

        // ---- rays whose stack ran dry: shade, continue the path or retire the job ----
        if (kSteal && exhausted) {
            const bool helper = (st & kHelperBit) != 0u;
            const uint32_t own = helper ? ((threadIdx.x & ~31u) + ((st >> kMasterShift) & 31u)) : threadIdx.x;
            const bool shared = job != RTR_NONE && (st & 0x10000u) == 0u && (helper || s_lane[threadIdx.x][2] != 0u);
            if (__any_sync(0xffffffffu, shared)) {
                // the lanes of a ray publish their best hit and prune with the best any of them has found
                if (shared && best_node != RTR_NONE) atomicMax(lane_best(own), pack_best(best_t, best_node));
                __syncwarp();
                if (shared) {
                    const unsigned long long p = *lane_best(own);
                    if (p != 0ull) limit = fminf(limit, limit_of(__uint_as_float(~(uint32_t)(p >> 32))));
                }
                // a helper whose stack ran dry has published what it found: it falls idle
                if (helper && a == kDry && pend_node == RTR_NONE) {
                    if (st & 0x20000u) atomicAdd(&tp->stack_overflows, 1u);
                    atomicSub(&s_lane[own][2], 1u);
                    job = RTR_NONE; st = 0u;
                }
                __syncwarp();
            }
        }
        const bool fin = job != RTR_NONE && a == kDry && pend_node == RTR_NONE && (st & kHelperBit) == 0u &&
                         (!kSteal || !exhausted || (st & 0x10000u) != 0u || s_lane[threadIdx.x][2] == 0u);   // (any-hit: the words are the stash)
        if (kSteal && exhausted && fin && (st & 0x10000u) == 0u) {   // what the helpers of this ray found
            const unsigned long long p = *lane_best(threadIdx.x);
            if (p != 0ull) {
                const float h_t = __uint_as_float(~(uint32_t)(p >> 32));
                const uint32_t h_n = (uint32_t)p;
                if (best_node == RTR_NONE || h_t < best_t || (h_t == best_t && h_n > best_node)) { best_t = h_t; best_node = h_n; }
                *lane_best(threadIdx.x) = 0ull;
            }
        }
        traced += __popc(__ballot_sync(0xffffffffu, fin));
        if (fin) {
            const bool any_mode = (st & 0x10000u) != 0u;
            // regenerate the record of the best hit (same inputs, same ops => same bits)
            Hit best = no_hit();
            TriWorld bw;
            bw.p0 = bw.p1 = bw.p2 = make_float3(0.f, 0.f, 0.f);
            if (best_node != RTR_NONE) {
                if (any_mode) best.did_hit = 1u;
                else {
                    uint4 c0, c1, c2, c3;
                    ld_nc_256(A.pairs + (size_t)best_node * 4, c0, c1);
                    ld_nc_256(A.pairs + (size_t)best_node * 4 + 2, c2, c3);
                    bw.p0 = make_float3(__uint_as_float(c1.z), __uint_as_float(c1.w), __uint_as_float(c2.x));
                    bw.p2 = make_float3(__uint_as_float(c2.y), __uint_as_float(c2.z), __uint_as_float(c2.w));
                    bw.p1 = make_float3(__uint_as_float(c3.x), __uint_as_float(c3.y), __uint_as_float(c3.z));
                    ray_triangle_w(r, bw, 0u, c3.w, best);
                }
            }
            if (jd.kind != 0u) {
                store_hit(jd.hits, job, best);
                if (st & 0x20000u) atomicAdd(&tp->stack_overflows, 1u);  // never silent: the host forms refuse the batch
                job = RTR_NONE;
            } else {
                uint32_t x, y, out_row;
                job_pixel(job, x, y, out_row);
                const size_t o = (size_t)out_row * jd.width + x;
                const uint32_t k = st & 0xFFFFu;
                bool path_done = false, shade = false;
                float nx = 0.f, ny = 0.f, nz = 0.f, ox = r.ox, oy = r.oy, oz = r.oz, dx = 0.f, dy = 0.f, dz = 0.f, c = 0.f;
                if (!any_mode) {
                    if (k == 0u && jd.hits) store_hit(jd.hits, o, best);
                    if (!best.did_hit) {
                        path_done = true;
                    } else {
                        hit_frame_w(r, best.t, bw, nx, ny, nz, ox, oy, oz);
                        dx = r.dx; dy = r.dy; dz = r.dz;
                        if (jd.shadow) {
                            const float sx = __fsub_rn(jd.lx, ox), sy = __fsub_rn(jd.ly, oy), sz = __fsub_rn(jd.lz, oz);
                            const float len = __fsqrt_rn(dot3(sx, sy, sz, sx, sy, sz));
                            s_lane[threadIdx.x][0] = __float_as_uint(nx); s_lane[threadIdx.x][1] = __float_as_uint(ny); s_lane[threadIdx.x][2] = __float_as_uint(nz);
                            s_lane[threadIdx.x][3] = __float_as_uint(dx); s_lane[threadIdx.x][4] = __float_as_uint(dy); s_lane[threadIdx.x][5] = __float_as_uint(dz);
                            r = make_ray(ox, oy, oz, __fdiv_rn(sx, len), __fdiv_rn(sy, len), __fdiv_rn(sz, len));
                            start_ray(true, len);
                        } else {
                            c = -dot3(nx, ny, nz, dx, dy, dz);
                            shade = true;
                        }
                    }
                } else {  // shadow ray finished; its origin is the bounce origin
                    nx = __uint_as_float(s_lane[threadIdx.x][0]); ny = __uint_as_float(s_lane[threadIdx.x][1]); nz = __uint_as_float(s_lane[threadIdx.x][2]);
                    dx = __uint_as_float(s_lane[threadIdx.x][3]); dy = __uint_as_float(s_lane[threadIdx.x][4]); dz = __uint_as_float(s_lane[threadIdx.x][5]);
                    s_lane[threadIdx.x][0] = 0u; s_lane[threadIdx.x][1] = 0u; s_lane[threadIdx.x][2] = 0u;   // free for the drain's use
                    const float ndl = dot3(nx, ny, nz, r.dx, r.dy, r.dz);
                    c = best.did_hit ? 0.f : fmaxf(0.f, ndl);
                    shade = true;
                }
                if (shade) {
                    const float wgt = ldexpf(1.f, -(int)k);  // 0.5^k, the product of k exact halvings
                    L = __fadd_rn(L, __fmul_rn(wgt, c));
                    if (k < jd.bounces) {
                        const float kk = __fmul_rn(2.f, dot3(dx, dy, dz, nx, ny, nz));
                        float rx = __fsub_rn(dx, __fmul_rn(nx, kk));
                        float ry = __fsub_rn(dy, __fmul_rn(ny, kk));
                        float rz = __fsub_rn(dz, __fmul_rn(nz, kk));
                        normalize3(rx, ry, rz);
                        r = make_ray(ox, oy, oz, rx, ry, rz);
                        st = (st & ~0xFFFFu) | (k + 1u);
                        start_ray(false, INFINITY);
                    } else {
                        path_done = true;
                    }
                }
                if (path_done) {
                    if (jd.rgba) jd.rgba[o] = make_float4(L, L, L, 1.f);
                    if (st & 0x20000u) atomicAdd(&tp->stack_overflows, 1u);
                    job = RTR_NONE;
                }
            }
        }
    }
    if (rays_traced && lane == 0u && traced) atomicAdd(rays_traced, (unsigned long long)traced);
}

// intersectBVH of raytracer.glsl:182-237 with its return code 2 (the ray enters within BVH_LINE_WIDTH / (depth + 1) of
// two of the three slab pairs: a box edge)
__device__ __forceinline__ uint32_t intersect_box_edge(const Ray& r, const float4 lo, const float4 hi, float threshold) {
    float t_entry;
    if (!intersect_box(r, lo, hi, t_entry)) return 0u;
    const float ex = __fadd_rn(r.ox, __fmul_rn(r.dx, t_entry));
    const float ey = __fadd_rn(r.oy, __fmul_rn(r.dy, t_entry));
    const float ez = __fadd_rn(r.oz, __fmul_rn(r.dz, t_entry));
    const bool cx = fabsf(__fsub_rn(ex, lo.x)) < threshold || fabsf(__fsub_rn(ex, hi.x)) < threshold;
    const bool cy = fabsf(__fsub_rn(ey, lo.y)) < threshold || fabsf(__fsub_rn(ey, hi.y)) < threshold;
    const bool cz = fabsf(__fsub_rn(ez, lo.z)) < threshold || fabsf(__fsub_rn(ez, hi.z)) < threshold;
    return ((cx && cy) || (cx && cz) || (cy && cz)) ? 2u : 1u;
}

// The bvhColor of getClosestHitBVH (raytracer.glsl:246-295): every intersected node at depth == display_depth
// overwrites the colour, so the node the shader visits LAST decides.  The shader pops the right child first and does
// not prune, and a node is only reached through intersected ancestors (which every intersected node has: merged
// boxes are unions and the slab test is monotone): the last one visited is the intersected depth-D node with the
// smallest flat index -- the first one a LEFT-first walk down to depth D meets.  One thread per pixel.
__global__ void __launch_bounds__(kTraceBlock)
depth_overlay_kernel(const rtr_node* __restrict__ nodes, const rtr_camera cam, uint32_t width, uint32_t height,
                     uint32_t denom_w, uint32_t denom_h, int display_depth, float4* __restrict__ out) {
    uint32_t x, y;
    if (!pixel_of_thread(width, height, x, y)) return;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x < denom_w && y < denom_h && display_depth >= 0 && display_depth < kStack) {
        const Ray r = camera_ray(cam, x, y, denom_w, denom_h);
        const float threshold = __fdiv_rn(0.05f, __fadd_rn((float)display_depth, 1.f));
        uint32_t stack[kStack];   // right siblings still to try ...
        uint8_t sdepth[kStack];   // ... and their depth (< kStack)
        int sp = 0;
        uint32_t cur = 0u, depth = 0u;
        while (true) {
            const NodeRec nd = load_node(nodes, cur);
            const uint32_t code = intersect_box_edge(r, nd.lo, nd.hi, threshold);
            bool descend = false;
            if (code != 0u) {
                if ((int)depth == display_depth) {
                    c = code == 2u ? make_float4(0.7f, 0.f, 0.7f, 0.1f) : make_float4(0.5f, 0.f, 0.5f, 0.1f);
                    break;
                }
                if (!is_leaf(nd.links)) {
                    stack[sp] = nd.links.z; sdepth[sp] = (uint8_t)(depth + 1u); ++sp;   // sp <= depth < display_depth < kStack
                    cur = nd.links.y; depth += 1u;
                    descend = true;
                }
            }
            if (!descend) {
                if (sp == 0) break;
                --sp;
                cur = stack[sp]; depth = sdepth[sp];
            }
        }
    }
    out[(size_t)y * width + x] = c;
}

// getColor + main of raytracer.glsl (:159-179, :299-331): the reference's frame from hit records
__global__ void __launch_bounds__(256)
shade_kernel(const rtr_hit* __restrict__ hits, uint64_t n, const rtr_triangle* __restrict__ tris,
             const rtr_mesh* __restrict__ meshes, const rtr_material* __restrict__ materials, uint32_t flags,
             const float4* __restrict__ bvh_rgba, float4* __restrict__ rgba) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2* hp = reinterpret_cast<const uint2*>(hits + i);  // 24-byte records: three 8-byte loads
    const uint2 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2);
    float4 c = make_float4(0.f, 0.f, 0.f, 1.f);
    if ((flags & RTR_SHADE_BVH) && bvh_rgba) c = __ldg(bvh_rgba + i);   // uIsBVHDisplayed, :161-163
    if (h2.x != 0u) {
        const uint32_t model = __ldg(&tris[h2.y].model_id);
        const float* m = materials[__ldg(&meshes[model].material_id)].color;
        c.x = __fadd_rn(c.x, __ldg(m)); c.y = __fadd_rn(c.y, __ldg(m + 1));
        c.z = __fadd_rn(c.z, __ldg(m + 2)); c.w = __fadd_rn(c.w, __ldg(m + 3));
        if (flags & RTR_SHADE_WIREFRAME) {
            const float th = 0.02f;  // WIREFRAME_LINE_WIDTH, raytracer.glsl:71
            if (__uint_as_float(h0.x) < th || __uint_as_float(h0.y) < th || __uint_as_float(h1.x) < th)
                c = make_float4(0.f, 0.f, 0.f, 1.f);
        }
    }
    rgba[i] = c;
}

inline Accel accel_of(const rtr_bvh* b) {
    Accel A;
    A.nodes = b->flat_view; A.wtri = b->wtri_view; A.pairs = b->pairs_view; A.by_rank = b->wtri_by_rank ? 1u : 0u;
    return A;
}

// one CTA per resident slot of the device
int launch_persistent(rtr_ctx* ctx, const rtr_bvh* b, const JobDesc& jd, uint64_t* rays) {
    if (ctx->trace_ctas_per_sm == 0) {  // per context = per device (one rtr_ctx per (host thread, GPU))
        int c = 0;
        RTR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, trace_persistent_kernel, kTraceBlock, 0));
        ctx->trace_ctas_per_sm = c < 1 ? 1 : c;
    }
    const int ctas_per_sm = ctx->trace_ctas_per_sm;
    const uint32_t need = (jd.total + kTraceBlock - 1) / kTraceBlock;
    // a launch on a stream of the ctx's SM partition (rtr_ctx_partition_sms) fills the partition and nothing else
    bool in_partition = false;
    for (cudaStream_t st : ctx->partition_streams) in_partition = in_partition || st == ctx->stream;
    const int sms = in_partition ? ctx->partition_sms : ctx->sm_count;
    const int keep_free = in_partition ? 0 : ctx->reserved_sms;
    const uint32_t full = (uint32_t)(sms * ctas_per_sm);
    uint32_t grid = full, reserve = 0u;
    if (need <= (uint32_t)((sms - keep_free) * ctas_per_sm)) grid = need;  // small job: room is left anyway
    else reserve = (uint32_t)keep_free;
    if (grid == 0) return RTR_OK;
    uint32_t* table = nullptr;
    if (reserve != 0u) {  // two tables in turn: consecutive launches may run on different streams and overlap
        if (!ctx->sm_table) RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->sm_table), 2 * 1025 * sizeof(uint32_t)));
        table = ctx->sm_table + 1025 * (ctx->sm_table_turn++ & 1u);
        RTR_CUDA(ctx, cudaMemsetAsync(table, 0, 1025 * sizeof(uint32_t), ctx->stream));
    }
    RTR_CUDA(ctx, cudaMemsetAsync(&b->tparams->job_counter, 0, sizeof(uint32_t), ctx->stream));
    trace_persistent_kernel<<<grid, kTraceBlock, 0, ctx->stream>>>(accel_of(b), b->tparams, jd,
                                                                   reinterpret_cast<unsigned long long*>(rays), table, reserve);
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}

}  // namespace

#ifdef RTR_STEAL_STATS
extern "C" int rtr_debug_steal_stats(unsigned long long out[8], int reset) {
    if (out && cudaMemcpyFromSymbol(out, g_steal_stats, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_steal_stats, z, sizeof(z)); }
    return 0;
}
#endif

static void resolve_denoms(uint32_t width, uint32_t height, uint32_t& dw, uint32_t& dh) {
    if (dw == 0) dw = (width / 16) * 16;   // application.cpp:225-226 + raytracer.glsl:304-305 (Q5)
    if (dh == 0) dh = (height / 16) * 16;
}

static JobDesc pixel_jobs(const rtr_camera& cam, uint32_t width, uint32_t denom_w, uint32_t denom_h, const RowMap& rm,
                          uint32_t bounces, int shadow, const float light[3], float* rgba, rtr_hit* hits) {
    JobDesc jd;
    memset(&jd, 0, sizeof(jd));
    jd.kind = 0u;
    jd.cam = cam; jd.width = width; jd.denom_w = denom_w; jd.denom_h = denom_h; jd.rm = rm;
    jd.bounces = bounces; jd.shadow = shadow;
    jd.lx = light ? light[0] : 0.f; jd.ly = light ? light[1] : 0.f; jd.lz = light ? light[2] : 0.f;
    jd.rgba = reinterpret_cast<float4*>(rgba); jd.hits = hits;
    jd.total = ((width + 7) / 8) * ((rm.rows + 3) / 4) * 32u;
    return jd;
}

int rtr_trace_primary_launch(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera& cam, uint32_t width, uint32_t height,
                             uint32_t denom_w, uint32_t denom_h, uint32_t row0, uint32_t row1, uint32_t flags,
                             rtr_hit* hits) {
    resolve_denoms(width, height, denom_w, denom_h);
    if (row1 == 0) row1 = height;
    if (width == 0 || row0 >= row1 || row1 > height || denom_w == 0 || denom_h == 0)
        return rtr_set_error(ctx, RTR_E_INVALID, "trace_primary: bad image geometry %ux%u rows [%u,%u) denom %ux%u",
                             width, height, row0, row1, denom_w, denom_h);
    const uint32_t rows = row1 - row0;
    if ((uint64_t)((width + 7) / 8) * ((rows + 3) / 4) * 32u >= 0xFFFFFF00ull)
        return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "trace_primary: image too large");
    RTR_PROF(ctx, "trace_primary_kernel");
    if (flags & (RTR_TRACE_REFERENCE_ORDER | RTR_TRACE_DEEP_STACK)) {
        if (flags & RTR_TRACE_DEEP_STACK)
            trace_primary_kernel<2><<<pixel_grid(width, rows), kTraceBlock, 0, ctx->stream>>>(
                accel_of(b), b->tparams, cam, width, denom_w, denom_h, row0, rows, hits);
        else
            trace_primary_kernel<0><<<pixel_grid(width, rows), kTraceBlock, 0, ctx->stream>>>(
                accel_of(b), b->tparams, cam, width, denom_w, denom_h, row0, rows, hits);
        RTR_LAUNCH_CHECK(ctx);
        return RTR_OK;
    }
    RowMap rm;
    rm.row0 = row0; rm.rows = rows; rm.height = height; rm.rpb = 1; rm.count = 1; rm.span = 1; rm.off[0] = 0;
    return launch_persistent(ctx, b, pixel_jobs(cam, width, denom_w, denom_h, rm, 0u, 0, nullptr, nullptr, hits), nullptr);
}

int rtr_trace_rays_launch(rtr_ctx* ctx, const rtr_bvh* b, const rtr_ray* rays, uint64_t n_rays, int any, const float* t_max,
                          uint32_t flags, rtr_hit* hits) {
    if (n_rays == 0) return RTR_OK;
    if (n_rays >= 0xFFFFFF00ull) return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "trace_rays: too many rays");
    RTR_PROF(ctx, "trace_rays_kernel");
    if (flags & (RTR_TRACE_REFERENCE_ORDER | RTR_TRACE_DEEP_STACK)) {
        const uint32_t grid = (uint32_t)((n_rays + kTraceBlock - 1) / kTraceBlock);
        if (flags & RTR_TRACE_DEEP_STACK)
            trace_rays_kernel<2><<<grid, kTraceBlock, 0, ctx->stream>>>(accel_of(b), b->tparams, rays, n_rays, any, t_max, hits);
        else
            trace_rays_kernel<0><<<grid, kTraceBlock, 0, ctx->stream>>>(accel_of(b), b->tparams, rays, n_rays, any, t_max, hits);
        RTR_LAUNCH_CHECK(ctx);
        return RTR_OK;
    }
    JobDesc jd;
    memset(&jd, 0, sizeof(jd));
    jd.kind = 1u; jd.total = (uint32_t)n_rays; jd.rays = rays; jd.t_max = t_max; jd.want_any = any; jd.hits = hits;
    return launch_persistent(ctx, b, jd, nullptr);
}

int rtr_render_launch(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera& cam, uint32_t width, uint32_t height,
                      uint32_t denom_w, uint32_t denom_h, uint32_t row0, uint32_t row1, uint32_t bounces, int shadow,
                      const float light[3], uint32_t flags, float* rgba, rtr_hit* hits, uint64_t* rays,
                      uint32_t rows_per_block, uint32_t total_stripes, uint32_t nb_stripes, const uint8_t* stripe_offsets) {
    resolve_denoms(width, height, denom_w, denom_h);
    if (row1 == 0) row1 = height;
    if (width == 0 || row0 >= row1 || row1 > height || denom_w == 0 || denom_h == 0)
        return rtr_set_error(ctx, RTR_E_INVALID, "render: bad image geometry %ux%u rows [%u,%u) denom %ux%u", width,
                             height, row0, row1, denom_w, denom_h);
    if (bounces > 0xFFFFu)  // the lane state keeps the bounce index in 16 bits
        return rtr_set_error(ctx, RTR_E_INVALID, "render: %u bounces (at most 65535)", bounces);
    RowMap rm;
    memset(&rm, 0, sizeof(rm));
    rm.row0 = row0; rm.rows = row1 - row0; rm.height = height; rm.rpb = 1; rm.count = 1; rm.span = 1; rm.off[0] = 0;
    if (total_stripes > 1) {
        if (nb_stripes == 0) return RTR_OK;  // a rank that owns no stripes (e.g. the one that rebuilds the BVH)
        if (rows_per_block == 0 || nb_stripes > RTR_MAX_STRIPES_PER_RANK || total_stripes > 255 || !stripe_offsets)
            return rtr_set_error(ctx, RTR_E_INVALID, "render: bad stripes %u of %u, rows_per_block %u", nb_stripes,
                                 total_stripes, rows_per_block);
        uint32_t first = total_stripes;
        for (uint32_t i = 0; i < nb_stripes; ++i) {
            if (stripe_offsets[i] >= total_stripes) return rtr_set_error(ctx, RTR_E_INVALID, "render: stripe offset out of range");
            rm.off[i] = stripe_offsets[i];
            if (stripe_offsets[i] < first) first = stripe_offsets[i];
        }
        const uint32_t blocks = (height + rows_per_block - 1) / rows_per_block;
        // local blocks are enumerated cycle by cycle: up to the last cycle that still holds one of this rank's blocks
        const uint32_t cycles = blocks > first ? (blocks - first + total_stripes - 1) / total_stripes : 0;
        if (cycles == 0) return RTR_OK;
        rm.row0 = 0; rm.rows = cycles * nb_stripes * rows_per_block; rm.rpb = rows_per_block;
        rm.count = total_stripes; rm.span = nb_stripes;
    }
    if ((uint64_t)((width + 7) / 8) * ((rm.rows + 3) / 4) * 32u >= 0xFFFFFF00ull)
        return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "render: image too large");
    RTR_PROF(ctx, "render_kernel");
    if (flags & (RTR_TRACE_REFERENCE_ORDER | RTR_TRACE_DEEP_STACK)) {
        const float lx = light ? light[0] : 0.f, ly = light ? light[1] : 0.f, lz = light ? light[2] : 0.f;
        if (flags & RTR_TRACE_DEEP_STACK)
            render_kernel<2><<<pixel_grid(width, rm.rows), kTraceBlock, 0, ctx->stream>>>(
                accel_of(b), b->tparams, cam, width, denom_w, denom_h, rm, bounces, shadow, lx, ly, lz,
                reinterpret_cast<float4*>(rgba), hits, reinterpret_cast<unsigned long long*>(rays));
        else
            render_kernel<0><<<pixel_grid(width, rm.rows), kTraceBlock, 0, ctx->stream>>>(
                accel_of(b), b->tparams, cam, width, denom_w, denom_h, rm, bounces, shadow, lx, ly, lz,
                reinterpret_cast<float4*>(rgba), hits, reinterpret_cast<unsigned long long*>(rays));
        RTR_LAUNCH_CHECK(ctx);
        return RTR_OK;
    }
    return launch_persistent(ctx, b, pixel_jobs(cam, width, denom_w, denom_h, rm, bounces, shadow, light, rgba, hits), rays);
}

int rtr_depth_overlay_launch(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera& cam, uint32_t width, uint32_t height,
                             uint32_t denom_w, uint32_t denom_h, int display_depth, float* bvh_rgba) {
    resolve_denoms(width, height, denom_w, denom_h);
    if (width == 0 || height == 0 || denom_w == 0 || denom_h == 0)
        return rtr_set_error(ctx, RTR_E_INVALID, "depth_overlay: bad image geometry %ux%u", width, height);
    if (display_depth >= kStack) return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "depth_overlay: depth %d >= %d", display_depth, kStack);
    RTR_PROF(ctx, "depth_overlay_kernel");
    depth_overlay_kernel<<<pixel_grid(width, height), kTraceBlock, 0, ctx->stream>>>(
        b->flat_view, cam, width, height, denom_w, denom_h, display_depth, reinterpret_cast<float4*>(bvh_rgba));
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}

int rtr_shade_launch(rtr_ctx* ctx, const rtr_hit* hits, uint64_t n, const rtr_triangle* tris, const rtr_mesh* meshes,
                     const rtr_material* materials, uint32_t flags, const float* bvh_rgba, float* rgba) {
    if (n == 0) return RTR_OK;
    if (n >= (1ull << 40)) return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "shade: too many pixels");
    RTR_PROF(ctx, "shade_kernel");
    shade_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(hits, n, tris, meshes, materials, flags,
                                                                       reinterpret_cast<const float4*>(bvh_rgba),
                                                                       reinterpret_cast<float4*>(rgba));
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}
