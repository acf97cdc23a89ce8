// ref_raytracer.cpp -- C-ABI driver around the REFERENCE's own compute shader
// (srcCommon/shaders/raytracer.glsl), compiled as C++ through the GLM 0.9.9.9 the
// reference vendors (srcVulkan/dep/slang/external/glm).  TEST INFRASTRUCTURE ONLY:
// it is linked into oracle/_ref/libref_raytracer*.so, which pins the traversal half
// of the C restatement (rtr_oracle.c), generates tests/golden/raytracer_golden.npz
// and is compared with the CUDA path directly in the -m gpu tests.
//
// No shader text lives in this repo.  oracle/Makefile pipes the shader, where it
// lies under /root/reference, through oracle/glsl2cpp.sed -- a purely syntactic
// GLSL -> C++ mapping (layout/uniform/buffer qualifiers dropped, SSBO arrays ->
// pointers, `inout T x` -> `T& x`, swizzles `.xyz` -> `.xyz()`, `main` renamed) --
// into oracle/_ref/raytracer_glsl.inc, which is included below and removed after the
// compile.  Every function body of the shader (getRay :92-100, rayTriangleIntersection
// :102-147, getAllHits :149-157, getColor :159-179, intersectBVH :182-237, isLeafBVH
// :239-244, getClosestHitBVH :246-295, main :299-331) is therefore the reference's
// own text, executed with GLM's vec/mat arithmetic (the same library the reference's
// C++ side uses).
//
// Two things GLSL leaves open are pinned by compile-time switches, and each variant
// is built as its own library:
//  * normalize().  GLSL does not define its rounding.  RTR_NORMALIZE_GLM (default):
//    glm::normalize = v * (1 / sqrt(dot(v, v))).  RTR_NORMALIZE_DIV: v / sqrt(dot(v, v)),
//    the formula the oracle and the kernels pin; with it the
//    restatement must equal this library bit for bit.
//  * the value of an uninitialised local (`Hit hit;` at :103, whose _Coords.w is read
//    at :279 after a miss -- SURVEY App. B Q6).  RTR_UNDEFINED_W: +inf (default; a miss
//    then never replaces a hit, which is exactly the guarded semantics of the shader's
//    own getAllHits :152-153) or 0 ("zeroed registers", libref_raytracer_zero.so: what
//    the letter of :279 does on hardware that hands out zeroed registers).
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <cmath>
#include <omp.h>

#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>

#ifndef RTR_UNDEFINED_W
#define RTR_UNDEFINED_W std::numeric_limits<float>::infinity()
#endif

namespace shader {
using glm::ivec2;
using glm::mat4;
using glm::uvec3;
using glm::vec2;
using glm::vec3;
using glm::vec4;
using glm::abs;
using glm::cross;
using glm::dot;
using glm::max;
using glm::min;
typedef unsigned int uint;

#ifdef RTR_NORMALIZE_DIV
// non-template overloads: they win over glm::normalize, which argument-dependent lookup also finds
static inline vec3 normalize(const vec3& v) { return v / std::sqrt(dot(v, v)); }
static inline vec4 normalize(const vec4& v) { return v / std::sqrt(dot(v, v)); }
#else
using glm::normalize;
#endif

// GL built-ins the shader's main() touches (:301-305, :330)
struct image2D {
    float* rgba;
    uint32_t width, height;
};
static thread_local uvec3 gl_GlobalInvocationID;
static uvec3 gl_NumWorkGroups;
static const uvec3 gl_WorkGroupSize(16, 16, 1);  // layout(local_size_x = 16, local_size_y = 16), :58
static inline void imageStore(image2D& img, ivec2 p, vec4 v) {
    if (p.x < 0 || p.y < 0 || (uint32_t)p.x >= img.width || (uint32_t)p.y >= img.height) return;  // GL: discarded
    float* o = img.rgba + 4 * ((size_t)p.y * img.width + (size_t)p.x);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}

// value of an uninitialised `Hit` (see header)
#define RTR_UNDEFINED_HIT shader::Hit{shader::vec4(0.f, 0.f, 0.f, RTR_UNDEFINED_W), 0u, 0u}

#include RTR_SHADER_INC
}  // namespace shader

namespace {
std::vector<shader::Triangle> g_tris;
std::vector<shader::Model> g_models;
std::vector<shader::Material> g_materials;
std::vector<shader::BVH_Node> g_nodes;

int g_threads = 0;  // 0: omp_get_max_threads(); set per library so the process-wide OpenMP setting is left alone
inline int threads() { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

struct HostRay { float o[4], d[4]; };
struct HostHit { float b0, b1, b2, t; uint32_t did_hit, tri; };

shader::Ray to_ray(const HostRay& r) {
    shader::Ray s;
    s._Origin = shader::vec4(r.o[0], r.o[1], r.o[2], r.o[3]);
    s._Direction = shader::vec4(r.d[0], r.d[1], r.d[2], r.d[3]);
    return s;
}
void from_hit(const shader::Hit& h, HostHit& o) {
    std::memset(&o, 0, sizeof(o));
    if (h._DidHit == 0) return;  // the fields of a miss are undefined in the shader: normalised to zero
    o.b0 = h._Coords.x; o.b1 = h._Coords.y; o.b2 = h._Coords.z; o.t = h._Coords.w;
    o.did_hit = h._DidHit; o.tri = h._TriangleId;
}
}  // namespace

extern "C" {

const char* ref_rt_variant() {
#ifdef RTR_NORMALIZE_DIV
    return "normalize=div";
#else
    return "normalize=glm";
#endif
}
float ref_rt_undefined_w() { return RTR_UNDEFINED_W; }
void ref_rt_set_threads(int n) { g_threads = n; }

// The four SSBOs of scene.cpp:112-176, decoded from the HOST byte layouts exactly as std430 would read them:
// triangles at stride 64 (the shader's struct names the second vec4 _P2 and the third _P1, Q8 -- the memory order
// is kept, the NAMES swap), nodes at stride 48 (vec3 _Min @0, vec3 _Max @16, uints @32/36/40), materials one vec4
// each.  Models: the host writes 68-byte records (mesh.hpp:12-15) and the shader reads 80-byte ones (Q7);
// model_stride = 68 decodes what the host meant, 80 what the shader's std430 layout fetches from those bytes
// (identical for model 0).
int ref_rt_set_scene(const void* tris, uint32_t nb_tris, const void* models, uint32_t nb_models,
                     uint32_t model_stride, const float* materials, uint32_t nb_materials,
                     const void* nodes, uint32_t nb_nodes) {
    if (model_stride != 68 && model_stride != 80) return -1;
    const uint8_t* tb = static_cast<const uint8_t*>(tris);
    g_tris.resize(nb_tris);
    for (uint32_t i = 0; i < nb_tris; ++i) {
        float f[12];
        uint32_t id;
        std::memcpy(f, tb + 64 * (size_t)i, 48);
        std::memcpy(&id, tb + 64 * (size_t)i + 48, 4);
        shader::Triangle t;
        t._P0 = shader::vec4(f[0], f[1], f[2], f[3]);
        t._P2 = shader::vec4(f[4], f[5], f[6], f[7]);   // second vec4 of the record
        t._P1 = shader::vec4(f[8], f[9], f[10], f[11]); // third vec4 of the record
        t._ModelId = id;
        g_tris[i] = t;
    }
    const uint8_t* mb = static_cast<const uint8_t*>(models);
    const size_t model_bytes = 68 * (size_t)nb_models;
    g_models.assign(nb_models, shader::Model{});
    for (uint32_t i = 0; i < nb_models; ++i) {
        const size_t off = (size_t)model_stride * i;
        if (off + 68 > model_bytes) break;  // beyond the uploaded bytes: the GL buffer holds zeros there
        float f[16];
        uint32_t id;
        std::memcpy(f, mb + off, 64);
        std::memcpy(&id, mb + off + 64, 4);
        shader::Model m;
        for (int c = 0; c < 4; ++c) m._ModelMatrix[c] = shader::vec4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
        m._MaterialId = id;
        g_models[i] = m;
    }
    g_materials.resize(nb_materials);
    for (uint32_t i = 0; i < nb_materials; ++i)
        g_materials[i]._Color = shader::vec4(materials[4 * i], materials[4 * i + 1], materials[4 * i + 2], materials[4 * i + 3]);
    const uint8_t* nb = static_cast<const uint8_t*>(nodes);
    g_nodes.resize(nb_nodes);
    for (uint32_t i = 0; i < nb_nodes; ++i) {
        float f[8];
        uint32_t u[3];
        std::memcpy(f, nb + 48 * (size_t)i, 32);
        std::memcpy(u, nb + 48 * (size_t)i + 32, 12);
        shader::BVH_Node n;
        n._BoundingBox._Min = shader::vec3(f[0], f[1], f[2]);
        n._BoundingBox._Max = shader::vec3(f[4], f[5], f[6]);
        n._TriangleId = u[0]; n._LeftChild = u[1]; n._RightChild = u[2];
        g_nodes[i] = n;
    }
    shader::uTriangles = g_tris.data();
    shader::uModels = g_models.data();
    shader::uMaterials = g_materials.data();
    shader::uBVH_Nodes = g_nodes.data();
    shader::uNbTriangles = nb_tris;      // scene.cpp:215-217
    shader::uNbMaterials = nb_materials;
    shader::uNbModels = nb_models;
    return 0;
}

// CameraGPU (camera.hpp:21-30) as application.cpp:236-243 uploads it: 4 column-major mat4, eye, 3 floats = 284 B
void ref_rt_set_camera(const float* cam) {
    shader::mat4* ms[4] = {&shader::uCamera._View, &shader::uCamera._Proj, &shader::uCamera._InvView, &shader::uCamera._InvProj};
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 4; ++c)
            (*ms[k])[c] = shader::vec4(cam[16 * k + 4 * c], cam[16 * k + 4 * c + 1], cam[16 * k + 4 * c + 2], cam[16 * k + 4 * c + 3]);
    shader::uCamera._Eye = shader::vec4(cam[64], cam[65], cam[66], cam[67]);
    shader::uCamera._PlaneWidth = cam[68];
    shader::uCamera._PlaneHeight = cam[69];
    shader::uCamera._PlaneNear = cam[70];
}

// application.cpp:229-232
void ref_rt_set_flags(int depth_display_bvh, int is_bvh_displayed, int is_wireframe_mode_on) {
    shader::uDepthDisplayBVH = depth_display_bvh;
    shader::uIsBVHDisplayed = is_bvh_displayed != 0;
    shader::uIsWireframeModeOn = is_wireframe_mode_on != 0;
}

// getRay(pos), raytracer.glsl:92-100
void ref_rt_get_ray(float pos_x, float pos_y, HostRay* out) {
    shader::Ray r = shader::getRay(shader::vec2(pos_x, pos_y));
    for (int k = 0; k < 4; ++k) { out->o[k] = r._Origin[k]; out->d[k] = r._Direction[k]; }
}

// The closest hits of a whole frame: pos as main() computes it (:303-305 restated here in two lines -- ref_rt_dispatch
// below runs main() itself and pins them), then getRay and getClosestHitBVH verbatim.  Pixels outside
// [0, denom_w) x [0, denom_h) have no invocation (Q5) and keep a zero record.  rays_out is nullable.
void ref_rt_trace_primary(uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h, HostHit* hits_out,
                          HostRay* rays_out) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads())
    for (int64_t y = 0; y < (int64_t)height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            const size_t i = (size_t)y * width + x;
            std::memset(&hits_out[i], 0, sizeof(HostHit));
            if (rays_out) std::memset(&rays_out[i], 0, sizeof(HostRay));
            if (x >= denom_w || (uint32_t)y >= denom_h) continue;
            shader::vec2 pos(float(int(x)) / denom_w, float(int(y)) / denom_h);
            shader::Ray r = shader::getRay(pos);
            shader::vec4 c(0.f, 0.f, 0.f, 0.f);
            from_hit(shader::getClosestHitBVH(r, 0u, c), hits_out[i]);
            if (rays_out) for (int k = 0; k < 4; ++k) { rays_out[i].o[k] = r._Origin[k]; rays_out[i].d[k] = r._Direction[k]; }
        }
}

// rayTriangleIntersection, :102-147
void ref_rt_ray_triangle(const HostRay* rays, const uint32_t* tri_index, uint64_t n, HostHit* out) {
#pragma omp parallel for schedule(static) num_threads(threads())
    for (int64_t i = 0; i < (int64_t)n; ++i)
        from_hit(shader::rayTriangleIntersection(to_ray(rays[i]), tri_index[i]), out[i]);
}

// intersectBVH, :182-237 (0 / 1 / 2) on node node_index[i] of the bound node array
void ref_rt_intersect_bvh(const HostRay* rays, const uint32_t* node_index, uint64_t n, uint32_t* out) {
#pragma omp parallel for schedule(static) num_threads(threads())
    for (int64_t i = 0; i < (int64_t)n; ++i)
        out[i] = shader::intersectBVH(to_ray(rays[i]), shader::uBVH_Nodes[node_index[i]]);
}

// getClosestHitBVH(ray, 0, bvhColor), :246-295; bvh_color_out (nullable): 4 floats per ray, starting from (0,0,0,0) (:321)
void ref_rt_closest_hit_bvh(const HostRay* rays, uint64_t n, HostHit* out, float* bvh_color_out) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads())
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        shader::vec4 c(0.f, 0.f, 0.f, 0.f);
        from_hit(shader::getClosestHitBVH(to_ray(rays[i]), 0u, c), out[i]);
        if (bvh_color_out) for (int k = 0; k < 4; ++k) bvh_color_out[4 * i + k] = c[k];
    }
}

// getAllHits, :149-157 -- the shader's own brute-force path (disabled at :309-312), no undefined reads
void ref_rt_all_hits(const HostRay* rays, uint64_t n, HostHit* out) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads())
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        shader::Hit closest = RTR_UNDEFINED_HIT;
        closest._DidHit = 0;  // :308
        shader::getAllHits(to_ray(rays[i]), shader::uNbTriangles, closest);
        from_hit(closest, out[i]);
    }
}

// getColor, :159-179
void ref_rt_get_color(const HostHit* hits, const float* bvh_color, uint64_t n, float* rgba_out) {
    for (uint64_t i = 0; i < n; ++i) {
        shader::Hit h = RTR_UNDEFINED_HIT;
        h._Coords = shader::vec4(hits[i].b0, hits[i].b1, hits[i].b2, hits[i].t);
        h._DidHit = hits[i].did_hit; h._TriangleId = hits[i].tri;
        shader::vec4 value(0.f, 0.f, 0.f, 1.f);  // :300
        shader::vec4 bc(0.f, 0.f, 0.f, 0.f);
        if (bvh_color) bc = shader::vec4(bvh_color[4 * i], bvh_color[4 * i + 1], bvh_color[4 * i + 2], bvh_color[4 * i + 3]);
        shader::getColor(h, bc, value);
        for (int k = 0; k < 4; ++k) rgba_out[4 * i + k] = value[k];
    }
}

// glDispatchCompute(groups_x, groups_y, 1) of application.cpp:245 on a width x height rgba32f image (:335,:341);
// the application passes floor(W/16), floor(H/16) (:225-226).  rgba_out is left untouched where no invocation stores.
void ref_rt_dispatch(uint32_t groups_x, uint32_t groups_y, uint32_t width, uint32_t height, float* rgba_out) {
    shader::oImage.rgba = rgba_out;
    shader::oImage.width = width;
    shader::oImage.height = height;
    shader::gl_NumWorkGroups = shader::uvec3(groups_x, groups_y, 1);
    const int64_t ny = 16 * (int64_t)groups_y, nx = 16 * (int64_t)groups_x;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads())
    for (int64_t y = 0; y < ny; ++y)
        for (int64_t x = 0; x < nx; ++x) {
            shader::gl_GlobalInvocationID = shader::uvec3((uint32_t)x, (uint32_t)y, 0);
            shader::shader_main();
        }
}

}  // extern "C"
