/*
 * rtr_oracle.h -- CPU restatement of the reference's acceleration-structure and
 * ray-cast path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (librtr_b200.so) never
 * links, loads or calls it; it has no CPU fallback.
 *
 * Parity status: the build half (Morton codes, pair sort, PLOC topology,
 * flatten) is PINNED against the reference's own bvh.cpp/triangle.cpp compiled
 * unmodified into oracle/_ref/ (see oracle/Makefile, tests/test_oracle_golden_cpu.py)
 * and against the committed fixtures in tests/golden/ generated from that build.
 * The sort pre-pass helpers are pinned against the known-answer vectors of the
 * reference's tests/testsSortGPU.  The traversal half (ray generation, slab test,
 * ray/triangle, stack traversal, getAllHits, getColor, the depth overlay, main's
 * pixel mapping) is PINNED against the reference's own raytracer.glsl compiled as
 * C++ through the reference's vendored GLM (oracle/ref_raytracer.cpp + glsl2cpp.sed
 * -> oracle/_ref/libref_raytracer_*.so) and against tests/golden/raytracer_golden.npz
 * generated from that library: bit for bit with normalize := v / sqrt(dot(v, v))
 * (tests/test_raytracer_pin_cpu.py), within 1e-5 relative with glm::normalize.
 * Documented deviations: Q6 (guarded misses, as the shader's own getAllHits) and Q11
 * (IEEE fminf/fmaxf where a zero direction component meets a slab plane: NaN).
 * Secondary rays (orc_render) do not exist in the reference and are defined here.
 *
 * All citations are relative to /root/reference.
 */
#ifndef RTR_ORACLE_H
#define RTR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* srcCommon/scene/geometry/triangle.hpp:9-14 -- 64 B, _ModelId at 48 */
typedef struct {
    float p0[4];
    float p1[4];
    float p2[4];
    uint32_t model_id;
    uint32_t pad[3];
} orc_triangle;

/* srcCommon/scene/geometry/mesh.hpp:12-15 -- 68 B, column-major mat4 */
typedef struct {
    float m[16];
    uint32_t material_id;
} orc_mesh;

/* srcCommon/scene/geometry/bvh.hpp:22-42 -- 48 B */
typedef struct {
    float bmin[3];
    uint32_t pad0;
    float bmax[3];
    uint32_t pad1;
    uint32_t triangle_id;
    uint32_t left;
    uint32_t right;
    uint32_t pad2;
} orc_node;

/* srcCommon/scene/camera.hpp:21-30 (column-major mat4s) */
typedef struct {
    float view[16];
    float proj[16];
    float inv_view[16];
    float inv_proj[16];
    float eye[4];
    float plane_width;
    float plane_height;
    float plane_near;
} orc_camera;

/* raytracer.glsl:14-17 */
typedef struct {
    float origin[4];
    float direction[4];
} orc_ray;

/* raytracer.glsl:36-40 ; 24 B */
typedef struct {
    float b0, b1, b2, t;
    uint32_t did_hit;
    uint32_t triangle_id;
} orc_hit;

/* ---- sort pre-passes pinned by tests/testsSortGPU ---- */
void orc_bit_histogram32(const uint32_t* keys, uint32_t n, uint32_t out[32]);
void orc_digitplace_exclusive_scan(const uint32_t in[32], uint32_t out[32]);

/* ---- Morton codes (bvh.cpp:235-372) ---- */
void orc_scene_aabb(const orc_triangle* tris, uint32_t array_len,
                    const orc_mesh* meshes, float out_minmax[6]);
void orc_circumscribed_cube(const float scene_minmax[6], float cube_minmax[6]);
void orc_morton_codes(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                      const orc_mesh* meshes, uint32_t* codes_out);

/* ---- sort (bvh.cpp:214-232): std::sort on pair<code,index> ---- */
void orc_sort_pairs(uint32_t* codes, uint32_t* indices, uint32_t n); /* comparison sort */
void orc_radix_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t n); /* stable LSD radix-256 */
void orc_radix_sort_pairs_mt(uint32_t* keys, uint32_t* vals, uint32_t n, int threads); /* the same on `threads` cores (<= 0: all) */
void orc_radix_sort_keys_u32(uint32_t* keys, uint32_t n);
void orc_radix_sort_keys_u64(uint64_t* keys, uint32_t n);
void orc_radix_sort_pairs_u64(uint64_t* keys, uint32_t* vals, uint32_t n);
void orc_morton_codes64(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                        const orc_mesh* meshes, uint64_t* codes_out);

/* ---- PLOC build (bvh.cpp:26-210) ---- */
typedef struct {
    uint32_t n;            /* triangles */
    uint32_t nb_clusters;  /* 2n-1 */
    uint32_t* morton_sorted;     /* [n] */
    uint64_t* morton_sorted64;   /* [n], orc_bvh_build64 only (else NULL) */
    uint32_t* triangle_indices;  /* [n] */
    orc_node* clusters;          /* [2n-1] by cluster id; internal links are 0 (Q10) */
    uint32_t* parent;            /* [2n-1], 0xFFFFFFFF = none */
    uint32_t* left;              /* [2n-1], 0xFFFFFFFF for leaves */
    uint32_t* right;             /* [2n-1] */
    uint32_t nb_iterations;
    uint32_t* trace_active;      /* [nb_iterations] n_i */
    uint32_t* trace_merges;      /* [nb_iterations] m_i */
} orc_bvh;

/* returns NULL on bad arguments */
orc_bvh* orc_bvh_build(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                       const orc_mesh* meshes, uint32_t nb_meshes,
                       uint32_t search_radius);
/* same builder, leaves ordered by 63-bit Morton keys (no reference definition, see rtr_oracle.c) */
orc_bvh* orc_bvh_build64(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                         const orc_mesh* meshes, uint32_t nb_meshes,
                         uint32_t search_radius);
const uint64_t* orc_bvh_morton_sorted64(const orc_bvh* b);
void orc_bvh_destroy(orc_bvh* b);
/* plain accessors for ctypes users */
uint32_t orc_bvh_nb_iterations(const orc_bvh* b);
const uint32_t* orc_bvh_morton_sorted(const orc_bvh* b);
const uint32_t* orc_bvh_triangle_indices(const orc_bvh* b);
const orc_node* orc_bvh_clusters(const orc_bvh* b);
const uint32_t* orc_bvh_parent(const orc_bvh* b);
const uint32_t* orc_bvh_left(const orc_bvh* b);
const uint32_t* orc_bvh_right(const orc_bvh* b);
const uint32_t* orc_bvh_trace_active(const orc_bvh* b);
const uint32_t* orc_bvh_trace_merges(const orc_bvh* b);

/* ---- flatten (srcOpenGL/scene/scene.cpp:189-208) ---- */
void orc_flatten(const orc_node* clusters, const uint32_t* left, const uint32_t* right,
                 uint32_t n, orc_node* flat_out /* [2n-1] */);

/* FNV-1a-64 variant over 32-bit words, SURVEY App. C */
uint64_t orc_hash_words(const uint32_t* w, uint64_t nwords);
uint64_t orc_hash_flat_nodes(const orc_node* flat, uint32_t nb_nodes);

/* ---- traversal (raytracer.glsl:92-147,182-295,299-331) ---- */
void orc_get_ray(const orc_camera* cam, uint32_t x, uint32_t y,
                 uint32_t denom_w, uint32_t denom_h, orc_ray* out);
void orc_get_rays(const orc_camera* cam, uint32_t width, uint32_t height,
                  uint32_t denom_w, uint32_t denom_h, orc_ray* out /* [height*width] */);
void orc_ray_triangle(const orc_ray* ray, const orc_triangle* tris,
                      const orc_mesh* meshes, uint32_t tri_index, orc_hit* out);
uint32_t orc_intersect_box(const orc_ray* ray, const orc_node* node);
/* the same with the shader's return code 2 (entry point near a box edge, :222-233; threshold 0.05 / (display_depth + 1)) */
uint32_t orc_intersect_box_edge(const orc_ray* ray, const orc_node* node, int display_depth);
void orc_closest_hit_bvh(const orc_ray* ray, const orc_node* flat,
                         const orc_triangle* tris, const orc_mesh* meshes,
                         orc_hit* out, uint64_t* nodes_visited);
void orc_closest_hit_brute(const orc_ray* ray, const orc_triangle* tris, uint32_t n,
                           const orc_mesh* meshes, orc_hit* out);
int orc_any_hit_bvh(const orc_ray* ray, float t_max, const orc_node* flat,
                    const orc_triangle* tris, const orc_mesh* meshes);

/* whole-image / batch drivers; threads<=0 -> all cores (OpenMP) */
void orc_trace_primary(const orc_node* flat, const orc_triangle* tris,
                       const orc_mesh* meshes, const orc_camera* cam,
                       uint32_t width, uint32_t height,
                       uint32_t denom_w, uint32_t denom_h,
                       orc_hit* hits_out /* [height*width] */, int threads);
void orc_trace_rays(const orc_node* flat, const orc_triangle* tris,
                    const orc_mesh* meshes, const orc_ray* rays, uint64_t n_rays,
                    orc_hit* hits_out, int threads);
/* rows [row0,row1) only; hits_out/rgba_out index from row0 */
void orc_render(const orc_node* flat, const orc_triangle* tris,
                const orc_mesh* meshes, const orc_camera* cam,
                uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h,
                uint32_t row0, uint32_t row1, uint32_t bounces, int shadow,
                const float light_pos[3],
                float* rgba_out, orc_hit* primary_hits_out, uint64_t* rays_traced,
                int threads);
/* getColor + main of raytracer.glsl (:159-179, :299-331) without the BVH overlay: the reference's actual frame.
 * value = (0,0,0,1); if the pixel hit: value += colour of the material of the triangle's model; wireframe: a hit
 * with a barycentric below WIREFRAME_LINE_WIDTH (0.02, :71) is (0,0,0,1).  materials: 4 floats each (material.hpp:9-11).
 * The material id is the host's meshes[model]._MaterialId (the shader reads it at stride 80 instead of the host's 68,
 * Q7: identical for model 0, i.e. for every single-model scene). */
void orc_shade(const orc_hit* hits, uint64_t n, const orc_triangle* tris, const orc_mesh* meshes,
               const float* materials, int wireframe, const float* bvh_rgba /* nullable: uIsBVHDisplayed */,
               float* rgba_out);
/* the bvhColor of getClosestHitBVH (raytracer.glsl:246-295, :222-233): colour of the last visited intersected node at
 * depth display_depth -- box (0.5,0,0.5,0.1), or line (0.7,0,0.7,0.1) where the ray enters near a box edge */
void orc_depth_overlay(const orc_node* flat, const orc_camera* cam, uint32_t width, uint32_t height,
                       uint32_t denom_w, uint32_t denom_h, int display_depth, float* out);
/* deterministic secondary-ray definitions (shared with the CUDA path) */
int orc_bounce_ray(const orc_ray* in, const orc_hit* hit, const orc_triangle* tris,
                   const orc_mesh* meshes, orc_ray* out);
int orc_shadow_ray(const orc_ray* in, const orc_hit* hit, const orc_triangle* tris,
                   const orc_mesh* meshes, const float light_pos[3],
                   orc_ray* out, float* t_max);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
