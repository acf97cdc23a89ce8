/*
 * rtr.h -- C ABI of librtr_b200.so: the B200-native (sm_100a) replacement for the
 * acceleration-structure + ray-cast path of MrBigoudi/RealTimeRaytracing.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * reference checkout).  Conventions:
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every call returns int: 0 = RTR_OK, negative = RTR_E_*; rtr_last_error() gives text;
 *     nothing exits or throws across the ABI (the reference prints and exit()s,
 *     srcCommon/core/errorHandler.cpp:13-29 -- the C++ shim in rtr_scene.hpp restores that);
 *   - inputs are borrowed for the duration of the call; outputs go to caller-owned buffers;
 *     opaque rtr_ctx / rtr_bvh own device memory, destroyed explicitly;
 *   - one rtr_ctx per (host thread, GPU); calls on one ctx are not re-entrant;
 *   - names ending in _dev take DEVICE pointers and are asynchronous on the ctx stream;
 *     the others take HOST pointers, copy in/out and return after completion;
 *   - there is NO CPU fallback: without a usable CUDA device rtr_ctx_create fails with
 *     RTR_E_NODEVICE and nothing else can be called.
 */
#ifndef RTR_H
#define RTR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTR_OK 0
#define RTR_E_INVALID (-1)     /* bad argument */
#define RTR_E_CUDA (-2)        /* CUDA runtime error, see rtr_last_error */
#define RTR_E_NOMEM (-3)       /* device or pinned-host allocation failed */
#define RTR_E_NODEVICE (-4)    /* no CUDA device / wrong architecture */
#define RTR_E_UNSUPPORTED (-5) /* size outside the supported range */
#define RTR_E_STATE (-6)       /* object not in the state the call needs */
#define RTR_E_COMM (-7)        /* NCCL error / libnccl not loadable */

#define RTR_DEFAULT_SEARCH_RADIUS 16u /* PlocParams::_SEARCH_RADIUS, bvh.hpp:66 */
#define RTR_MAX_SEARCH_RADIUS 16u
#define RTR_NONE 0xFFFFFFFFu

/* ---- reference record layouts (SURVEY.md App. A rule 10) ---- */

/* cr::TriangleGPU, srcCommon/scene/geometry/triangle.hpp:9-14 (64 B, _ModelId at 48) */
typedef struct rtr_triangle {
    float p0[4];
    float p1[4];
    float p2[4];
    uint32_t model_id;
    uint32_t pad[3];
} rtr_triangle;

/* cr::MeshModelGPU, srcCommon/scene/geometry/mesh.hpp:12-15 (68 B host layout, Q7) */
typedef struct rtr_mesh {
    float model[16]; /* column-major */
    uint32_t material_id;
} rtr_mesh;

/* cr::BVH_NodeGPU, srcCommon/scene/geometry/bvh.hpp:22-42 (48 B); leaf <=> left==0 && right==0 */
typedef struct rtr_node {
    float bmin[3];
    uint32_t pad0;
    float bmax[3];
    uint32_t pad1;
    uint32_t triangle_id;
    uint32_t left;
    uint32_t right;
    uint32_t pad2;
} rtr_node;

/* cr::CameraGPU, srcCommon/scene/camera.hpp:21-30 */
typedef struct rtr_camera {
    float view[16];
    float proj[16];
    float inv_view[16];
    float inv_proj[16];
    float eye[4];
    float plane_width;
    float plane_height;
    float plane_near;
} rtr_camera;

/* Ray, srcCommon/shaders/raytracer.glsl:14-17 */
typedef struct rtr_ray {
    float origin[4];
    float direction[4];
} rtr_ray;

/* Hit, srcCommon/shaders/raytracer.glsl:36-40 ; (b0,b1,b2,t), DidHit, TriangleId */
typedef struct rtr_hit {
    float b0, b1, b2, t;
    uint32_t did_hit;
    uint32_t triangle_id;
} rtr_hit;

typedef struct rtr_ctx rtr_ctx;
typedef struct rtr_bvh rtr_bvh;

/* ---- context ---- */
const char* rtr_version(void);
/* replaces glr::Application::dummyApplication() (srcOpenGL/application.cpp:416-423) as the
 * "give me a device context" step of the tests and of Application::init (:281-295) */
int rtr_ctx_create(int device, rtr_ctx** out);
int rtr_ctx_destroy(rtr_ctx* ctx);
/* ctx may be NULL: returns the last error of a failed rtr_ctx_create on this thread */
const char* rtr_last_error(const rtr_ctx* ctx);
int rtr_ctx_sync(rtr_ctx* ctx);
void* rtr_ctx_stream(rtr_ctx* ctx);                 /* cudaStream_t the ctx launches on */
int rtr_ctx_set_stream(rtr_ctx* ctx, void* stream); /* adopt a caller-owned cudaStream_t */
/* Pipelined frames: later calls on this ctx are enqueued on `stream` (borrowed); unlike rtr_ctx_set_stream the
 * work already enqueued is not waited for, so a caller can alternate between two streams -- rebuild + broadcast
 * of frame f+1 on one, rays of frame f on the other -- and order them with its own events
 * (SURVEY.md 8b "build + trace may overlap on different streams"). */
int rtr_ctx_switch_stream(rtr_ctx* ctx, void* cuda_stream);
/* the persistent traversal kernel leaves `sms` SMs free, e.g. for the NCCL kernels of a broadcast in flight */
int rtr_ctx_reserve_sms(rtr_ctx* ctx, uint32_t sms);
/* An SM partition for the rays (a CUDA green context): `sms` SMs, rounded up to the granularity the architecture
 * partitions by (8 on sm_90+; *sms_out gets the count), with n_streams (<= 8) streams of its own in streams_out.
 * Traversal launches on those streams (rtr_ctx_switch_stream) fill the partition only, so several frames' rays can be
 * in flight at once and the SMs outside it stay free for a concurrent collective.  Once per ctx. */
int rtr_ctx_partition_sms(rtr_ctx* ctx, uint32_t sms, uint32_t n_streams, void** streams_out, uint32_t* sms_out);
int rtr_ctx_device(const rtr_ctx* ctx);
int rtr_ctx_sm_count(const rtr_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py reports it as gpu_launches) */
uint64_t rtr_ctx_launch_count(const rtr_ctx* ctx);

/* per-kernel device times: while enabled, the launches of the kernels that dominate the path are
 * bracketed by CUDA events; read returns one line per kernel name in `names` ('\n' separated) with its
 * summed milliseconds and launch count since enable.  Replaces the reference's commented-out
 * glfwGetTime phase timers (bvh.cpp:21-23,65-103). */
int rtr_ctx_profile_enable(rtr_ctx* ctx, int enable);
int rtr_ctx_profile_read(rtr_ctx* ctx, char* names, size_t names_bytes, float* total_ms, uint32_t* counts,
                         uint32_t capacity, uint32_t* n_out);

/* pinned host buffers for the host-pointer entry points (plain malloc'd memory also works) */
int rtr_host_alloc(size_t bytes, void** out);
int rtr_host_free(void* p);
/* page-lock / release host memory the caller owns (a mapping shared by the processes of a multi-GPU job, for
 * rtr_download_stripes_async) */
int rtr_host_register(void* p, size_t bytes);
int rtr_host_unregister(void* p);
/* device buffers for C/C++ hosts (replace glCreateBuffers/glNamedBufferStorage,
 * tests/testsSortGPU/testHistogramCreation.cpp:75-85, and glGetNamedBufferSubData :131) */
int rtr_dev_alloc(rtr_ctx* ctx, size_t bytes, void** out);
int rtr_dev_free(rtr_ctx* ctx, void* p);
int rtr_dev_upload(rtr_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int rtr_dev_download(rtr_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
/* the same, enqueued on the ctx stream without waiting (the host buffer must stay valid, and be pinned -- rtr_host_alloc --
 * for the copy to run beside kernels): with rtr_ctx_switch_stream the triangles of frame f+1 can travel while frame f
 * is traced */
int rtr_dev_upload_async(rtr_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int rtr_dev_download_async(rtr_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int rtr_dev_zero(rtr_ctx* ctx, void* dst_dev, size_t bytes);

/* ---- sort pre-passes pinned by tests/testsSortGPU ---- */
/* srcCommon/shaders/ploc/preprocessing/sort/histogramOfGlobalDigitCounts.glsl:31-59 :
 * out[b] = number of keys with bit b set.  The _dev form ADDS into out_dev (the shader
 * atomically accumulates into a pre-zeroed SSBO, testHistogramCreation.cpp:113-116). */
int rtr_bit_histogram32(rtr_ctx* ctx, const uint32_t* keys, uint32_t n, uint32_t out[32]);
int rtr_bit_histogram32_dev(rtr_ctx* ctx, const uint32_t* keys_dev, uint32_t n, uint32_t* out_dev);
/* .../sort/prefixSumOfGlobalDigitCounts.glsl:24-69 : exclusive scan inside each group of 4 bins */
int rtr_digitplace_exclusive_scan(rtr_ctx* ctx, const uint32_t in[32], uint32_t out[32]);
int rtr_digitplace_exclusive_scan_dev(rtr_ctx* ctx, const uint32_t* in_dev, uint32_t* out_dev);

/* ---- LSD radix sort (Onesweep).  Replaces the std::sort of (code,index) pairs at
 * srcCommon/scene/geometry/bvh.cpp:223-231 and the unfinished chainedScanDigitBinning.glsl.
 * Stable; ascending; in place from the caller's point of view.  n < 2^30. ---- */
int rtr_sort_keys_u32(rtr_ctx* ctx, uint32_t* keys, uint32_t n);
int rtr_sort_pairs_u32(rtr_ctx* ctx, uint32_t* keys, uint32_t* values, uint32_t n);
int rtr_sort_keys_u64(rtr_ctx* ctx, uint64_t* keys, uint32_t n);
int rtr_sort_pairs_u64(rtr_ctx* ctx, uint64_t* keys, uint32_t* values, uint32_t n);
/* device forms: values_dev may be NULL (keys only); bits [begin_bit, end_bit) take part */
int rtr_sort_pairs_u32_dev(rtr_ctx* ctx, uint32_t* keys_dev, uint32_t* values_dev, uint32_t n,
                           int begin_bit, int end_bit);
int rtr_sort_pairs_u64_dev(rtr_ctx* ctx, uint64_t* keys_dev, uint32_t* values_dev, uint32_t n,
                           int begin_bit, int end_bit);

/* ---- Morton codes: BVH::getMortonCodes, bvh.cpp:330-348 (:235-328, :350-372) and the
 * stub shader ploc/preprocessing/preprocessing.glsl.  tris_array_len is the length of the
 * vector the reference computes the scene box over (Q2); codes are in input order. ---- */
int rtr_morton_codes(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t tris_array_len,
                     const rtr_mesh* meshes, uint32_t nb_meshes, uint32_t* codes_out);
int rtr_morton_codes_dev(rtr_ctx* ctx, const rtr_triangle* tris_dev, uint32_t n, uint32_t tris_array_len,
                         const rtr_mesh* meshes_dev, uint32_t nb_meshes, uint32_t* codes_dev);
/* 21-bit-per-axis extension (no reference definition) */
int rtr_morton_codes64(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t tris_array_len,
                       const rtr_mesh* meshes, uint32_t nb_meshes, uint64_t* codes_out);
/* out[0..5] = scene AABB min,max (bvh.cpp:235-251); out[6..11] = "circumscribed cube" (bvh.cpp:253-303) */
int rtr_scene_bounds(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t tris_array_len,
                     const rtr_mesh* meshes, uint32_t nb_meshes, float out[12]);

/* ---- BVH build: replaces cr::BVH::BVH(nbTriangles, unsortedTriangles, meshesInTheScene)
 * (bvh.hpp:89-91, bvh.cpp:11-24) as called from glr::Scene::bindSSBO (scene.cpp:148) plus the
 * flatten glr::Scene::getBVH_NodesToGPUData (scene.cpp:189-208). ---- */
/* host inputs: copies them to the device (as the reference ctor copies its vectors) and
 * returns after the build finished.  *out must be NULL or a BVH from this ctx to reuse. */
int rtr_bvh_build(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t nb_triangles, uint32_t tris_array_len,
                  const rtr_mesh* meshes, uint32_t nb_meshes, uint32_t search_radius, rtr_bvh** out);
/* device inputs, borrowed (must stay alive while the BVH is traced); asynchronous */
int rtr_bvh_build_dev(rtr_ctx* ctx, const rtr_triangle* tris_dev, uint32_t nb_triangles, uint32_t tris_array_len,
                      const rtr_mesh* meshes_dev, uint32_t nb_meshes, uint32_t search_radius, rtr_bvh** out);
/* the same builder with the leaves ordered by 63-bit Morton keys (21 bits per axis, rtr_morton_codes64; 8 sort passes
 * over (u64 key, u32 index) pairs).  The reference has 30-bit codes only (bvh.cpp:358-372): this is the "64-bit keys"
 * variant of north_star, defined by oracle/rtr_oracle.c: orc_bvh_build64; its leaf order refines the 32-bit one
 * (code64 >> 33 == code32) and only differs where 10 bits per axis cannot tell centroids apart. */
int rtr_bvh_build64(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t nb_triangles, uint32_t tris_array_len,
                    const rtr_mesh* meshes, uint32_t nb_meshes, uint32_t search_radius, rtr_bvh** out);
int rtr_bvh_build64_dev(rtr_ctx* ctx, const rtr_triangle* tris_dev, uint32_t nb_triangles, uint32_t tris_array_len,
                        const rtr_mesh* meshes_dev, uint32_t nb_meshes, uint32_t search_radius, rtr_bvh** out);
int rtr_bvh_destroy(rtr_bvh* bvh);
uint32_t rtr_bvh_nb_triangles(const rtr_bvh* bvh);
uint32_t rtr_bvh_nb_nodes(const rtr_bvh* bvh); /* 2n-1 */
/* PLOC iterations of the last build and its trace (n_i active clusters, m_i merges) */
int rtr_bvh_iteration_trace(rtr_bvh* bvh, uint32_t* active, uint32_t* merges, uint32_t capacity, uint32_t* count);
/* diagnostics: device clock (%globaltimer, ns) at the start of every PLOC iteration of the last build; count + 1
 * values, the last one is the end of the loop.  The loop runs on the device without the host (see rtr_bvh_build_dev). */
int rtr_bvh_iteration_times(rtr_bvh* bvh, uint64_t* start_ns, uint32_t capacity, uint32_t* count);
/* per-stage device times of the last build in ms: [0] scene box + Morton, [1] sort, [2] leaf init,
 * [3] PLOC loop, [4] flatten, [5] total.  Only measured when enabled. */
int rtr_bvh_enable_stage_timing(rtr_bvh* bvh, int enable);
int rtr_bvh_stage_ms(rtr_bvh* bvh, float out[6]);
/* sorted Morton codes (PlocParams::_MortonCodes) and BVH_Params::_TriangleIndices (bvh.hpp:55) */
int rtr_bvh_morton_codes(rtr_bvh* bvh, uint32_t* out);
int rtr_bvh_morton_codes64(rtr_bvh* bvh, uint64_t* out); /* sorted 63-bit codes of a rtr_bvh_build64 BVH */
int rtr_bvh_triangle_indices(rtr_bvh* bvh, uint32_t* out);
/* BVH_Params::_Clusters/_Parent/_LeftChild/_RightChild/_IsLeaf (bvh.hpp:50-54) by cluster id,
 * ids in serial-merge order (Q3); absent links are RTR_NONE; any pointer may be NULL */
int rtr_bvh_clusters(rtr_bvh* bvh, rtr_node* clusters, uint32_t* parent, uint32_t* left, uint32_t* right,
                     uint8_t* is_leaf);
/* DFS pre-order array of scene.cpp:203-208 */
int rtr_bvh_flat_nodes(rtr_bvh* bvh, rtr_node* out);
const rtr_node* rtr_bvh_device_nodes(const rtr_bvh* bvh);
const rtr_triangle* rtr_bvh_device_triangles(const rtr_bvh* bvh);
const rtr_mesh* rtr_bvh_device_meshes(const rtr_bvh* bvh);
/* wrap device arrays produced elsewhere (e.g. received by a broadcast) as a traceable BVH; borrowed */
int rtr_bvh_adopt_dev(rtr_ctx* ctx, const rtr_node* nodes_dev, uint32_t nb_triangles,
                      const rtr_triangle* tris_dev, const rtr_mesh* meshes_dev, uint32_t nb_meshes,
                      rtr_bvh** out);

/* ---- traversal: replaces glDispatchCompute on raytracer.glsl (srcOpenGL/application.cpp:245).
 * denom_w/denom_h are the pixel-position divisors numGroups*16 of raytracer.glsl:304-305 (Q5);
 * pass 0 for the reference formula floor(W/16)*16.  Pixels at or beyond the divisors are not
 * traced (their records are zero).  Rows [row0,row1) are produced (row1 = 0 means height). ---- */
#define RTR_TRACE_DEFAULT 0u
/* visit nodes exactly as getClosestHitBVH does (raytracer.glsl:246-295): no pruning, right child
 * first.  The default order prunes by the closest hit found so far with a conservative margin
 * and returns identical results. */
#define RTR_TRACE_REFERENCE_ORDER 1u
/* the shader's loop with the shader's own stack: `const uint STACK_SIZE = 1024` (raytracer.glsl:251) where the other
 * kernels keep 128 entries per lane.  Same records, slower (the stack lives in local memory).  The host-pointer calls
 * fall back to it by themselves when a ray runs out of the 128 entries; see rtr_bvh_stack_overflows. */
#define RTR_TRACE_DEEP_STACK 2u

int rtr_trace_primary(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_camera* camera,
                      uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h,
                      uint32_t flags, rtr_hit* hits_out /* [height*width] */);
int rtr_trace_primary_dev(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_camera* camera,
                          uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h,
                          uint32_t row0, uint32_t row1, uint32_t flags, rtr_hit* hits_dev /* [(row1-row0)*width] */);
/* explicit ray batches (getClosestHitBVH semantics); any_hit != 0: did_hit = 1 iff a hit with
 * t < t_max[i] exists (t_max NULL = +inf), other fields zero.
 * Directions: the shader's rays are unit vectors and the default order's pruning margin is derived for them; it holds
 * with room up to |direction| = 8.  A batch of rtr_trace_rays that holds a longer direction is traced in the shader's
 * own order instead (RTR_TRACE_REFERENCE_ORDER, which takes any length -- same records, no pruning); the caller of
 * rtr_trace_rays_dev is responsible for the same choice. */
int rtr_trace_rays(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_ray* rays, uint64_t n_rays, int any_hit,
                   const float* t_max, uint32_t flags, rtr_hit* hits_out);
int rtr_trace_rays_dev(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_ray* rays_dev, uint64_t n_rays, int any_hit,
                       const float* t_max_dev, uint32_t flags, rtr_hit* hits_dev);
/* multi-bounce frame (definition in DESIGN.md "secondary rays"): rgba32f like the reference's
 * image unit 0 (srcOpenGL/application.cpp:335,341).  rgba / primary_hits / rays_traced may be NULL. */
int rtr_render(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_camera* camera,
               uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h,
               uint32_t row0, uint32_t row1, uint32_t bounces, int shadow, const float light_pos[3],
               uint32_t flags, float* rgba_out, rtr_hit* primary_hits_out, uint64_t* rays_traced);
int rtr_render_dev(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_camera* camera,
                   uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h,
                   uint32_t row0, uint32_t row1, uint32_t bounces, int shadow, const float light_pos[3],
                   uint32_t flags, float* rgba_dev, rtr_hit* primary_hits_dev, uint64_t* rays_traced_dev);

/* the same frame, but only the row blocks dealt to shard_rank (block b of rows_per_block rows belongs
 * to rank b % shard_count); rgba_dev / primary_hits_dev are FULL-image buffers written at global rows */
int rtr_render_sharded_dev(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_camera* camera,
                           uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h,
                           uint32_t rows_per_block, uint32_t shard_rank, uint32_t shard_count,
                           uint32_t bounces, int shadow, const float light_pos[3], uint32_t flags,
                           float* rgba_dev, rtr_hit* primary_hits_dev, uint64_t* rays_traced_dev);

/* The shader's private stack is 1024 entries deep (raytracer.glsl:251), the fast kernels' 128.  A ray that runs out of
 * it drops subtrees; the kernels count such rays.  rtr_trace_primary / rtr_trace_rays / rtr_render (host pointers)
 * then trace the call again with RTR_TRACE_DEEP_STACK -- the shader's loop with the shader's stack -- and fail with
 * RTR_E_UNSUPPORTED only if a ray exhausts those 1024 entries too (where the shader itself writes out of bounds);
 * after the asynchronous _dev forms, this call waits for the ctx stream and returns the count since the last call
 * (and clears it), and the caller may repeat the launch with RTR_TRACE_DEEP_STACK.  0 on every scene measured so far. */
int rtr_bvh_stack_overflows(const rtr_bvh* bvh, uint32_t* count_out);

/* ---- shading: getColor + main of raytracer.glsl (:159-179, :299-331), the reference's rgba32f frame from the hit
 * records of rtr_trace_primary.  value = (0,0,0,1) -- or, with RTR_SHADE_BVH (uIsBVHDisplayed), the pixel's colour
 * from rtr_bvh_depth_overlay; a hit adds the colour of the material of the triangle's model; with
 * RTR_SHADE_WIREFRAME a hit whose barycentric falls below WIREFRAME_LINE_WIDTH = 0.02 (:71, :170-178) is (0,0,0,1).
 * The material id is the host's meshes[model]._MaterialId (the shader reads it at stride 80 where the host wrote
 * stride 68, SURVEY Q7: the same value for model 0, i.e. for every single-model scene). */
/* cr::MaterialGPU, srcCommon/scene/pbr/material.hpp:9-11 */
typedef struct rtr_material {
    float color[4];
} rtr_material;
#define RTR_SHADE_WIREFRAME 1u
#define RTR_SHADE_BVH 2u
int rtr_shade(rtr_ctx* ctx, const rtr_hit* hits, uint64_t n, const rtr_triangle* triangles, uint32_t nb_triangles,
              const rtr_mesh* meshes, uint32_t nb_meshes, const rtr_material* materials, uint32_t nb_materials,
              uint32_t flags, const float* bvh_rgba /* [n*4], only read with RTR_SHADE_BVH */, float* rgba_out /* [n*4] */);
int rtr_shade_dev(rtr_ctx* ctx, const rtr_hit* hits_dev, uint64_t n, const rtr_triangle* triangles_dev,
                  const rtr_mesh* meshes_dev, const rtr_material* materials_dev, uint32_t flags,
                  const float* bvh_rgba_dev, float* rgba_dev);
/* the BVH-depth overlay of the shader (uDepthDisplayBVH, raytracer.glsl:269-275 with intersectBVH's edge code 2,
 * :222-233): per pixel the colour of the node at depth `display_depth` (root = 0) that getClosestHitBVH visits last
 * among those the primary ray intersects -- (0.5,0,0.5,0.1), or (0.7,0,0.7,0.1) where the ray enters the box near one
 * of its edges; (0,0,0,0) where there is none.  out: 4 floats per pixel. */
int rtr_bvh_depth_overlay(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_camera* camera, uint32_t width, uint32_t height,
                          uint32_t denom_w, uint32_t denom_h, int display_depth, float* bvh_rgba_out);
int rtr_bvh_depth_overlay_dev(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_camera* camera, uint32_t width, uint32_t height,
                              uint32_t denom_w, uint32_t denom_h, int display_depth, float* bvh_rgba_dev);

/* weighted dealing: rank r owns stripes_of_rank[r] stripes (host array of nranks entries, at most
 * RTR_MAX_STRIPES_PER_RANK each, 255 in total; 0 = renders nothing, e.g. a rank busy rebuilding); block b belongs to
 * stripe b % sum(stripes_of_rank) and the stripes of a cycle are handed out round by round: round k to every rank
 * with more than k stripes, in rank order.  rtr_render_sharded_dev is the case of one stripe per rank. */
#define RTR_MAX_STRIPES_PER_RANK 16
int rtr_render_stripes_dev(rtr_ctx* ctx, const rtr_bvh* bvh, const rtr_camera* camera,
                           uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h,
                           uint32_t rows_per_block, const uint32_t* stripes_of_rank, uint32_t nranks, uint32_t rank,
                           uint32_t bounces, int shadow, const float light_pos[3], uint32_t flags,
                           float* rgba_dev, rtr_hit* primary_hits_dev, uint64_t* rays_traced_dev);

/* ---- scene ingestion (host only: no rtr_ctx, no device; errors through rtr_last_error(NULL)) ----
 * What produces the TriangleGPU[] / MeshModelGPU[] arrays the build consumes. */

/* replaces cr::Mesh::load (srcCommon/scene/geometry/mesh.cpp:186-263): the triangles of a Wavefront OBJ file in
 * face order, w = 1, every _ModelId = model_id (Mesh::_Id).  Parsing follows tinyobjloader 1.2.0 (the copy in the
 * reference's tree) bit for bit, including its decimal reader and its ear clipping of polygons; a malformed `f`
 * statement or a corner that names a vertex the file does not have is RTR_E_INVALID (the reference exits / reads
 * out of bounds).  *out_tris is malloc'ed by the library: release it with rtr_obj_free. */
int rtr_obj_load(const char* path, uint32_t model_id, rtr_triangle** out_tris, uint64_t* out_n);
int rtr_obj_parse(const char* text, uint64_t len, uint32_t model_id, rtr_triangle** out_tris, uint64_t* out_n);
void rtr_obj_free(rtr_triangle* tris);

/* replaces cr::Mesh::primitiveTriangle/Square/Cube/Sphere (mesh.cpp:64-183); the sphere is empty there too */
#define RTR_PRIMITIVE_TRIANGLE 0
#define RTR_PRIMITIVE_SQUARE 1
#define RTR_PRIMITIVE_CUBE 2
#define RTR_PRIMITIVE_SPHERE 3
int rtr_mesh_primitive(int which, uint32_t model_id, rtr_triangle* out, uint64_t cap, uint64_t* out_n);

/* replace cr::Mesh::setModel / setPosition / setScale / setRotation / setMaterial (mesh.cpp:18-62) on the
 * MeshModelGPU record; same GLM operation order, so the matrices are bit-identical to the reference's */
void rtr_mesh_init(rtr_mesh* m); /* identity, material 0 (mesh.hpp:13-14) */
void rtr_mesh_set_model(rtr_mesh* m, const float model[16]);
void rtr_mesh_set_position(rtr_mesh* m, float x, float y, float z);
void rtr_mesh_set_scale(rtr_mesh* m, float scale);
void rtr_mesh_set_rotation(rtr_mesh* m, float theta_x, float theta_y, float theta_z);
void rtr_mesh_set_material(rtr_mesh* m, uint32_t material_id);
/* replaces cr::Triangle::getCentroid(triangle, model) (triangle.cpp:30-32), the point the Morton codes are taken of */
void rtr_triangle_centroid(const rtr_triangle* t, const float model[16], float out[3]);

/* replaces cr::Camera (srcCommon/scene/camera.{hpp,cpp}) for hosts that cannot include rtr_scene.hpp: the CameraGPU
 * of a camera constructed like application.cpp:16-20 after replaying input events on it -- kind 0:
 * ProcessMouseMovement(a, b); kind 1..6: processKeyboard(FORWARD, BACKWARD, LEFT, RIGHT, UP, DOWN with deltaTime a),
 * _Accelerate = (b != 0).  Same code as cr::Camera of rtr_scene.hpp: bit-identical to the reference's getGpuData(). */
int rtr_camera_gpu_data(const float eye[3], float aspect, float fov, float near_plane, float far_plane, int n_events,
                        const int* kind, const float* a, const float* b, rtr_camera* out);

/* ---- multi-GPU (one process per GPU; NCCL resolved at run time with dlopen("libnccl.so.2"),
 * so inside a torch process it is the very library torch.distributed already loaded) ---- */
#define RTR_NCCL_UNIQUE_ID_BYTES 128
int rtr_comm_unique_id(void* id_out /* 128 B */);
int rtr_comm_init(rtr_ctx* ctx, const void* unique_id, int rank, int nranks);
int rtr_comm_destroy(rtr_ctx* ctx);
/* root: sends flat nodes + triangles + meshes of *bvh; others: receive into a BVH they own */
int rtr_bvh_broadcast(rtr_ctx* ctx, rtr_bvh** bvh, int root);
/* lighter replica for ranks that only trace in the default order: the 64-byte traversal records, node 0 and the
 * trace constants (64 B*(2n-1) instead of ~208 B per triangle).  Receivers cannot use RTR_TRACE_REFERENCE_ORDER,
 * the node/triangle accessors or re-broadcast (RTR_E_STATE).  expected_triangles: 0 = the root announces the size
 * first (one host round trip on every rank); otherwise every rank passes the same triangle count and the call
 * only enqueues -- a receiver can then post it ahead of the root's rebuild (pipelined frames). */
int rtr_bvh_broadcast_traversal(rtr_ctx* ctx, rtr_bvh** bvh, int root, uint32_t expected_triangles);
/* rows of the image are dealt to ranks in blocks of `rows_per_block`, block b -> rank b % nranks */
int rtr_allgather_rows(rtr_ctx* ctx, void* image_dev, uint32_t width, uint32_t height, uint32_t bytes_per_pixel,
                       uint32_t rows_per_block);

/* the same for the weighted dealing of rtr_render_stripes_dev (stripes_of_rank: host array of nranks entries) */
int rtr_allgather_stripes(rtr_ctx* ctx, void* image_dev, uint32_t width, uint32_t height, uint32_t bytes_per_pixel,
                          uint32_t rows_per_block, const uint32_t* stripes_of_rank);

/* only `root` ends up with the whole image: every block travels once, from its owner to the root (ncclSend/ncclRecv) */
int rtr_gather_stripes(rtr_ctx* ctx, void* image_dev, uint32_t width, uint32_t height, uint32_t bytes_per_pixel,
                       uint32_t rows_per_block, const uint32_t* stripes_of_rank, int root);

/* Sharded host traffic of a multi-GPU frame loop (no reference counterpart: the reference is single-GPU).
 * rtr_gather_slices: an array of n_elems elements (the frame's TriangleGPU records) is cut into nranks contiguous
 * slices, rank r holding [n_elems*r/nranks, n_elems*(r+1)/nranks) in slice_dev -- uploaded over ITS OWN PCIe link --
 * and the slices are assembled in full_dev on the root over NVLink (every slice travels once; full_dev may be NULL on
 * the other ranks).  rtr_download_stripes_async: this rank's row blocks of a striped frame (rtr_render_stripes_dev)
 * are copied from image_dev to the same rows of image_host, a full-size host image (e.g. one pinned buffer shared by
 * all ranks); asynchronous on the context's stream, no collective. */
int rtr_gather_slices(rtr_ctx* ctx, const void* slice_dev, void* full_dev, uint64_t n_elems, uint32_t elem_bytes, int root);
int rtr_download_stripes_async(rtr_ctx* ctx, void* image_host, const void* image_dev, uint32_t width, uint32_t height,
                               uint32_t bytes_per_pixel, uint32_t rows_per_block, const uint32_t* stripes_of_rank);

#ifdef __cplusplus
}
#endif
#endif /* RTR_H */
