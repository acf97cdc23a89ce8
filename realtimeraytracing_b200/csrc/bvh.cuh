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
    uint32_t pad1;
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
    bool built = false;

    // inputs (device)
    const rtr_triangle* tris = nullptr;
    const rtr_mesh* meshes = nullptr;
    rtr_triangle* tris_own = nullptr;  size_t tris_own_cap = 0;   // triangles
    rtr_mesh* meshes_own = nullptr;    size_t meshes_own_cap = 0; // meshes

    // build arrays (device), sized by capacity
    uint32_t* codes = nullptr;     // [cap]  sorted Morton codes
    uint32_t* tri_idx = nullptr;   // [cap]  BVH_Params::_TriangleIndices
    float4* node = nullptr;        // [2*(2cap-1)] by cluster id, 32 B records: (min.xyz, max.x)(max.y, max.z, bits(left | triangle id), bits(right | NONE))
    uint32_t* isize = nullptr;     // [cap] nodes in the subtree of internal cluster (id - n)
    uint32_t* ipos = nullptr;      // [cap] DFS pre-order position of internal cluster (id - n)
    uint32_t* order = nullptr;     // [2cap-1] cluster id at every DFS pre-order position (flatten pass 1)
    uint32_t* cin = nullptr;       // [cap] active list, ping
    uint32_t* cout = nullptr;      // [cap] active list, pong
    uint64_t* tile_status = nullptr;  // [ceil(cap/tile)] decoupled look-back words
    PlocState* state = nullptr;    // [2]
    uint32_t* trace_active = nullptr;   // [kMaxPlocIterations]
    uint32_t* trace_merges = nullptr;   // [kMaxPlocIterations]
    uint32_t* iter_first_id = nullptr;  // [kMaxPlocIterations + 1] first cluster id created by iteration i
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
    // child-pair records read by the default traversal: one 64-byte record per inner node of the flat array
    // (indexed by flat index; leaf entries unused) = both children's boxes + their links, copied bit for
    // bit from the flat nodes by pack_pairs_kernel:
    //   (L.min.xyz, L.max.x) (L.max.yz, R.min.xy) (R.min.z, R.max.xyz) (L.word, R.word, L.aux, R.aux)
    //   word = child flat index, or 0x80000000 | index for a leaf whose aux is its wtri slot
    uint4* pairs = nullptr;        // [4*(2cap-1)]
    uint4* pairs_own = nullptr;    // adopted / received BVHs
    size_t pairs_own_cap = 0;      // triangles pairs_own was sized for
    const uint4* pairs_view = nullptr;

    // host mirrors of the last build
    uint32_t iterations = 0;
    std::vector<uint32_t> h_first_id;  // [iterations + 1]
    bool timing = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float stage_ms[6] = {0, 0, 0, 0, 0, 0};
};

// ploc.cu
int rtr_bvh_run_build(rtr_bvh* b);
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
                      uint32_t rows_per_block = 0, uint32_t shard_rank = 0, uint32_t shard_count = 1);
