// common.cuh -- shared declarations of librtr_b200 (sm_100a only, no CPU fallback).
//
// Bit-exactness: the whole library is compiled with -fmad=false so that every fp32
// expression is a sequence of individually rounded IEEE operations in the association
// order written in the source, which is the reference's order.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rtr.h"

#define RTR_SM_COUNT_B200 148

struct rtr_ctx {
    int device = 0;
    int sm_count = RTR_SM_COUNT_B200;
    int reserved_sms = 0;  // SMs the persistent traversal leaves free (rtr_ctx_reserve_sms)
    uint32_t* sm_table = nullptr;  // 2 x [1 + 1024], per launch: SMs claimed so far, then one state word per %smid
    uint32_t sm_table_turn = 0;
    int ploc_ctas_per_sm[2] = {0, 0};  // persistent PLOC loop kernel (full radius / masked radius), resident CTAs per SM
    int flatten_ctas_per_sm = 0;
    // SM partition for the rays (rtr_ctx_partition_sms): a green context and its streams
    void* green_ctx = nullptr;
    int partition_sms = 0;
    std::vector<cudaStream_t> partition_streams;
    int trace_ctas_per_sm = 0;     // resident CTAs per SM of the persistent traversal kernel on this device (0: not asked yet)
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    uint64_t launches = 0;
    std::string err;
    // growable device workspace (sort double buffers, tile status, histograms ...)
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // small pinned host scratch for counters read back during a build
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // optional per-kernel timing (rtr_ctx_profile_*): CUDA events around the launches that matter
    bool profiling = false;
    struct ProfRec { const char* name; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
    bool prof_open = false;
    // NCCL (resolved with dlopen at rtr_comm_init)
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
};

int rtr_set_error(rtr_ctx* ctx, int code, const char* fmt, ...);
int rtr_ws_reserve(rtr_ctx* ctx, size_t bytes);  // ensures ctx->ws has >= bytes

#define RTR_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess)                                                               \
            return rtr_set_error((ctx), (_e == cudaErrorMemoryAllocation) ? RTR_E_NOMEM : RTR_E_CUDA, \
                                 "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)

#define RTR_CHECK(expr)              \
    do {                             \
        int _r = (expr);             \
        if (_r != RTR_OK) return _r; \
    } while (0)

void rtr_prof_begin(rtr_ctx* ctx, const char* name);
void rtr_prof_end(rtr_ctx* ctx);

// put before a kernel launch whose duration should be reported by rtr_ctx_profile_read
#define RTR_PROF(ctx, name)                               \
    do {                                                  \
        if ((ctx)->profiling) rtr_prof_begin((ctx), name); \
    } while (0)

#define RTR_LAUNCH_CHECK(ctx)                   \
    do {                                        \
        (ctx)->launches++;                      \
        if ((ctx)->prof_open) rtr_prof_end(ctx); \
        RTR_CUDA((ctx), cudaGetLastError());    \
    } while (0)

static inline size_t rtr_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Weighted dealing of image row blocks (rtr_render_stripes_dev / rtr_allgather_stripes): block b belongs to
// stripe b % V, V = sum of stripes_of_rank; the stripes of one cycle are handed out round by round -- round k
// goes to every rank with more than k stripes, in rank order -- so that the ranks' blocks interleave and a
// partial last cycle cannot favour anyone by more than a block.  Returns the owner of every stripe.
static inline std::vector<int> rtr_stripe_owners(const uint32_t* stripes_of_rank, int nranks) {
    std::vector<int> owner;
    uint32_t most = 0;
    for (int r = 0; r < nranks; ++r) most = stripes_of_rank[r] > most ? stripes_of_rank[r] : most;
    for (uint32_t k = 0; k < most; ++k)
        for (int r = 0; r < nranks; ++r)
            if (stripes_of_rank[r] > k) owner.push_back(r);
    return owner;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Reference record views used by the kernels (identical layouts to rtr.h)
struct alignas(16) TriRec {
    float4 p0, p1, p2;
    uint32_t model_id, pad0, pad1, pad2;
};
static_assert(sizeof(TriRec) == 64, "TriangleGPU is 64 B");
static_assert(sizeof(rtr_mesh) == 68, "MeshModelGPU is 68 B");
static_assert(sizeof(rtr_node) == 48, "BVH_NodeGPU is 48 B");
static_assert(sizeof(rtr_hit) == 24, "Hit is 24 B");
static_assert(sizeof(rtr_ray) == 32, "Ray is 32 B");

// rows 0..2 of a column-major mat4 (the w row is never needed on this path)
struct Mat3x4 {
    float c0x, c0y, c0z;
    float c1x, c1y, c1z;
    float c2x, c2y, c2z;
    float c3x, c3y, c3z;
};

__device__ __forceinline__ Mat3x4 load_model(const rtr_mesh* __restrict__ meshes, uint32_t id) {
    const float* m = meshes[id].model;  // 68-byte stride: only 4-byte aligned
    Mat3x4 M;
    M.c0x = __ldg(m + 0);  M.c0y = __ldg(m + 1);  M.c0z = __ldg(m + 2);
    M.c1x = __ldg(m + 4);  M.c1y = __ldg(m + 5);  M.c1z = __ldg(m + 6);
    M.c2x = __ldg(m + 8);  M.c2y = __ldg(m + 9);  M.c2z = __ldg(m + 10);
    M.c3x = __ldg(m + 12); M.c3y = __ldg(m + 13); M.c3z = __ldg(m + 14);
    return M;
}

// glm 0.9.9.9 mat4*vec4 (type_mat4x4.inl:536-583): (m0*v0 + m1*v1) + (m2*v2 + m3*v3) per row
__device__ __forceinline__ float3 mat_mul_point(const Mat3x4& M, const float4 v) {
    float3 r;
    r.x = __fadd_rn(__fadd_rn(__fmul_rn(M.c0x, v.x), __fmul_rn(M.c1x, v.y)),
                    __fadd_rn(__fmul_rn(M.c2x, v.z), __fmul_rn(M.c3x, v.w)));
    r.y = __fadd_rn(__fadd_rn(__fmul_rn(M.c0y, v.x), __fmul_rn(M.c1y, v.y)),
                    __fadd_rn(__fmul_rn(M.c2y, v.z), __fmul_rn(M.c3y, v.w)));
    r.z = __fadd_rn(__fadd_rn(__fmul_rn(M.c0z, v.x), __fmul_rn(M.c1z, v.y)),
                    __fadd_rn(__fmul_rn(M.c2z, v.z), __fmul_rn(M.c3z, v.w)));
    return r;
}

__device__ __forceinline__ TriRec load_tri(const rtr_triangle* __restrict__ tris, uint32_t i) {
    const uint4* p = reinterpret_cast<const uint4*>(tris) + (size_t)i * 4;
    uint4 a = __ldg(p + 0), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    TriRec t;
    t.p0 = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
    t.p1 = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
    t.p2 = make_float4(__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z), __uint_as_float(c.w));
    t.model_id = d.x; t.pad0 = t.pad1 = t.pad2 = 0;
    return t;
}

// order-preserving float <-> uint map, for atomicMin/atomicMax on floats
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// acquire/release accessors for decoupled look-back status words
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------
// internal entry points (one .cu per subsystem)
// ---------------------------------------------------------------------------------------
// sort.cu
int rtr_sort_impl_u32(rtr_ctx* ctx, uint32_t* keys, uint32_t* vals, uint32_t n, int begin_bit, int end_bit);
int rtr_sort_impl_u64(rtr_ctx* ctx, uint64_t* keys, uint32_t* vals, uint32_t n, int begin_bit, int end_bit);
size_t rtr_sort_ws_bytes(uint32_t n, int key_bytes, bool pairs);
int rtr_bit_histogram32_launch(rtr_ctx* ctx, const uint32_t* keys, uint32_t n, uint32_t* out);
int rtr_digitplace_scan_launch(rtr_ctx* ctx, const uint32_t* in, uint32_t* out);
// morton.cu
int rtr_morton_launch(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t array_len,
                      const rtr_mesh* meshes, uint32_t* codes, uint32_t* indices /*nullable: iota*/,
                      uint64_t* codes64 /*nullable*/, float* bounds12 /*device, 12 floats*/,
                      uint32_t* bounds_ordered6 /*device scratch, 6 u32*/);
