// morton.cu -- scene box reduction + Morton codes over triangle centroids.
//
// Replaces BVH::getMortonCodes (bvh.cpp:330-348) = getSceneAABB (:235-251) ->
// getCircumscribedCube (:253-303, including the Q1 axis pick) -> getTrianglesCentroids
// (:305-312, triangle.cpp:31-33) -> getNormalizedCentroids (:314-328) -> morton3D (:358-372),
// and the stub shader ploc/preprocessing/preprocessing.glsl.
//
// Two HBM-streaming kernels, 64 B read per triangle each:
//   scene_aabb_kernel : min/max of the 3 transformed vertices over the WHOLE passed array (Q2),
//                       warp shuffle -> CTA -> 6 ordered-uint atomics per CTA
//   morton_kernel     : every CTA re-derives the cube from the 6 reduced values (20 flops),
//                       one thread per triangle writes code (+ iota index, + optional 63-bit code)
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock)
scene_aabb_kernel(const rtr_triangle* __restrict__ tris, uint32_t array_len,
                  const rtr_mesh* __restrict__ meshes, uint32_t* __restrict__ ordered6) {
    float mn[3] = {INFINITY, INFINITY, INFINITY};
    float mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < array_len; i += gridDim.x * kBlock) {
        TriRec t = load_tri(tris, i);
        Mat3x4 M = load_model(meshes, t.model_id);
        float3 a = mat_mul_point(M, t.p0), b = mat_mul_point(M, t.p1), c = mat_mul_point(M, t.p2);
        mn[0] = fminf(mn[0], fminf(a.x, fminf(b.x, c.x)));
        mn[1] = fminf(mn[1], fminf(a.y, fminf(b.y, c.y)));
        mn[2] = fminf(mn[2], fminf(a.z, fminf(b.z, c.z)));
        mx[0] = fmaxf(mx[0], fmaxf(a.x, fmaxf(b.x, c.x)));
        mx[1] = fmaxf(mx[1], fmaxf(a.y, fmaxf(b.y, c.y)));
        mx[2] = fmaxf(mx[2], fmaxf(a.z, fmaxf(b.z, c.z)));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    __shared__ float s[kBlock / 32][6];
    const uint32_t w = threadIdx.x >> 5, l = lane_id();
    if (l == 0) {
        s[w][0] = mn[0]; s[w][1] = mn[1]; s[w][2] = mn[2];
        s[w][3] = mx[0]; s[w][4] = mx[1]; s[w][5] = mx[2];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[0][threadIdx.x];
        for (int i = 1; i < kBlock / 32; ++i)
            v = (threadIdx.x < 3) ? fminf(v, s[i][threadIdx.x]) : fmaxf(v, s[i][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(&ordered6[threadIdx.x], float_to_ordered(v));
        else atomicMax(&ordered6[threadIdx.x], float_to_ordered(v));
    }
}

__global__ void init_ordered_kernel(uint32_t* ordered6) {
    if (threadIdx.x < 3) ordered6[threadIdx.x] = float_to_ordered(INFINITY);
    else if (threadIdx.x < 6) ordered6[threadIdx.x] = float_to_ordered(-INFINITY);
}

// bvh.cpp:253-303.  The else-if chain picks X whenever distX > 0 (Q1).
__device__ __forceinline__ void circumscribed_cube(const float s[6], float c[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i] = s[i];
    const float distX = s[3] - s[0], distY = s[4] - s[1], distZ = s[5] - s[2];
    float maxDist = 0.f;
    int axis;
    if (distX > maxDist) { maxDist = distX; axis = 0; }
    else if (distY > maxDist) { maxDist = distY; axis = 1; }
    else { maxDist = distZ; axis = 2; }
    float delta;
    if (axis == 0) {
        delta = (maxDist - distY) / 2.f; c[4] += delta; c[1] -= delta;
        delta = (maxDist - distZ) / 2.f; c[5] += delta; c[2] -= delta;
    } else if (axis == 1) {
        delta = (maxDist - distX) / 2.f; c[3] += delta; c[0] -= delta;
        delta = (maxDist - distZ) / 2.f; c[5] += delta; c[2] -= delta;
    } else {
        delta = (maxDist - distX) / 2.f; c[3] += delta; c[0] -= delta;
        delta = (maxDist - distY) / 2.f; c[4] += delta; c[1] -= delta;
    }
}

// bvh.cpp:350-356
__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint64_t expand_bits21(uint32_t v) {
    uint64_t x = v & 0x1FFFFFu;
    x = (x | (x << 32)) & 0x001F00000000FFFFull;
    x = (x | (x << 16)) & 0x001F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

// std::min(std::max(v*S, 0), S-1) then C cast (bvh.cpp:363-369); finite inputs only (Q4)
__device__ __forceinline__ uint32_t quantise(float v, float scale, float top) {
    float t = v * scale;
    t = (t < 0.0f) ? 0.0f : t;   // std::max(t, 0.0f)
    t = (top < t) ? top : t;     // std::min(t, top)
    return __float2uint_rz(t);
}

__global__ void __launch_bounds__(kBlock)
morton_kernel(const rtr_triangle* __restrict__ tris, uint32_t n,
              const rtr_mesh* __restrict__ meshes, const uint32_t* __restrict__ ordered6,
              uint32_t* __restrict__ codes, uint32_t* __restrict__ indices,
              uint64_t* __restrict__ codes64, float* __restrict__ bounds12) {
    __shared__ float s_cube[6];
    if (threadIdx.x == 0) {
        float s[6], c[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) s[i] = ordered_to_float(ordered6[i]);
        circumscribed_cube(s, c);
#pragma unroll
        for (int i = 0; i < 6; ++i) s_cube[i] = c[i];
        if (blockIdx.x == 0 && bounds12) {
#pragma unroll
            for (int i = 0; i < 6; ++i) { bounds12[i] = s[i]; bounds12[6 + i] = c[i]; }
        }
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const float minX = s_cube[0], minY = s_cube[1], minZ = s_cube[2];
    const float lenX = s_cube[3] - minX, lenY = s_cube[4] - minY, lenZ = s_cube[5] - minZ;

    TriRec t = load_tri(tris, i);
    Mat3x4 M = load_model(meshes, t.model_id);
    // triangle.cpp:32 : ((1.f/3.f) * model) * ((P0 + P1) + P2)
    const float third = 1.f / 3.f;
    Mat3x4 S;
    S.c0x = M.c0x * third; S.c0y = M.c0y * third; S.c0z = M.c0z * third;
    S.c1x = M.c1x * third; S.c1y = M.c1y * third; S.c1z = M.c1z * third;
    S.c2x = M.c2x * third; S.c2y = M.c2y * third; S.c2z = M.c2z * third;
    S.c3x = M.c3x * third; S.c3y = M.c3y * third; S.c3z = M.c3z * third;
    float4 sum;
    sum.x = __fadd_rn(__fadd_rn(t.p0.x, t.p1.x), t.p2.x);
    sum.y = __fadd_rn(__fadd_rn(t.p0.y, t.p1.y), t.p2.y);
    sum.z = __fadd_rn(__fadd_rn(t.p0.z, t.p1.z), t.p2.z);
    sum.w = __fadd_rn(__fadd_rn(t.p0.w, t.p1.w), t.p2.w);
    const float3 c = mat_mul_point(S, sum);
    // bvh.cpp:319-324
    const float nx = __fdiv_rn(__fsub_rn(c.x, minX), lenX);
    const float ny = __fdiv_rn(__fsub_rn(c.y, minY), lenY);
    const float nz = __fdiv_rn(__fsub_rn(c.z, minZ), lenZ);
    if (codes) {
        const uint32_t xx = expand_bits(quantise(nx, 1024.0f, 1023.0f));
        const uint32_t yy = expand_bits(quantise(ny, 1024.0f, 1023.0f));
        const uint32_t zz = expand_bits(quantise(nz, 1024.0f, 1023.0f));
        codes[i] = (xx << 2) | (yy << 1) | zz;
    }
    if (indices) indices[i] = i;
    if (codes64) {
        const uint64_t xx = expand_bits21(quantise(nx, 2097152.0f, 2097151.0f));
        const uint64_t yy = expand_bits21(quantise(ny, 2097152.0f, 2097151.0f));
        const uint64_t zz = expand_bits21(quantise(nz, 2097152.0f, 2097151.0f));
        codes64[i] = (xx << 2) | (yy << 1) | zz;
    }
}

}  // namespace

int rtr_morton_launch(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t array_len,
                      const rtr_mesh* meshes, uint32_t* codes, uint32_t* indices, uint64_t* codes64,
                      float* bounds12, uint32_t* ordered6) {
    init_ordered_kernel<<<1, 32, 0, ctx->stream>>>(ordered6);
    RTR_LAUNCH_CHECK(ctx);
    if (array_len) {
        uint32_t blocks = (array_len + kBlock - 1) / kBlock;
        const uint32_t cap = (uint32_t)ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
        if (blocks > cap) blocks = cap;
        RTR_PROF(ctx, "scene_aabb_kernel");
        scene_aabb_kernel<<<blocks, kBlock, 0, ctx->stream>>>(tris, array_len, meshes, ordered6);
        RTR_LAUNCH_CHECK(ctx);
    }
    const uint32_t blocks = n ? (n + kBlock - 1) / kBlock : 1;
    RTR_PROF(ctx, "morton_kernel");
    morton_kernel<<<blocks, kBlock, 0, ctx->stream>>>(tris, n, meshes, ordered6, codes, indices, codes64, bounds12);
    RTR_LAUNCH_CHECK(ctx);
    return RTR_OK;
}
