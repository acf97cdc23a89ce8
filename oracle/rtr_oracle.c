/*
 * rtr_oracle.c -- CPU restatement of the reference's acceleration-structure and
 * ray-cast path (plain C99).  TEST INFRASTRUCTURE ONLY -- see rtr_oracle.h for
 * who may load it and for the parity-pinning status of each half.
 *
 * Build: gcc -std=c99 -O2 -ffp-contract=off -fopenmp -fPIC -shared  (oracle/Makefile)
 * -ffp-contract=off is REQUIRED: the reference is built without FMA contraction
 * (CMakeLists.txt:26-28, x86-64 SSE2 scalar), every float expression below keeps the
 * reference's association order.
 *
 * Citations are relative to /root/reference.
 */
#include "rtr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_NONE 0xFFFFFFFFu

/* ------------------------------------------------------------------------- */
/* small float helpers: every op is one rounded IEEE fp32 op                  */
/* ------------------------------------------------------------------------- */

/* glm 0.9.9.9 type_mat4x4.inl:536-583 (mat4 * vec4): per row
 * (m[0][r]*v0 + m[1][r]*v1) + (m[2][r]*v2 + m[3][r]*v3), column-major m[c*4+r] */
static void mat4_mul_vec4(const float* m, const float* v, float* out) {
    for (int r = 0; r < 4; ++r) {
        float mul0 = m[0 * 4 + r] * v[0];
        float mul1 = m[1 * 4 + r] * v[1];
        float add0 = mul0 + mul1;
        float mul2 = m[2 * 4 + r] * v[2];
        float mul3 = m[3 * 4 + r] * v[3];
        float add1 = mul2 + mul3;
        out[r] = add0 + add1;
    }
}

static float fmin2(float a, float b) { return (b < a) ? b : a; } /* std::min(a,b) */
static float fmax2(float a, float b) { return (a < b) ? b : a; } /* std::max(a,b) */

static void cross3(const float* x, const float* y, float* o) { /* GLSL cross */
    float o0 = x[1] * y[2] - y[1] * x[2];
    float o1 = x[2] * y[0] - y[2] * x[0];
    float o2 = x[0] * y[1] - y[0] * x[1];
    o[0] = o0; o[1] = o1; o[2] = o2;
}
static float dot3(const float* a, const float* b) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
/* normalize(v) is defined for this build as v / sqrt(dot(v,v)) component-wise.
 * (GLSL leaves the exact formula to the driver; this is the pinned definition
 * shared by the oracle and the CUDA kernels.) */
static void normalize3(const float* v, float* o) {
    float len = sqrtf(dot3(v, v));
    o[0] = v[0] / len; o[1] = v[1] / len; o[2] = v[2] / len;
}

/* ------------------------------------------------------------------------- */
/* sort pre-passes (tests/testsSortGPU)                                       */
/* ------------------------------------------------------------------------- */

/* histogramOfGlobalDigitCounts.glsl:31-59 as asserted by
 * testHistogramCreation.cpp:29-41 : out[b] = #keys with bit b set */
void orc_bit_histogram32(const uint32_t* keys, uint32_t n, uint32_t out[32]) {
    for (int b = 0; b < 32; ++b) out[b] = 0;
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t b = 0; b < 32; ++b)
            if (keys[i] & (1u << b)) out[b]++;
}

/* prefixSumOfGlobalDigitCounts.glsl:24-69 as asserted by
 * testHistogramPrefixSum.cpp:43-51 : exclusive scan inside each group of 4 bins */
void orc_digitplace_exclusive_scan(const uint32_t in[32], uint32_t out[32]) {
    for (uint32_t j = 0; j < 8; ++j) {
        out[j * 4] = 0;
        for (uint32_t i = 1; i < 4; ++i)
            out[j * 4 + i] = in[j * 4 + i - 1] + out[j * 4 + i - 1];
    }
}

/* ------------------------------------------------------------------------- */
/* Morton codes                                                               */
/* ------------------------------------------------------------------------- */

/* bvh.cpp:235-251 -- iterates the WHOLE vector (Q2), hence array_len */
void orc_scene_aabb(const orc_triangle* tris, uint32_t array_len,
                    const orc_mesh* meshes, float out[6]) {
    float mn[3] = {INFINITY, INFINITY, INFINITY};
    float mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = 0; i < array_len; ++i) {
        const orc_triangle* t = &tris[i];
        const float* M = meshes[t->model_id].m;
        float p0[4], p1[4], p2[4];
        mat4_mul_vec4(M, t->p0, p0);
        mat4_mul_vec4(M, t->p1, p1);
        mat4_mul_vec4(M, t->p2, p2);
        for (int a = 0; a < 3; ++a) {
            mx[a] = fmax2(mx[a], fmax2(p0[a], fmax2(p1[a], p2[a])));
            mn[a] = fmin2(mn[a], fmin2(p0[a], fmin2(p1[a], p2[a])));
        }
    }
    out[0] = mn[0]; out[1] = mn[1]; out[2] = mn[2];
    out[3] = mx[0]; out[4] = mx[1]; out[5] = mx[2];
}

/* bvh.cpp:253-303 -- including the else-if axis pick (Q1): X whenever distX > 0 */
void orc_circumscribed_cube(const float s[6], float c[6]) {
    for (int i = 0; i < 6; ++i) c[i] = s[i];
    float distX = s[3] - s[0];
    float distY = s[4] - s[1];
    float distZ = s[5] - s[2];
    float maxDist = 0.f;
    int axis = 0;
    if (distX > maxDist) { maxDist = distX; axis = 0; }
    else if (distY > maxDist) { maxDist = distY; axis = 1; }
    else { maxDist = distZ; axis = 2; }
    float delta;
    switch (axis) {
    case 0:
        delta = (maxDist - distY) / 2.f; c[4] += delta; c[1] -= delta;
        delta = (maxDist - distZ) / 2.f; c[5] += delta; c[2] -= delta;
        break;
    case 1:
        delta = (maxDist - distX) / 2.f; c[3] += delta; c[0] -= delta;
        delta = (maxDist - distZ) / 2.f; c[5] += delta; c[2] -= delta;
        break;
    default:
        delta = (maxDist - distX) / 2.f; c[3] += delta; c[0] -= delta;
        delta = (maxDist - distY) / 2.f; c[4] += delta; c[1] -= delta;
        break;
    }
}

/* triangle.cpp:31-33: vec3(((1.f/3.f) * model) * ((P0 + P1) + P2)) */
static void centroid(const orc_triangle* t, const float* M, float* c3) {
    const float third = 1.f / 3.f;
    float Ms[16];
    for (int i = 0; i < 16; ++i) Ms[i] = M[i] * third; /* type_mat4x4.inl:526-533 */
    float s[4], o[4];
    for (int i = 0; i < 4; ++i) s[i] = (t->p0[i] + t->p1[i]) + t->p2[i];
    mat4_mul_vec4(Ms, s, o);
    c3[0] = o[0]; c3[1] = o[1]; c3[2] = o[2];
}

/* bvh.cpp:350-356 */
static uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

/* bvh.cpp:358-372 */
static uint32_t morton3d(const float p[3]) {
    float x = fmin2(fmax2(p[0] * 1024.0f, 0.0f), 1023.0f);
    float y = fmin2(fmax2(p[1] * 1024.0f, 0.0f), 1023.0f);
    float z = fmin2(fmax2(p[2] * 1024.0f, 0.0f), 1023.0f);
    uint32_t xx = expand_bits((uint32_t)x);
    uint32_t yy = expand_bits((uint32_t)y);
    uint32_t zz = expand_bits((uint32_t)z);
    return (xx << 2) | (yy << 1) | zz;
}

/* bvh.cpp:330-348 via :305-328 */
void orc_morton_codes(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                      const orc_mesh* meshes, uint32_t* codes) {
    float scene[6], cube[6];
    orc_scene_aabb(tris, array_len, meshes, scene);
    orc_circumscribed_cube(scene, cube);
    float lenX = cube[3] - cube[0];
    float lenY = cube[4] - cube[1];
    float lenZ = cube[5] - cube[2];
    for (uint32_t i = 0; i < n; ++i) {
        float c[3], p[3];
        centroid(&tris[i], meshes[tris[i].model_id].m, c);
        p[0] = (c[0] - cube[0]) / lenX;
        p[1] = (c[1] - cube[1]) / lenY;
        p[2] = (c[2] - cube[2]) / lenZ;
        codes[i] = morton3d(p);
    }
}

/* 64-bit extension (no reference definition exists): same normalised
 * centroid, 21 bits per axis, q = (uint32)min(max(v*2097152, 0), 2097151),
 * interleaved x (bit 2), y (bit 1), z (bit 0) per triple as in morton3D. */
static uint64_t expand_bits21(uint32_t v) {
    uint64_t x = v & 0x1FFFFFu;
    x = (x | (x << 32)) & 0x001F00000000FFFFull;
    x = (x | (x << 16)) & 0x001F0000FF0000FFull;
    x = (x | (x << 8))  & 0x100F00F00F00F00Full;
    x = (x | (x << 4))  & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2))  & 0x1249249249249249ull;
    return x;
}
void orc_morton_codes64(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                        const orc_mesh* meshes, uint64_t* codes) {
    float scene[6], cube[6];
    orc_scene_aabb(tris, array_len, meshes, scene);
    orc_circumscribed_cube(scene, cube);
    float lenX = cube[3] - cube[0];
    float lenY = cube[4] - cube[1];
    float lenZ = cube[5] - cube[2];
    for (uint32_t i = 0; i < n; ++i) {
        float c[3], p[3];
        centroid(&tris[i], meshes[tris[i].model_id].m, c);
        p[0] = (c[0] - cube[0]) / lenX;
        p[1] = (c[1] - cube[1]) / lenY;
        p[2] = (c[2] - cube[2]) / lenZ;
        uint32_t q[3];
        for (int a = 0; a < 3; ++a) {
            float v = fmin2(fmax2(p[a] * 2097152.0f, 0.0f), 2097151.0f);
            q[a] = (uint32_t)v;
        }
        codes[i] = (expand_bits21(q[0]) << 2) | (expand_bits21(q[1]) << 1) | expand_bits21(q[2]);
    }
}

/* ------------------------------------------------------------------------- */
/* sort                                                                       */
/* ------------------------------------------------------------------------- */

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}

/* bvh.cpp:223-231: std::sort over pair<code,index>; lexicographic pair order is
 * a total order here (indices unique) so any comparison sort gives the same
 * result, which equals a stable sort by code. */
void orc_sort_pairs(uint32_t* codes, uint32_t* indices, uint32_t n) {
    uint64_t* tmp = (uint64_t*)malloc((size_t)n * sizeof(uint64_t));
    for (uint32_t i = 0; i < n; ++i) tmp[i] = ((uint64_t)codes[i] << 32) | indices[i];
    qsort(tmp, n, sizeof(uint64_t), cmp_u64);
    for (uint32_t i = 0; i < n; ++i) {
        codes[i] = (uint32_t)(tmp[i] >> 32);
        indices[i] = (uint32_t)tmp[i];
    }
    free(tmp);
}

/* Host LSD radix-256 sort "of the same design" as the GPU path:
 * one up-front histogram for all digit places, then one stable scatter per place. */
void orc_radix_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t n) {
    uint32_t* k2 = (uint32_t*)malloc((size_t)n * 4);
    uint32_t* v2 = vals ? (uint32_t*)malloc((size_t)n * 4) : NULL;
    static const int P = 4;
    uint32_t (*hist)[256] = calloc(P, sizeof(*hist));
    for (uint32_t i = 0; i < n; ++i)
        for (int p = 0; p < P; ++p) hist[p][(keys[i] >> (8 * p)) & 255]++;
    uint32_t *ks = keys, *kd = k2, *vs = vals, *vd = v2;
    for (int p = 0; p < P; ++p) {
        uint32_t off[256], sum = 0;
        for (int d = 0; d < 256; ++d) { off[d] = sum; sum += hist[p][d]; }
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t d = (ks[i] >> (8 * p)) & 255;
            uint32_t o = off[d]++;
            kd[o] = ks[i];
            if (vals) vd[o] = vs[i];
        }
        uint32_t* t = ks; ks = kd; kd = t;
        t = vs; vs = vd; vd = t;
    }
    /* P even: result is back in keys/vals */
    free(hist); free(k2); free(v2);
}
/* The same stable LSD radix-256 sort on `threads` host cores (OpenMP): every thread histograms and scatters a
 * contiguous chunk, chunk offsets per digit come from a (digit-major, thread-minor) exclusive scan, which keeps the
 * sort stable.  The host-side counterpart of the Onesweep kernel for the per-stage CPU baseline. */
void orc_radix_sort_pairs_mt(uint32_t* keys, uint32_t* vals, uint32_t n, int threads) {
    if (n < 2) return;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    uint32_t* k2 = (uint32_t*)malloc((size_t)n * 4);
    uint32_t* v2 = (uint32_t*)malloc((size_t)n * 4);
    uint32_t* hist = (uint32_t*)malloc((size_t)threads * 256 * 4);
    uint32_t *kin = keys, *vin = vals, *kout = k2, *vout = v2;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
        {
#ifdef _OPENMP
            const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
            const int t = 0, nt = 1;
#endif
            const uint64_t lo = (uint64_t)n * t / nt, hi = (uint64_t)n * (t + 1) / nt;
            uint32_t* h = hist + (size_t)t * 256;
            memset(h, 0, 256 * 4);
            for (uint64_t i = lo; i < hi; ++i) h[(kin[i] >> shift) & 255u]++;
#ifdef _OPENMP
#pragma omp barrier
#pragma omp single
#endif
            {
                uint32_t run = 0;
                for (int d = 0; d < 256; ++d)
                    for (int q = 0; q < nt; ++q) { uint32_t c = hist[(size_t)q * 256 + d]; hist[(size_t)q * 256 + d] = run; run += c; }
            }
            for (uint64_t i = lo; i < hi; ++i) {
                const uint32_t dst = h[(kin[i] >> shift) & 255u]++;
                kout[dst] = kin[i]; vout[dst] = vin[i];
            }
        }
        uint32_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    /* 4 passes: the result is back in keys / vals */
    free(k2); free(v2); free(hist);
}

void orc_radix_sort_keys_u32(uint32_t* keys, uint32_t n) { orc_radix_sort_pairs(keys, NULL, n); }

void orc_radix_sort_pairs_u64(uint64_t* keys, uint32_t* vals, uint32_t n) {
    uint64_t* k2 = (uint64_t*)malloc((size_t)n * 8);
    uint32_t* v2 = vals ? (uint32_t*)malloc((size_t)n * 4) : NULL;
    uint64_t *ks = keys, *kd = k2; uint32_t *vs = vals, *vd = v2;
    for (int p = 0; p < 8; ++p) {
        uint32_t hist[256]; memset(hist, 0, sizeof(hist));
        for (uint32_t i = 0; i < n; ++i) hist[(ks[i] >> (8 * p)) & 255]++;
        uint32_t off[256], sum = 0;
        for (int d = 0; d < 256; ++d) { off[d] = sum; sum += hist[d]; }
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t d = (uint32_t)((ks[i] >> (8 * p)) & 255);
            uint32_t o = off[d]++;
            kd[o] = ks[i];
            if (vals) vd[o] = vs[i];
        }
        uint64_t* t = ks; ks = kd; kd = t;
        uint32_t* u = vs; vs = vd; vd = u;
    }
    free(k2); free(v2);
}
void orc_radix_sort_keys_u64(uint64_t* keys, uint32_t n) { orc_radix_sort_pairs_u64(keys, NULL, n); }

/* ------------------------------------------------------------------------- */
/* PLOC                                                                       */
/* ------------------------------------------------------------------------- */

/* bvh.cpp:383-399 */
static void aabb_from_triangle(const orc_triangle* t, const orc_mesh* mesh, orc_node* nd) {
    float p0[4], p1[4], p2[4];
    mat4_mul_vec4(mesh->m, t->p0, p0);
    mat4_mul_vec4(mesh->m, t->p1, p1);
    mat4_mul_vec4(mesh->m, t->p2, p2);
    for (int a = 0; a < 3; ++a) {
        nd->bmin[a] = fmin2(p0[a], fmin2(p1[a], p2[a]));
        nd->bmax[a] = fmax2(p0[a], fmax2(p1[a], p2[a]));
    }
}

/* bvh.cpp:401-413 then :378-381 : 2 * ((dx*dy + dy*dz) + dz*dx) */
static float merged_area(const orc_node* a, const orc_node* b) {
    float mn[3], mx[3];
    for (int k = 0; k < 3; ++k) {
        mn[k] = fmin2(a->bmin[k], b->bmin[k]);
        mx[k] = fmax2(a->bmax[k], b->bmax[k]);
    }
    float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    return 2 * (dx * dy + dy * dz + dz * dx);
}

static orc_bvh* bvh_build_impl(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                               const orc_mesh* meshes, uint32_t nb_meshes,
                               uint32_t search_radius, int key_bits);

orc_bvh* orc_bvh_build(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                       const orc_mesh* meshes, uint32_t nb_meshes,
                       uint32_t search_radius) {
    return bvh_build_impl(tris, n, array_len, meshes, nb_meshes, search_radius, 32);
}
/* the same builder over 63-bit Morton keys (21 bits per axis, orc_morton_codes64): only the leaf order changes.
 * No reference definition exists (bvh.cpp has 30-bit codes only); the 64-bit order refines the 32-bit one
 * (code64 >> 33 == code32), ties are broken by the original index like std::sort over pairs does (bvh.cpp:223-231). */
orc_bvh* orc_bvh_build64(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                         const orc_mesh* meshes, uint32_t nb_meshes,
                         uint32_t search_radius) {
    return bvh_build_impl(tris, n, array_len, meshes, nb_meshes, search_radius, 64);
}

static orc_bvh* bvh_build_impl(const orc_triangle* tris, uint32_t n, uint32_t array_len,
                               const orc_mesh* meshes, uint32_t nb_meshes,
                               uint32_t search_radius, int key_bits) {
    if (!tris || !meshes || n == 0 || array_len < n || nb_meshes == 0) return NULL;
    orc_bvh* b = (orc_bvh*)calloc(1, sizeof(orc_bvh));
    uint32_t nc = 2 * n - 1;
    b->n = n; b->nb_clusters = nc;
    b->morton_sorted = (uint32_t*)malloc((size_t)n * 4);
    b->triangle_indices = (uint32_t*)malloc((size_t)n * 4);
    b->clusters = (orc_node*)calloc(nc, sizeof(orc_node));
    b->parent = (uint32_t*)malloc((size_t)nc * 4);
    b->left = (uint32_t*)malloc((size_t)nc * 4);
    b->right = (uint32_t*)malloc((size_t)nc * 4);
    for (uint32_t i = 0; i < nc; ++i) b->parent[i] = b->left[i] = b->right[i] = ORC_NONE;
    uint32_t cap_it = 256;
    b->trace_active = (uint32_t*)malloc(cap_it * 4);
    b->trace_merges = (uint32_t*)malloc(cap_it * 4);

    /* bvh.cpp:214-232 */
    for (uint32_t i = 0; i < n; ++i) b->triangle_indices[i] = i;
    if (key_bits == 64) {
        b->morton_sorted64 = (uint64_t*)malloc((size_t)n * 8);
        orc_morton_codes64(tris, n, array_len, meshes, b->morton_sorted64);
        orc_radix_sort_pairs_u64(b->morton_sorted64, b->triangle_indices, n);  /* stable: ties keep index order */
        for (uint32_t i = 0; i < n; ++i) b->morton_sorted[i] = (uint32_t)(b->morton_sorted64[i] >> 33);
    } else {
        orc_morton_codes(tris, n, array_len, meshes, b->morton_sorted);
        orc_sort_pairs(b->morton_sorted, b->triangle_indices, n);
    }

    /* bvh.cpp:26-46 */
    uint32_t* c_in = (uint32_t*)malloc((size_t)n * 4);
    uint32_t* c_out = (uint32_t*)malloc((size_t)n * 4);
    uint32_t* nn = (uint32_t*)calloc(n, 4);      /* persists across iterations (Q4) */
    uint32_t* prefix = (uint32_t*)malloc((size_t)n * 4);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t ti = b->triangle_indices[i];
        orc_node* leaf = &b->clusters[i];
        leaf->triangle_id = ti;
        aabb_from_triangle(&tris[ti], &meshes[tris[ti].model_id], leaf);
        c_in[i] = i;
    }
    uint32_t iter_n = n, total = n;

    /* bvh.cpp:62-122 (serial order == OMP_NUM_THREADS=1 == canonical ids, Q3) */
    while (iter_n > 1) {
        /* :193-210 */
        for (uint32_t i = 0; i < iter_n; ++i) {
            float min_dist = INFINITY;
            const orc_node* ci = &b->clusters[c_in[i]];
            int s0 = (int)i - (int)search_radius; if (s0 < 0) s0 = 0;
            uint32_t start = (uint32_t)s0;
            uint32_t end = i + search_radius + 1; if (end > iter_n) end = iter_n;
            for (uint32_t j = start; j < end; ++j) {
                if (j == i) continue;
                float d = merged_area(ci, &b->clusters[c_in[j]]);
                if (d < min_dist) { min_dist = d; nn[i] = j; }
            }
        }
        /* :159-191 ; invalid == ORC_NONE */
        uint32_t merges = 0;
        for (uint32_t i = 0; i < iter_n; ++i) {
            uint32_t j = nn[i];
            if (nn[j] == i && i < j) {
                uint32_t ci = c_in[i], cj = c_in[j];
                uint32_t id = total++;
                orc_node* m = &b->clusters[id];   /* :415-420 links/tri stay 0 */
                for (int k = 0; k < 3; ++k) {
                    m->bmin[k] = fmin2(b->clusters[ci].bmin[k], b->clusters[cj].bmin[k]);
                    m->bmax[k] = fmax2(b->clusters[ci].bmax[k], b->clusters[cj].bmax[k]);
                }
                b->left[id] = ci; b->right[id] = cj;
                b->parent[ci] = id; b->parent[cj] = id;
                c_in[j] = ORC_NONE;
                c_in[i] = id;
                merges++;
            }
        }
        /* :125-148 exclusive scan of validity, :150-156 compaction */
        uint32_t run = 0;
        for (uint32_t i = 0; i < iter_n; ++i) {
            prefix[i] = run;
            if (c_in[i] != ORC_NONE) { c_out[run] = c_in[i]; run++; }
        }
        if (b->nb_iterations == cap_it) {
            cap_it *= 2;
            b->trace_active = (uint32_t*)realloc(b->trace_active, cap_it * 4);
            b->trace_merges = (uint32_t*)realloc(b->trace_merges, cap_it * 4);
        }
        b->trace_active[b->nb_iterations] = iter_n;
        b->trace_merges[b->nb_iterations] = merges;
        b->nb_iterations++;
        /* :105-113 */
        iter_n = run;
        uint32_t* t = c_in; c_in = c_out; c_out = t;
        if (merges == 0) break; /* Q4 guard: reference would loop forever */
    }
    free(c_in); free(c_out); free(nn); free(prefix);
    return b;
}

void orc_bvh_destroy(orc_bvh* b) {
    if (!b) return;
    free(b->morton_sorted); free(b->morton_sorted64); free(b->triangle_indices); free(b->clusters);
    free(b->parent); free(b->left); free(b->right);
    free(b->trace_active); free(b->trace_merges);
    free(b);
}
uint32_t orc_bvh_nb_iterations(const orc_bvh* b) { return b->nb_iterations; }
const uint32_t* orc_bvh_morton_sorted(const orc_bvh* b) { return b->morton_sorted; }
const uint64_t* orc_bvh_morton_sorted64(const orc_bvh* b) { return b->morton_sorted64; }
const uint32_t* orc_bvh_triangle_indices(const orc_bvh* b) { return b->triangle_indices; }
const orc_node* orc_bvh_clusters(const orc_bvh* b) { return b->clusters; }
const uint32_t* orc_bvh_parent(const orc_bvh* b) { return b->parent; }
const uint32_t* orc_bvh_left(const orc_bvh* b) { return b->left; }
const uint32_t* orc_bvh_right(const orc_bvh* b) { return b->right; }
const uint32_t* orc_bvh_trace_active(const orc_bvh* b) { return b->trace_active; }
const uint32_t* orc_bvh_trace_merges(const orc_bvh* b) { return b->trace_merges; }

/* scene.cpp:189-208, iterative so deep trees cannot overflow the C stack.
 * Pre-order from root 2n-2; node at output position p gets Left = p+1 and
 * Right = first position after the left subtree; leaves keep 0/0 + TriangleId. */
void orc_flatten(const orc_node* clusters, const uint32_t* left, const uint32_t* right,
                 uint32_t n, orc_node* flat) {
    uint32_t nc = 2 * n - 1;
    /* stack entries: (cluster id, parent flat position or NONE) for pending right children */
    uint32_t* st_id = (uint32_t*)malloc((size_t)nc * 4);
    uint32_t* st_parent = (uint32_t*)malloc((size_t)nc * 4);
    uint32_t sp = 0, out = 0;
    st_id[sp] = nc - 1; st_parent[sp] = ORC_NONE; sp++;
    while (sp) {
        sp--;
        uint32_t id = st_id[sp], par = st_parent[sp];
        for (;;) {
            uint32_t pos = out++;
            flat[pos] = clusters[id];
            if (par != ORC_NONE) { flat[par].right = pos; par = ORC_NONE; }
            if (left[id] == ORC_NONE) break; /* leaf (Q9) */
            flat[pos].left = pos + 1;
            st_id[sp] = right[id]; st_parent[sp] = pos; sp++;
            id = left[id];
        }
    }
    free(st_id); free(st_parent);
}

uint64_t orc_hash_words(const uint32_t* w, uint64_t nwords) {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t i = 0; i < nwords; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return h;
}
uint64_t orc_hash_flat_nodes(const orc_node* flat, uint32_t nb) {
    uint64_t h = 1469598103934665603ull;
    for (uint32_t i = 0; i < nb; ++i) {
        uint32_t w[9];
        memcpy(&w[0], flat[i].bmin, 12);
        memcpy(&w[3], flat[i].bmax, 12);
        w[6] = flat[i].triangle_id; w[7] = flat[i].left; w[8] = flat[i].right;
        for (int k = 0; k < 9; ++k) h = (h ^ w[k]) * 1099511628211ull;
    }
    return h;
}

/* ------------------------------------------------------------------------- */
/* traversal                                                                  */
/* ------------------------------------------------------------------------- */

/* raytracer.glsl:92-100 + :303-305 */
void orc_get_ray(const orc_camera* cam, uint32_t x, uint32_t y,
                 uint32_t denom_w, uint32_t denom_h, orc_ray* ray) {
    float px = (float)x / (float)denom_w;
    float py = (float)y / (float)denom_h;
    float pv[4];
    pv[0] = (px - 0.5f) * cam->plane_width;
    pv[1] = (py - 0.5f) * cam->plane_height;
    pv[2] = 1.f * cam->plane_near;
    pv[3] = 1.f;
    float pw[4];
    mat4_mul_vec4(cam->inv_view, pv, pw);
    float d[4];
    for (int i = 0; i < 4; ++i) { ray->origin[i] = cam->eye[i]; d[i] = pw[i] - cam->eye[i]; }
    /* vec4 normalize: ((x2+y2)+z2)+w2 */
    float len = sqrtf(((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) + d[3] * d[3]);
    ray->direction[0] = d[0] / len;
    ray->direction[1] = d[1] / len;
    ray->direction[2] = d[2] / len;
    ray->direction[3] = 0.f;
}

/* every pixel's ray, row-major; pixels outside the denominators (Q5) keep a zero record */
void orc_get_rays(const orc_camera* cam, uint32_t width, uint32_t height,
                  uint32_t denom_w, uint32_t denom_h, orc_ray* out) {
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            orc_ray* r = &out[(size_t)y * width + x];
            memset(r, 0, sizeof(*r));
            if (x < denom_w && y < denom_h) orc_get_ray(cam, x, y, denom_w, denom_h, r);
        }
}

/* world-space vertices in SHADER naming (Q8): shader _P1 = host _P2, shader _P2 = host _P1 */
static void shader_vertices(const orc_triangle* t, const orc_mesh* meshes,
                            float p0[4], float p1[4], float p2[4]) {
    const float* M = meshes[t->model_id].m;
    mat4_mul_vec4(M, t->p0, p0);
    mat4_mul_vec4(M, t->p2, p1);
    mat4_mul_vec4(M, t->p1, p2);
}

/* raytracer.glsl:102-147 */
void orc_ray_triangle(const orc_ray* ray, const orc_triangle* tris,
                      const orc_mesh* meshes, uint32_t tri_index, orc_hit* hit) {
    float p0[4], p1[4], p2[4];
    shader_vertices(&tris[tri_index], meshes, p0, p1, p2);
    const float* d = ray->direction;
    float e0[3], e1[3], nrm[3], n[3], q[3];
    for (int i = 0; i < 3; ++i) { e0[i] = p1[i] - p0[i]; e1[i] = p2[i] - p0[i]; }
    cross3(e1, e0, nrm);
    normalize3(nrm, n);
    cross3(d, e1, q);
    float a = dot3(e0, q);
    const float epsilon = 1e-4f;
    memset(hit, 0, sizeof(*hit));
    if (dot3(n, d) >= 0 || fabsf(a) < epsilon) return;
    float s[3], r[3];
    for (int i = 0; i < 3; ++i) s[i] = (ray->origin[i] - p0[i]) / a;
    cross3(s, e0, r);
    float bx = dot3(s, q);
    float by = dot3(r, d);
    float bz = 1 - bx - by;
    if (bx < 0 || by < 0 || bz < 0) return;
    float t = dot3(e1, r);
    if (t < 0) return;
    hit->b0 = bx; hit->b1 = by; hit->b2 = bz; hit->t = t;
    hit->did_hit = 1; hit->triangle_id = tri_index;
}

/* raytracer.glsl:182-237; GLSL min/max follow IEEE fminf/fmaxf here (Q11);
 * the edge-overlay return code 2 is a debug colour and is folded into 1 */
uint32_t orc_intersect_box(const orc_ray* ray, const orc_node* node) {
    float tMin, tMax;
    float ix = 1.0f / ray->direction[0];
    float tx1 = (node->bmin[0] - ray->origin[0]) * ix;
    float tx2 = (node->bmax[0] - ray->origin[0]) * ix;
    tMin = fminf(tx1, tx2);
    tMax = fmaxf(tx1, tx2);
    if (tMax < 0.f || tMin > tMax) return 0;
    float iy = 1.0f / ray->direction[1];
    float ty1 = (node->bmin[1] - ray->origin[1]) * iy;
    float ty2 = (node->bmax[1] - ray->origin[1]) * iy;
    tMin = fmaxf(tMin, fminf(ty1, ty2));
    tMax = fminf(tMax, fmaxf(ty1, ty2));
    if (tMax < 0.f || tMin > tMax) return 0;
    float iz = 1.0f / ray->direction[2];
    float tz1 = (node->bmin[2] - ray->origin[2]) * iz;
    float tz2 = (node->bmax[2] - ray->origin[2]) * iz;
    tMin = fmaxf(tMin, fminf(tz1, tz2));
    tMax = fminf(tMax, fmaxf(tz1, tz2));
    if (tMax >= 0.f && tMin <= tMax) return 1;
    return 0;
}

/* raytracer.glsl:246-295 with the guarded-miss semantics of :152-153 (Q6).
 * LIFO stack, push Left then Right => Right visited first; strict '<' keeps the
 * first-visited leaf on equal t. */
void orc_closest_hit_bvh(const orc_ray* ray, const orc_node* flat,
                         const orc_triangle* tris, const orc_mesh* meshes,
                         orc_hit* out, uint64_t* nodes_visited) {
    orc_hit closest; memset(&closest, 0, sizeof(closest));
    uint32_t stack[1024];
    int sp = 0;
    uint64_t visited = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        uint32_t idx = stack[--sp];
        const orc_node* nd = &flat[idx];
        visited++;
        if (orc_intersect_box(ray, nd) != 0) {
            if (nd->left == 0 && nd->right == 0) {
                orc_hit h;
                orc_ray_triangle(ray, tris, meshes, nd->triangle_id, &h);
                if (h.did_hit && (closest.did_hit == 0 || h.t < closest.t)) closest = h;
            } else if (sp + 2 <= 1024) {
                stack[sp++] = nd->left;
                stack[sp++] = nd->right;
            }
        }
    }
    *out = closest;
    if (nodes_visited) *nodes_visited = visited;
}

/* raytracer.glsl:149-157 */
void orc_closest_hit_brute(const orc_ray* ray, const orc_triangle* tris, uint32_t n,
                           const orc_mesh* meshes, orc_hit* out) {
    orc_hit closest; memset(&closest, 0, sizeof(closest));
    for (uint32_t i = 0; i < n; ++i) {
        orc_hit h;
        orc_ray_triangle(ray, tris, meshes, i, &h);
        if (h.did_hit == 0) continue;
        if (closest.did_hit == 0 || h.t < closest.t) closest = h;
    }
    *out = closest;
}

/* any-hit (new; shadow rays do not exist in the reference, TODO.md:13-14):
 * 1 iff some triangle whose leaf box chain is hit passes rayTriangleIntersection
 * with t < t_max.  Order-independent. */
int orc_any_hit_bvh(const orc_ray* ray, float t_max, const orc_node* flat,
                    const orc_triangle* tris, const orc_mesh* meshes) {
    uint32_t stack[1024];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        uint32_t idx = stack[--sp];
        const orc_node* nd = &flat[idx];
        if (orc_intersect_box(ray, nd) != 0) {
            if (nd->left == 0 && nd->right == 0) {
                orc_hit h;
                orc_ray_triangle(ray, tris, meshes, nd->triangle_id, &h);
                if (h.did_hit && h.t < t_max) return 1;
            } else if (sp + 2 <= 1024) {
                stack[sp++] = nd->left;
                stack[sp++] = nd->right;
            }
        }
    }
    return 0;
}

/* ---- secondary rays: definitions of THIS build (none exist in the reference) ----
 * front-face unit normal n = normalize(cross(e1, e0)) exactly as raytracer.glsl:113;
 * hit point h = o + d*t; both secondary rays start at h + n*1e-3.
 * bounce: r = d - n*(2*dot(d,n)), direction = normalize(r).
 * shadow: l = light - origin, direction = l/|l|, t_max = |l|. */
static void hit_frame(const orc_ray* in, const orc_hit* hit, const orc_triangle* tris,
                      const orc_mesh* meshes, float n[3], float org[3]) {
    float p0[4], p1[4], p2[4], e0[3], e1[3], nrm[3];
    shader_vertices(&tris[hit->triangle_id], meshes, p0, p1, p2);
    for (int i = 0; i < 3; ++i) { e0[i] = p1[i] - p0[i]; e1[i] = p2[i] - p0[i]; }
    cross3(e1, e0, nrm);
    normalize3(nrm, n);
    for (int i = 0; i < 3; ++i) {
        float h = in->origin[i] + in->direction[i] * hit->t;
        org[i] = h + n[i] * 1e-3f;
    }
}

int orc_bounce_ray(const orc_ray* in, const orc_hit* hit, const orc_triangle* tris,
                   const orc_mesh* meshes, orc_ray* out) {
    if (!hit->did_hit) return 0;
    float n[3], org[3], r[3], rn[3];
    hit_frame(in, hit, tris, meshes, n, org);
    float k = 2.f * dot3(in->direction, n);
    for (int i = 0; i < 3; ++i) r[i] = in->direction[i] - n[i] * k;
    normalize3(r, rn);
    for (int i = 0; i < 3; ++i) { out->origin[i] = org[i]; out->direction[i] = rn[i]; }
    out->origin[3] = 1.f; out->direction[3] = 0.f;
    return 1;
}

int orc_shadow_ray(const orc_ray* in, const orc_hit* hit, const orc_triangle* tris,
                   const orc_mesh* meshes, const float light_pos[3],
                   orc_ray* out, float* t_max) {
    if (!hit->did_hit) return 0;
    float n[3], org[3], l[3];
    hit_frame(in, hit, tris, meshes, n, org);
    for (int i = 0; i < 3; ++i) l[i] = light_pos[i] - org[i];
    float len = sqrtf(dot3(l, l));
    for (int i = 0; i < 3; ++i) { out->origin[i] = org[i]; out->direction[i] = l[i] / len; }
    out->origin[3] = 1.f; out->direction[3] = 0.f;
    *t_max = len;
    return 1;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_trace_primary(const orc_node* flat, const orc_triangle* tris,
                       const orc_mesh* meshes, const orc_camera* cam,
                       uint32_t width, uint32_t height,
                       uint32_t denom_w, uint32_t denom_h,
                       orc_hit* hits, int threads) {
    (void)threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
#endif
    for (int64_t y = 0; y < (int64_t)height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            orc_ray ray;
            orc_get_ray(cam, x, (uint32_t)y, denom_w, denom_h, &ray);
            orc_closest_hit_bvh(&ray, flat, tris, meshes, &hits[(size_t)y * width + x], NULL);
        }
}

void orc_trace_rays(const orc_node* flat, const orc_triangle* tris,
                    const orc_mesh* meshes, const orc_ray* rays, uint64_t n_rays,
                    orc_hit* hits, int threads) {
    (void)threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads)
#endif
    for (int64_t i = 0; i < (int64_t)n_rays; ++i)
        orc_closest_hit_bvh(&rays[i], flat, tris, meshes, &hits[i], NULL);
}

/* Image definition of THIS build for multi-bounce frames (the reference shades
 * primary hits with a flat material colour only, raytracer.glsl:159-179):
 *   L = sum_k 0.5^k * c_k over path vertices k = 0..bounces,
 *   c_k = -dot(n_k, d_k)                        (shadow == 0, head-light term)
 *   c_k = vis_k * max(0, dot(n_k, l_k))          (shadow != 0, point light)
 * rgba = (L, L, L, 1).  Pixels outside [0,denom) x [0,denom) (Q5) are not traced
 * and stay (0,0,0,1).  Sum order is k ascending, one rounded op each. */
void orc_render(const orc_node* flat, const orc_triangle* tris,
                const orc_mesh* meshes, const orc_camera* cam,
                uint32_t width, uint32_t height, uint32_t denom_w, uint32_t denom_h,
                uint32_t row0, uint32_t row1, uint32_t bounces, int shadow,
                const float light_pos[3],
                float* rgba, orc_hit* primary_hits, uint64_t* rays_traced, int threads) {
    (void)threads; (void)height;
    uint64_t total = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads) reduction(+:total)
#endif
    for (int64_t y = row0; y < (int64_t)row1; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            size_t o = (size_t)(y - row0) * width + x;
            float L = 0.f, w = 1.f;
            orc_hit first; memset(&first, 0, sizeof(first));
            if (x < denom_w && (uint32_t)y < denom_h) {
                orc_ray ray;
                orc_get_ray(cam, x, (uint32_t)y, denom_w, denom_h, &ray);
                for (uint32_t k = 0; k <= bounces; ++k) {
                    orc_hit h;
                    orc_closest_hit_bvh(&ray, flat, tris, meshes, &h, NULL);
                    total++;
                    if (k == 0) first = h;
                    if (!h.did_hit) break;
                    float n[3], org[3], c;
                    hit_frame(&ray, &h, tris, meshes, n, org);
                    if (shadow) {
                        orc_ray sr; float tmax;
                        orc_shadow_ray(&ray, &h, tris, meshes, light_pos, &sr, &tmax);
                        int occ = orc_any_hit_bvh(&sr, tmax, flat, tris, meshes);
                        total++;
                        float ndl = dot3(n, sr.direction);
                        c = occ ? 0.f : fmaxf(0.f, ndl);
                    } else {
                        c = -dot3(n, ray.direction);
                    }
                    L = L + w * c;
                    w = w * 0.5f;
                    if (k < bounces) {
                        orc_ray nr;
                        orc_bounce_ray(&ray, &h, tris, meshes, &nr);
                        ray = nr;
                    }
                }
            }
            if (rgba) { rgba[o * 4 + 0] = L; rgba[o * 4 + 1] = L; rgba[o * 4 + 2] = L; rgba[o * 4 + 3] = 1.f; }
            if (primary_hits) primary_hits[o] = first;
        }
    if (rays_traced) *rays_traced = total;
}


/* raytracer.glsl:159-179 (getColor) + :299-331 (main), uIsBVHDisplayed == false */
void orc_shade(const orc_hit* hits, uint64_t n, const orc_triangle* tris, const orc_mesh* meshes,
               const float* materials, int wireframe, const float* bvh_rgba, float* rgba_out) {
    for (uint64_t i = 0; i < n; ++i) {
        float c[4] = {0.f, 0.f, 0.f, 1.f};                         /* :300 */
        if (bvh_rgba) memcpy(c, bvh_rgba + 4 * i, 16);             /* :161-163, uIsBVHDisplayed */
        const orc_hit* h = &hits[i];
        if (h->did_hit != 0) {                                      /* :165 */
            const uint32_t model = tris[h->triangle_id].model_id;           /* :166 */
            const float* m = materials + 4 * (size_t)meshes[model].material_id;
            for (int k = 0; k < 4; ++k) c[k] = c[k] + m[k];         /* :167 */
            if (wireframe) {                                        /* :170-178 */
                const float th = 0.02f;                             /* :71 */
                if (h->b0 < th || h->b1 < th || h->b2 < th) { c[0] = c[1] = c[2] = 0.f; c[3] = 1.f; }
            }
        }
        for (int k = 0; k < 4; ++k) rgba_out[4 * i + k] = c[k];    /* :330 */
    }
}

/* raytracer.glsl:182-237 with its return code 2: the ray enters the box within BVH_LINE_WIDTH / (depth + 1) of two of
 * its three slab pairs (a box edge, drawn as a line by the overlay) */
uint32_t orc_intersect_box_edge(const orc_ray* ray, const orc_node* node, int display_depth) {
    float tMin, tMax;
    float ix = 1.0f / ray->direction[0];
    float tx1 = (node->bmin[0] - ray->origin[0]) * ix;
    float tx2 = (node->bmax[0] - ray->origin[0]) * ix;
    tMin = fminf(tx1, tx2);
    tMax = fmaxf(tx1, tx2);
    if (tMax < 0.f || tMin > tMax) return 0;
    float iy = 1.0f / ray->direction[1];
    float ty1 = (node->bmin[1] - ray->origin[1]) * iy;
    float ty2 = (node->bmax[1] - ray->origin[1]) * iy;
    tMin = fmaxf(tMin, fminf(ty1, ty2));
    tMax = fminf(tMax, fmaxf(ty1, ty2));
    if (tMax < 0.f || tMin > tMax) return 0;
    float iz = 1.0f / ray->direction[2];
    float tz1 = (node->bmin[2] - ray->origin[2]) * iz;
    float tz2 = (node->bmax[2] - ray->origin[2]) * iz;
    tMin = fmaxf(tMin, fminf(tz1, tz2));
    tMax = fminf(tMax, fmaxf(tz1, tz2));
    if (tMax >= 0.f && tMin <= tMax) {
        const float threshold = 0.05f / ((float)display_depth + 1.f);      /* :72, :223 */
        float e[3];
        for (int k = 0; k < 3; ++k) e[k] = ray->origin[k] + ray->direction[k] * tMin;  /* :224 */
        const int cx = fabsf(e[0] - node->bmin[0]) < threshold || fabsf(e[0] - node->bmax[0]) < threshold;
        const int cy = fabsf(e[1] - node->bmin[1]) < threshold || fabsf(e[1] - node->bmax[1]) < threshold;
        const int cz = fabsf(e[2] - node->bmin[2]) < threshold || fabsf(e[2] - node->bmax[2]) < threshold;
        if ((cx && cy) || (cx && cz) || (cy && cz)) return 2;
        return 1;
    }
    return 0;
}

/* The bvhColor of getClosestHitBVH (raytracer.glsl:246-295), the shader's loop to the letter: LIFO stack with a depth
 * stack, Left pushed before Right, no pruning; every intersected node at depth == display_depth overwrites the
 * colour (:269-275), so the LAST one visited decides.  out: 4 floats per pixel, (0,0,0,0) where nothing at that depth
 * is hit (:321). */
void orc_depth_overlay(const orc_node* flat, const orc_camera* cam, uint32_t width, uint32_t height,
                       uint32_t denom_w, uint32_t denom_h, int display_depth, float* out) {
    static const float kBox[4] = {0.5f, 0.f, 0.5f, 0.1f}, kLine[4] = {0.7f, 0.f, 0.7f, 0.1f};  /* :69-70 */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int64_t y = 0; y < (int64_t)height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            float* c = out + 4 * ((size_t)y * width + x);
            c[0] = c[1] = c[2] = c[3] = 0.f;
            if (x >= denom_w || (uint32_t)y >= denom_h) continue;
            orc_ray ray;
            orc_get_ray(cam, x, (uint32_t)y, denom_w, denom_h, &ray);
            uint32_t stack[1024], depth[1024];
            int sp = 0;
            stack[sp] = 0; depth[sp] = 0; sp++;
            while (sp > 0) {
                --sp;
                const uint32_t idx = stack[sp], d = depth[sp];
                const orc_node* nd = &flat[idx];
                const uint32_t code = orc_intersect_box_edge(&ray, nd, display_depth);
                if (code != 0) {
                    if ((int)d == display_depth) memcpy(c, code == 2 ? kLine : kBox, 16);
                    if (!(nd->left == 0 && nd->right == 0) && sp + 2 <= 1024) {
                        stack[sp] = nd->left; depth[sp] = d + 1; sp++;
                        stack[sp] = nd->right; depth[sp] = d + 1; sp++;
                    }
                }
            }
        }
}
