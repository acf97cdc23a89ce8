// api.cu -- the extern "C" surface declared in include/rtr.h: context, buffers, host-pointer
// wrappers around the device entry points, BVH object lifetime and accessors.
#include <stdarg.h>

#include <new>

#include "bvh.cuh"
#include <stdlib.h>
#ifndef RTR_L2_FETCH_DEFAULT
#define RTR_L2_FETCH_DEFAULT 0
#endif

static thread_local std::string g_create_error;

int rtr_set_error(rtr_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

int rtr_ws_reserve(rtr_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return RTR_OK;
    if (ctx->ws) {
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        RTR_CUDA(ctx, cudaFree(ctx->ws));
        ctx->ws = nullptr; ctx->ws_bytes = 0;
    }
    const size_t want = rtr_align_up(bytes + bytes / 8, 1 << 20);
    RTR_CUDA(ctx, cudaMalloc(&ctx->ws, want));
    ctx->ws_bytes = want;
    return RTR_OK;
}

static cudaEvent_t prof_event(rtr_ctx* ctx) {
    cudaEvent_t e = nullptr;
    if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
void rtr_prof_begin(rtr_ctx* ctx, const char* name) {
    rtr_ctx::ProfRec r{name, prof_event(ctx), prof_event(ctx)};
    cudaEventRecord(r.e0, ctx->stream);
    ctx->prof.push_back(r);
    ctx->prof_open = true;
}
void rtr_prof_end(rtr_ctx* ctx) {
    cudaEventRecord(ctx->prof.back().e1, ctx->stream);
    ctx->prof_open = false;
}

namespace {

template <typename T>
int dev_alloc(rtr_ctx* ctx, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    return RTR_OK;
}

void bvh_free_arrays(rtr_bvh* b) {
    void* ptrs[] = {b->codes, b->tri_idx, b->node, b->isize, b->ipos, b->order, b->codes64, b->cin, b->cout, b->tile_status,
                    b->state, b->ctl, b->iter_ns, b->trace_active, b->trace_merges, b->iter_first_id, b->bounds12, b->ordered6, b->flat,
                    b->tparams, b->tris_own, b->meshes_own, b->flat_recv, b->wtri, b->wtri_own, b->pairs, b->pairs_own};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    b->codes = b->tri_idx = b->ipos = b->order = b->cin = b->cout = nullptr; b->isize = nullptr;
    b->codes64 = nullptr; b->codes64_cap = 0;
    b->node = nullptr;
    b->tile_status = nullptr; b->state = nullptr; b->ctl = nullptr;
    b->trace_active = b->trace_merges = b->iter_first_id = nullptr; b->iter_ns = nullptr;
    b->bounds12 = nullptr; b->ordered6 = nullptr; b->flat = nullptr; b->tparams = nullptr;
    b->tris_own = nullptr; b->meshes_own = nullptr; b->tris_own_cap = b->meshes_own_cap = 0;
    b->flat_recv = nullptr; b->recv_cap = 0;
    b->wtri = nullptr; b->wtri_own = nullptr; b->wtri_own_cap = 0; b->wtri_view = nullptr;
    b->pairs = nullptr; b->pairs_own = nullptr; b->pairs_own_cap = 0; b->pairs_view = nullptr;
    b->capacity = 0;
}

int bvh_reserve(rtr_bvh* b, uint32_t n) {
    rtr_ctx* ctx = b->ctx;
    if (b->capacity >= n && !b->adopted) return RTR_OK;
    if (b->capacity) RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rtr_triangle* keep_t = b->tris_own; size_t keep_tc = b->tris_own_cap;
    rtr_mesh* keep_m = b->meshes_own; size_t keep_mc = b->meshes_own_cap;
    b->tris_own = nullptr; b->meshes_own = nullptr;
    bvh_free_arrays(b);  // also drops a broadcast receive buffer
    b->tris_own = keep_t; b->tris_own_cap = keep_tc; b->meshes_own = keep_m; b->meshes_own_cap = keep_mc;
    b->adopted = false; b->trav_only = false;
    const size_t cap = n, nc = 2 * (size_t)n - 1;
    const size_t tiles = (cap + kPlocTile - 1) / kPlocTile + 1;
    RTR_CHECK(dev_alloc(ctx, &b->codes, cap));
    RTR_CHECK(dev_alloc(ctx, &b->tri_idx, cap));
    RTR_CHECK(dev_alloc(ctx, &b->node, 2 * nc));
    RTR_CHECK(dev_alloc(ctx, &b->isize, cap));
    RTR_CHECK(dev_alloc(ctx, &b->ipos, cap));
    RTR_CHECK(dev_alloc(ctx, &b->order, nc));
    RTR_CHECK(dev_alloc(ctx, &b->cin, cap));
    RTR_CHECK(dev_alloc(ctx, &b->cout, cap));
    RTR_CHECK(dev_alloc(ctx, &b->tile_status, tiles));
    RTR_CHECK(dev_alloc(ctx, &b->state, 3));
    RTR_CHECK(dev_alloc(ctx, &b->ctl, 4));
    RTR_CHECK(dev_alloc(ctx, &b->trace_active, kMaxPlocIterations));
    RTR_CHECK(dev_alloc(ctx, &b->trace_merges, kMaxPlocIterations));
    RTR_CHECK(dev_alloc(ctx, &b->iter_first_id, kMaxPlocIterations + 1));
    RTR_CHECK(dev_alloc(ctx, &b->iter_ns, kMaxPlocIterations + 2));
    RTR_CHECK(dev_alloc(ctx, &b->bounds12, 12));
    RTR_CHECK(dev_alloc(ctx, &b->ordered6, 8));
    RTR_CHECK(dev_alloc(ctx, &b->flat, nc));
    RTR_CHECK(dev_alloc(ctx, &b->tparams, 1));
    RTR_CHECK(dev_alloc(ctx, &b->wtri, 3 * cap));
    RTR_CHECK(dev_alloc(ctx, &b->pairs, 4 * nc));
    b->capacity = n;
    return RTR_OK;
}

int check_build_args(rtr_ctx* ctx, const void* tris, uint32_t n, uint32_t array_len, const void* meshes,
                     uint32_t nb_meshes, uint32_t radius, rtr_bvh** out) {
    if (!ctx) return RTR_E_INVALID;
    if (!out) return rtr_set_error(ctx, RTR_E_INVALID, "bvh_build: out is NULL");
    if (!tris || !meshes) return rtr_set_error(ctx, RTR_E_INVALID, "bvh_build: NULL triangles or meshes");
    if (n == 0) return rtr_set_error(ctx, RTR_E_INVALID, "bvh_build: nb_triangles == 0");
    if (array_len < n) return rtr_set_error(ctx, RTR_E_INVALID, "bvh_build: tris_array_len %u < nb_triangles %u", array_len, n);
    if (nb_meshes == 0) return rtr_set_error(ctx, RTR_E_INVALID, "bvh_build: nb_meshes == 0");
    if (radius == 0 || radius > RTR_MAX_SEARCH_RADIUS)
        return rtr_set_error(ctx, RTR_E_INVALID, "bvh_build: search radius %u outside [1,%u]", radius, RTR_MAX_SEARCH_RADIUS);
    if (n >= (1u << 27)) return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "bvh_build: %u triangles >= 2^27", n);
    if (*out && (*out)->ctx != ctx) return rtr_set_error(ctx, RTR_E_INVALID, "bvh_build: *out belongs to another ctx");
    return RTR_OK;
}

int bvh_build_common(rtr_ctx* ctx, const rtr_triangle* tris_dev, uint32_t n, uint32_t array_len,
                     const rtr_mesh* meshes_dev, uint32_t nb_meshes, uint32_t radius, rtr_bvh* b) {
    RTR_CHECK(bvh_reserve(b, n));
    RTR_CHECK(rtr_ws_reserve(ctx, rtr_sort_ws_bytes(n, 4, true)));
    b->n = n; b->array_len = array_len; b->nb_meshes = nb_meshes; b->radius = radius;
    b->tris = tris_dev; b->meshes = meshes_dev;
    RTR_CUDA(ctx, cudaMemsetAsync(b->tile_status, 0, sizeof(uint64_t) * ((n + kPlocTile - 1) / kPlocTile + 1), ctx->stream));
    return rtr_bvh_run_build(b);
}

template <typename F>
int with_bvh(rtr_bvh* b, bool need_build_arrays, F&& f) {
    if (!b) return RTR_E_INVALID;
    RTR_CHECK(rtr_bvh_finish(b));
    if (!b->built) return rtr_set_error(b->ctx, RTR_E_STATE, "BVH has not been built");
    if (need_build_arrays && b->adopted)
        return rtr_set_error(b->ctx, RTR_E_STATE, "adopted BVH has no build arrays (flat nodes only)");
    return f();
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host-pointer wrappers use the front of the ctx workspace as staging, the device entry points
// take the rest
// ---------------------------------------------------------------------------------------
struct Staging {
    rtr_ctx* ctx;
    void* saved_ws; size_t saved_bytes;
    char* base; size_t off;
};
// Reserve `front` bytes of staging plus `back` bytes that remain visible as ctx->ws to callees.
static int staging_begin(rtr_ctx* ctx, size_t front, size_t back, Staging* st) {
    front = rtr_align_up(front, 256);
    RTR_CHECK(rtr_ws_reserve(ctx, front + back + 256));
    st->ctx = ctx; st->saved_ws = ctx->ws; st->saved_bytes = ctx->ws_bytes;
    st->base = static_cast<char*>(ctx->ws); st->off = 0;
    ctx->ws = st->base + front;
    ctx->ws_bytes = st->saved_bytes - front;
    return RTR_OK;
}
static void* staging_take(Staging* st, size_t bytes) {
    void* p = st->base + st->off;
    st->off = rtr_align_up(st->off + bytes, 256);
    return p;
}
static void staging_end(Staging* st) { st->ctx->ws = st->saved_ws; st->ctx->ws_bytes = st->saved_bytes; }

template <typename KeyT>
static int sort_host(rtr_ctx* ctx, KeyT* keys, uint32_t* vals, uint32_t n) {
    if (!ctx) return RTR_E_INVALID;
    if (n && !keys) return rtr_set_error(ctx, RTR_E_INVALID, "sort: NULL keys");
    if (n >= (1u << 30)) return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "sort: n=%u >= 2^30", n);
    if (n <= 1) return RTR_OK;
    const size_t kb = (size_t)n * sizeof(KeyT), vb = vals ? (size_t)n * 4 : 0;
    Staging st;
    RTR_CHECK(staging_begin(ctx, rtr_align_up(kb, 256) + rtr_align_up(vb, 256) + 512,
                            rtr_sort_ws_bytes(n, (int)sizeof(KeyT), vals != nullptr), &st));
    KeyT* d_keys = static_cast<KeyT*>(staging_take(&st, kb));
    uint32_t* d_vals = vals ? static_cast<uint32_t*>(staging_take(&st, vb)) : nullptr;
    auto body = [&]() -> int {
        RTR_CUDA(ctx, cudaMemcpyAsync(d_keys, keys, kb, cudaMemcpyHostToDevice, ctx->stream));
        if (vals) RTR_CUDA(ctx, cudaMemcpyAsync(d_vals, vals, vb, cudaMemcpyHostToDevice, ctx->stream));
        if (sizeof(KeyT) == 4) RTR_CHECK(rtr_sort_impl_u32(ctx, reinterpret_cast<uint32_t*>(d_keys), d_vals, n, 0, 32));
        else RTR_CHECK(rtr_sort_impl_u64(ctx, reinterpret_cast<uint64_t*>(d_keys), d_vals, n, 0, 64));
        RTR_CUDA(ctx, cudaMemcpyAsync(keys, d_keys, kb, cudaMemcpyDeviceToHost, ctx->stream));
        if (vals) RTR_CUDA(ctx, cudaMemcpyAsync(vals, d_vals, vb, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    const int r = body();
    staging_end(&st);
    return r;
}

template <typename CodeT>
static int morton_host(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t array_len, const rtr_mesh* meshes,
                       uint32_t nb_meshes, CodeT* codes_out, float* bounds_out) {
    const size_t tb = (size_t)array_len * sizeof(rtr_triangle), mb = (size_t)nb_meshes * sizeof(rtr_mesh);
    const size_t cb = (size_t)n * sizeof(CodeT);
    Staging st;
    RTR_CHECK(staging_begin(ctx, tb + mb + cb + 2048, 0, &st));
    rtr_triangle* d_tris = static_cast<rtr_triangle*>(staging_take(&st, tb));
    rtr_mesh* d_meshes = static_cast<rtr_mesh*>(staging_take(&st, mb));
    CodeT* d_codes = static_cast<CodeT*>(staging_take(&st, cb));
    float* d_bounds = static_cast<float*>(staging_take(&st, 64));
    uint32_t* d_ordered = static_cast<uint32_t*>(staging_take(&st, 64));
    auto body = [&]() -> int {
        RTR_CUDA(ctx, cudaMemcpyAsync(d_tris, tris, tb, cudaMemcpyHostToDevice, ctx->stream));
        RTR_CUDA(ctx, cudaMemcpyAsync(d_meshes, meshes, mb, cudaMemcpyHostToDevice, ctx->stream));
        if (sizeof(CodeT) == 4)
            RTR_CHECK(rtr_morton_launch(ctx, d_tris, n, array_len, d_meshes, reinterpret_cast<uint32_t*>(d_codes), nullptr,
                                        nullptr, d_bounds, d_ordered));
        else
            RTR_CHECK(rtr_morton_launch(ctx, d_tris, n, array_len, d_meshes, nullptr, nullptr,
                                        reinterpret_cast<uint64_t*>(d_codes), d_bounds, d_ordered));
        if (codes_out && n) RTR_CUDA(ctx, cudaMemcpyAsync(codes_out, d_codes, cb, cudaMemcpyDeviceToHost, ctx->stream));
        if (bounds_out) RTR_CUDA(ctx, cudaMemcpyAsync(bounds_out, d_bounds, 48, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    const int r = body();
    staging_end(&st);
    return r;
}

extern "C" {

const char* rtr_version(void) { return "rtr_b200 0.1 (sm_100a)"; }

int rtr_ctx_create(int device, rtr_ctx** out) {
    if (!out) return RTR_E_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return rtr_set_error(nullptr, RTR_E_NODEVICE, "no CUDA device: %s (librtr_b200 has no CPU fallback)",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return rtr_set_error(nullptr, RTR_E_INVALID, "device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return rtr_set_error(nullptr, RTR_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return rtr_set_error(nullptr, RTR_E_NODEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device,
                             prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return rtr_set_error(nullptr, RTR_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    rtr_ctx* ctx = new (std::nothrow) rtr_ctx();
    if (!ctx) return RTR_E_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // The build gathers 32-byte cluster records and the traversal 64-byte child-pair records at random: a
    // smaller L2 fetch granularity keeps DRAM from delivering neighbours nobody asked for.  (A hint; the
    // streaming kernels touch whole lines anyway.)  RTR_L2_FETCH=0 leaves the device default.
    {
        const char* env = getenv("RTR_L2_FETCH");
        const long gran = env ? atol(env) : RTR_L2_FETCH_DEFAULT;
        if (gran > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
        cudaGetLastError();  // a hint: never fatal
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return rtr_set_error(nullptr, RTR_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    ctx->pinned_bytes = 4096;
    if ((e = cudaMallocHost(&ctx->pinned, ctx->pinned_bytes)) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return rtr_set_error(nullptr, RTR_E_NOMEM, "cudaMallocHost: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return RTR_OK;
}

int rtr_ctx_destroy(rtr_ctx* ctx) {
    if (!ctx) return RTR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rtr_comm_destroy(ctx);
    for (auto& r : ctx->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto& e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->sm_table) cudaFree(ctx->sm_table);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RTR_OK;
}

const char* rtr_last_error(const rtr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rtr_ctx_sync(rtr_ctx* ctx) {
    if (!ctx) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RTR_OK;
}
void* rtr_ctx_stream(rtr_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int rtr_ctx_set_stream(rtr_ctx* ctx, void* stream) {
    if (!ctx) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->owns_stream = false;
    return RTR_OK;
}
int rtr_ctx_switch_stream(rtr_ctx* ctx, void* stream) {
    if (!ctx) return RTR_E_INVALID;
    if (ctx->owns_stream && ctx->stream) {  // the ctx's own stream is retired first (this once, with a sync)
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaStreamDestroy(ctx->stream);
    }
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->owns_stream = false;
    return RTR_OK;
}
int rtr_ctx_reserve_sms(rtr_ctx* ctx, uint32_t sms) {
    if (!ctx) return RTR_E_INVALID;
    if ((int)sms >= ctx->sm_count) return rtr_set_error(ctx, RTR_E_INVALID, "reserve_sms: %u of %d SMs", sms, ctx->sm_count);
    ctx->reserved_sms = (int)sms;
    return RTR_OK;
}
// An SM partition for the rays: a green context (CUDA >= 12.4, driver API resolved through the runtime so the library
// does not link libcuda) holding `sms` SMs -- rounded up to the granularity the architecture partitions by, 8 SMs on
// sm_90+ -- with `n_streams` streams of its own.  Persistent traversal launches on those streams size themselves for
// the partition, several of them may be in flight at once (the rays of frame f+1 fill in as the long paths of frame
// f finish), and the SMs outside the partition stay free for the kernels of a concurrent collective.
#include <cuda.h>
int rtr_ctx_partition_sms(rtr_ctx* ctx, uint32_t sms, uint32_t n_streams, void** streams_out, uint32_t* sms_out) {
    if (!ctx) return RTR_E_INVALID;
    if (!streams_out || n_streams == 0 || n_streams > 8 || sms == 0 || (int)sms > ctx->sm_count)
        return rtr_set_error(ctx, RTR_E_INVALID, "partition_sms: %u SMs, %u streams", sms, n_streams);
    if (ctx->green_ctx) return rtr_set_error(ctx, RTR_E_STATE, "partition_sms: this ctx already has a partition");
    RTR_CUDA(ctx, cudaSetDevice(ctx->device));
    typedef CUresult (*GetDevRes)(CUdevice, CUdevResource*, CUdevResourceType);
    typedef CUresult (*SplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    typedef CUresult (*GenDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    typedef CUresult (*GreenCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*GreenStream)(CUstream*, CUgreenCtx, unsigned int, int);
    typedef CUresult (*DevGet)(CUdevice*, int);
    GetDevRes get_res = nullptr; SplitByCount split = nullptr; GenDesc gen = nullptr; GreenCreate create = nullptr;
    GreenStream mkstream = nullptr; DevGet dev_get = nullptr;
    struct { const char* name; void** fn; } syms[] = {
        {"cuDeviceGetDevResource", (void**)&get_res}, {"cuDevSmResourceSplitByCount", (void**)&split},
        {"cuDevResourceGenerateDesc", (void**)&gen}, {"cuGreenCtxCreate", (void**)&create},
        {"cuGreenCtxStreamCreate", (void**)&mkstream}, {"cuDeviceGet", (void**)&dev_get}};
    for (auto& sy : syms) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(sy.name, sy.fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*sy.fn)
            return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "partition_sms: the driver has no %s (green contexts need CUDA >= 12.4)", sy.name);
    }
#define RTR_CU(call)                                                                                               \
    do {                                                                                                           \
        CUresult _r = (call);                                                                                      \
        if (_r != CUDA_SUCCESS) return rtr_set_error(ctx, RTR_E_CUDA, "partition_sms: %s -> CUresult %d", #call, (int)_r); \
    } while (0)
    CUdevice dev;
    RTR_CU(dev_get(&dev, ctx->device));
    CUdevResource all, part, rest;
    RTR_CU(get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    unsigned int groups = 1;
    RTR_CU(split(&part, &groups, &all, &rest, 0, sms));
    if (groups != 1) return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "partition_sms: cannot carve %u SMs out of %u", sms, all.sm.smCount);
    CUdevResourceDesc desc;
    RTR_CU(gen(&desc, &part, 1));
    CUgreenCtx g;
    RTR_CU(create(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    ctx->green_ctx = g;
    ctx->partition_sms = (int)part.sm.smCount;
    for (uint32_t i = 0; i < n_streams; ++i) {
        CUstream st;
        RTR_CU(mkstream(&st, g, CU_STREAM_NON_BLOCKING, 0));
        ctx->partition_streams.push_back(reinterpret_cast<cudaStream_t>(st));
        streams_out[i] = st;
    }
#undef RTR_CU
    if (sms_out) *sms_out = part.sm.smCount;
    return RTR_OK;
}
int rtr_ctx_device(const rtr_ctx* ctx) { return ctx ? ctx->device : -1; }
int rtr_ctx_sm_count(const rtr_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t rtr_ctx_launch_count(const rtr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rtr_ctx_profile_enable(rtr_ctx* ctx, int enable) {
    if (!ctx) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& r : ctx->prof) { ctx->prof_pool.push_back(r.e0); ctx->prof_pool.push_back(r.e1); }
    ctx->prof.clear();
    ctx->prof_open = false;
    ctx->profiling = enable != 0;
    return RTR_OK;
}

int rtr_ctx_profile_read(rtr_ctx* ctx, char* names, size_t names_bytes, float* total_ms, uint32_t* counts,
                         uint32_t capacity, uint32_t* n_out) {
    if (!ctx || !n_out) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<const char*> keys;
    std::vector<float> ms;
    std::vector<uint32_t> cnt;
    for (auto& r : ctx->prof) {
        float t = 0.f;
        RTR_CUDA(ctx, cudaEventElapsedTime(&t, r.e0, r.e1));
        size_t k = 0;
        for (; k < keys.size(); ++k) if (strcmp(keys[k], r.name) == 0) break;
        if (k == keys.size()) { keys.push_back(r.name); ms.push_back(0.f); cnt.push_back(0); }
        ms[k] += t; cnt[k] += 1;
    }
    *n_out = (uint32_t)keys.size();
    size_t off = 0;
    if (names && names_bytes) names[0] = 0;
    for (size_t k = 0; k < keys.size() && k < capacity; ++k) {
        if (total_ms) total_ms[k] = ms[k];
        if (counts) counts[k] = cnt[k];
        if (names) {
            const size_t len = strlen(keys[k]);
            if (off + len + 2 <= names_bytes) { memcpy(names + off, keys[k], len); off += len; names[off++] = '\n'; names[off] = 0; }
        }
    }
    return RTR_OK;
}

int rtr_host_alloc(size_t bytes, void** out) {
    if (!out) return RTR_E_INVALID;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return rtr_set_error(nullptr, RTR_E_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    return RTR_OK;
}
int rtr_host_free(void* p) {
    if (p) cudaFreeHost(p);
    return RTR_OK;
}
// page-locks memory the caller owns (e.g. a file mapping shared by the processes of a multi-GPU job), so that the
// asynchronous copies can target it
int rtr_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return RTR_E_INVALID;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) return rtr_set_error(nullptr, RTR_E_NOMEM, "cudaHostRegister(%zu): %s", bytes, cudaGetErrorString(e));
    return RTR_OK;
}
int rtr_host_unregister(void* p) {
    if (p) cudaHostUnregister(p);
    return RTR_OK;
}
int rtr_dev_alloc(rtr_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return RTR_E_INVALID;
    *out = nullptr;
    RTR_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return RTR_OK;
}
int rtr_dev_free(rtr_ctx* ctx, void* p) {
    if (!ctx) return RTR_E_INVALID;
    if (p) {
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        RTR_CUDA(ctx, cudaFree(p));
    }
    return RTR_OK;
}
int rtr_dev_upload(rtr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (bytes && (!dst || !src))) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RTR_OK;
}
int rtr_dev_download(rtr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (bytes && (!dst || !src))) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RTR_OK;
}
int rtr_dev_upload_async(rtr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (bytes && (!dst || !src))) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return RTR_OK;
}
int rtr_dev_download_async(rtr_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (bytes && (!dst || !src))) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return RTR_OK;
}
int rtr_dev_zero(rtr_ctx* ctx, void* dst, size_t bytes) {
    if (!ctx || (bytes && !dst)) return RTR_E_INVALID;
    RTR_CUDA(ctx, cudaMemsetAsync(dst, 0, bytes, ctx->stream));
    return RTR_OK;
}

int rtr_bit_histogram32_dev(rtr_ctx* ctx, const uint32_t* keys_dev, uint32_t n, uint32_t* out_dev) {
    if (!ctx || !out_dev || (n && !keys_dev)) return ctx ? rtr_set_error(ctx, RTR_E_INVALID, "bit_histogram32: NULL argument") : RTR_E_INVALID;
    return rtr_bit_histogram32_launch(ctx, keys_dev, n, out_dev);
}
int rtr_bit_histogram32(rtr_ctx* ctx, const uint32_t* keys, uint32_t n, uint32_t out[32]) {
    if (!ctx || !out || (n && !keys)) return ctx ? rtr_set_error(ctx, RTR_E_INVALID, "bit_histogram32: NULL argument") : RTR_E_INVALID;
    Staging st;
    RTR_CHECK(staging_begin(ctx, (size_t)n * 4 + 1024, 0, &st));
    uint32_t* d_keys = static_cast<uint32_t*>(staging_take(&st, (size_t)n * 4));
    uint32_t* d_out = static_cast<uint32_t*>(staging_take(&st, 128));
    int r = RTR_OK;
    auto body = [&]() -> int {
        if (n) RTR_CUDA(ctx, cudaMemcpyAsync(d_keys, keys, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        RTR_CUDA(ctx, cudaMemsetAsync(d_out, 0, 128, ctx->stream));
        RTR_CHECK(rtr_bit_histogram32_launch(ctx, d_keys, n, d_out));
        RTR_CUDA(ctx, cudaMemcpyAsync(out, d_out, 128, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    r = body();
    staging_end(&st);
    return r;
}

int rtr_digitplace_exclusive_scan_dev(rtr_ctx* ctx, const uint32_t* in_dev, uint32_t* out_dev) {
    if (!ctx || !in_dev || !out_dev) return ctx ? rtr_set_error(ctx, RTR_E_INVALID, "digitplace_scan: NULL argument") : RTR_E_INVALID;
    return rtr_digitplace_scan_launch(ctx, in_dev, out_dev);
}
int rtr_digitplace_exclusive_scan(rtr_ctx* ctx, const uint32_t in[32], uint32_t out[32]) {
    if (!ctx || !in || !out) return ctx ? rtr_set_error(ctx, RTR_E_INVALID, "digitplace_scan: NULL argument") : RTR_E_INVALID;
    Staging st;
    RTR_CHECK(staging_begin(ctx, 1024, 0, &st));
    uint32_t* d_in = static_cast<uint32_t*>(staging_take(&st, 128));
    uint32_t* d_out = static_cast<uint32_t*>(staging_take(&st, 128));
    auto body = [&]() -> int {
        RTR_CUDA(ctx, cudaMemcpyAsync(d_in, in, 128, cudaMemcpyHostToDevice, ctx->stream));
        RTR_CHECK(rtr_digitplace_scan_launch(ctx, d_in, d_out));
        RTR_CUDA(ctx, cudaMemcpyAsync(out, d_out, 128, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    const int r = body();
    staging_end(&st);
    return r;
}

int rtr_sort_keys_u32(rtr_ctx* ctx, uint32_t* keys, uint32_t n) { return sort_host<uint32_t>(ctx, keys, nullptr, n); }
int rtr_sort_pairs_u32(rtr_ctx* ctx, uint32_t* keys, uint32_t* values, uint32_t n) {
    if (ctx && n && !values) return rtr_set_error(ctx, RTR_E_INVALID, "sort_pairs: NULL values");
    return sort_host<uint32_t>(ctx, keys, values, n);
}
int rtr_sort_keys_u64(rtr_ctx* ctx, uint64_t* keys, uint32_t n) { return sort_host<uint64_t>(ctx, keys, nullptr, n); }
int rtr_sort_pairs_u64(rtr_ctx* ctx, uint64_t* keys, uint32_t* values, uint32_t n) {
    if (ctx && n && !values) return rtr_set_error(ctx, RTR_E_INVALID, "sort_pairs: NULL values");
    return sort_host<uint64_t>(ctx, keys, values, n);
}
int rtr_sort_pairs_u32_dev(rtr_ctx* ctx, uint32_t* keys_dev, uint32_t* values_dev, uint32_t n, int begin_bit, int end_bit) {
    if (!ctx) return RTR_E_INVALID;
    if (n && !keys_dev) return rtr_set_error(ctx, RTR_E_INVALID, "sort: NULL keys");
    return rtr_sort_impl_u32(ctx, keys_dev, values_dev, n, begin_bit, end_bit);
}
int rtr_sort_pairs_u64_dev(rtr_ctx* ctx, uint64_t* keys_dev, uint32_t* values_dev, uint32_t n, int begin_bit, int end_bit) {
    if (!ctx) return RTR_E_INVALID;
    if (n && !keys_dev) return rtr_set_error(ctx, RTR_E_INVALID, "sort: NULL keys");
    return rtr_sort_impl_u64(ctx, keys_dev, values_dev, n, begin_bit, end_bit);
}

// ---- Morton ----
static int morton_args_ok(rtr_ctx* ctx, const void* tris, uint32_t n, uint32_t array_len, const void* meshes,
                          uint32_t nb_meshes, const void* out) {
    if (!ctx) return RTR_E_INVALID;
    if (!tris || !meshes || (n && !out)) return rtr_set_error(ctx, RTR_E_INVALID, "morton: NULL argument");
    if (array_len < n || array_len == 0) return rtr_set_error(ctx, RTR_E_INVALID, "morton: tris_array_len %u < n %u (or 0)", array_len, n);
    if (nb_meshes == 0) return rtr_set_error(ctx, RTR_E_INVALID, "morton: nb_meshes == 0");
    return RTR_OK;
}

int rtr_morton_codes_dev(rtr_ctx* ctx, const rtr_triangle* tris_dev, uint32_t n, uint32_t array_len,
                         const rtr_mesh* meshes_dev, uint32_t nb_meshes, uint32_t* codes_dev) {
    RTR_CHECK(morton_args_ok(ctx, tris_dev, n, array_len, meshes_dev, nb_meshes, codes_dev));
    Staging st;
    RTR_CHECK(staging_begin(ctx, 1024, 0, &st));
    float* bounds = static_cast<float*>(staging_take(&st, 64));
    uint32_t* ordered = static_cast<uint32_t*>(staging_take(&st, 64));
    const int r = rtr_morton_launch(ctx, tris_dev, n, array_len, meshes_dev, codes_dev, nullptr, nullptr, bounds, ordered);
    staging_end(&st);
    return r;
}

int rtr_morton_codes(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t array_len, const rtr_mesh* meshes,
                     uint32_t nb_meshes, uint32_t* codes_out) {
    RTR_CHECK(morton_args_ok(ctx, tris, n, array_len, meshes, nb_meshes, codes_out));
    return morton_host<uint32_t>(ctx, tris, n, array_len, meshes, nb_meshes, codes_out, nullptr);
}
int rtr_morton_codes64(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t array_len, const rtr_mesh* meshes,
                       uint32_t nb_meshes, uint64_t* codes_out) {
    RTR_CHECK(morton_args_ok(ctx, tris, n, array_len, meshes, nb_meshes, codes_out));
    return morton_host<uint64_t>(ctx, tris, n, array_len, meshes, nb_meshes, codes_out, nullptr);
}
int rtr_scene_bounds(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t array_len, const rtr_mesh* meshes,
                     uint32_t nb_meshes, float out[12]) {
    RTR_CHECK(morton_args_ok(ctx, tris, 0, array_len, meshes, nb_meshes, out));
    if (!out) return rtr_set_error(ctx, RTR_E_INVALID, "scene_bounds: NULL out");
    return morton_host<uint32_t>(ctx, tris, 0, array_len, meshes, nb_meshes, nullptr, out);
}

// ---- BVH ----
static int build_dev_impl(rtr_ctx* ctx, const rtr_triangle* tris_dev, uint32_t n, uint32_t array_len,
                          const rtr_mesh* meshes_dev, uint32_t nb_meshes, uint32_t radius, uint32_t key_bits, rtr_bvh** out) {
    RTR_CHECK(check_build_args(ctx, tris_dev, n, array_len, meshes_dev, nb_meshes, radius, out));
    rtr_bvh* b = *out;
    if (!b) {
        b = new (std::nothrow) rtr_bvh();
        if (!b) return rtr_set_error(ctx, RTR_E_NOMEM, "bvh_build: host allocation failed");
        b->ctx = ctx;
    }
    b->key_bits = key_bits;
    const int r = bvh_build_common(ctx, tris_dev, n, array_len, meshes_dev, nb_meshes, radius, b);
    if (r != RTR_OK && !*out) { rtr_bvh_destroy(b); return r; }
    *out = b;
    return r;
}

int rtr_bvh_build_dev(rtr_ctx* ctx, const rtr_triangle* tris_dev, uint32_t n, uint32_t array_len,
                      const rtr_mesh* meshes_dev, uint32_t nb_meshes, uint32_t radius, rtr_bvh** out) {
    return build_dev_impl(ctx, tris_dev, n, array_len, meshes_dev, nb_meshes, radius, 32, out);
}
int rtr_bvh_build64_dev(rtr_ctx* ctx, const rtr_triangle* tris_dev, uint32_t n, uint32_t array_len,
                        const rtr_mesh* meshes_dev, uint32_t nb_meshes, uint32_t radius, rtr_bvh** out) {
    return build_dev_impl(ctx, tris_dev, n, array_len, meshes_dev, nb_meshes, radius, 64, out);
}

static int build_host_impl(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t array_len, const rtr_mesh* meshes,
                           uint32_t nb_meshes, uint32_t radius, uint32_t key_bits, rtr_bvh** out) {
    RTR_CHECK(check_build_args(ctx, tris, n, array_len, meshes, nb_meshes, radius, out));
    rtr_bvh* b = *out;
    const bool fresh = (b == nullptr);
    if (fresh) {
        b = new (std::nothrow) rtr_bvh();
        if (!b) return rtr_set_error(ctx, RTR_E_NOMEM, "bvh_build: host allocation failed");
        b->ctx = ctx;
    }
    b->key_bits = key_bits;
    auto body = [&]() -> int {
        // like the reference ctor, keep private copies of both vectors (bvh.cpp:16-17)
        if (b->tris_own_cap < array_len) {
            if (b->tris_own) { RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(b->tris_own); b->tris_own = nullptr; b->tris_own_cap = 0; }
            RTR_CHECK(dev_alloc(ctx, &b->tris_own, array_len));
            b->tris_own_cap = array_len;
        }
        if (b->meshes_own_cap < nb_meshes) {
            if (b->meshes_own) { RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(b->meshes_own); b->meshes_own = nullptr; b->meshes_own_cap = 0; }
            RTR_CHECK(dev_alloc(ctx, &b->meshes_own, nb_meshes));
            b->meshes_own_cap = nb_meshes;
        }
        RTR_CUDA(ctx, cudaMemcpyAsync(b->tris_own, tris, (size_t)array_len * sizeof(rtr_triangle), cudaMemcpyHostToDevice, ctx->stream));
        RTR_CUDA(ctx, cudaMemcpyAsync(b->meshes_own, meshes, (size_t)nb_meshes * sizeof(rtr_mesh), cudaMemcpyHostToDevice, ctx->stream));
        RTR_CHECK(bvh_build_common(ctx, b->tris_own, n, array_len, b->meshes_own, nb_meshes, radius, b));
        return rtr_bvh_finish(b);  // synchronous like the reference ctor (bvh.cpp:22): waits and reports
    };
    const int r = body();
    if (r != RTR_OK && fresh) { rtr_bvh_destroy(b); return r; }
    *out = b;
    return r;
}

int rtr_bvh_build(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t array_len, const rtr_mesh* meshes,
                  uint32_t nb_meshes, uint32_t radius, rtr_bvh** out) {
    return build_host_impl(ctx, tris, n, array_len, meshes, nb_meshes, radius, 32, out);
}
int rtr_bvh_build64(rtr_ctx* ctx, const rtr_triangle* tris, uint32_t n, uint32_t array_len, const rtr_mesh* meshes,
                    uint32_t nb_meshes, uint32_t radius, rtr_bvh** out) {
    return build_host_impl(ctx, tris, n, array_len, meshes, nb_meshes, radius, 64, out);
}

int rtr_bvh_adopt_dev(rtr_ctx* ctx, const rtr_node* nodes_dev, uint32_t nb_triangles, const rtr_triangle* tris_dev,
                      const rtr_mesh* meshes_dev, uint32_t nb_meshes, rtr_bvh** out) {
    if (!ctx) return RTR_E_INVALID;
    if (!out || !nodes_dev || !tris_dev || !meshes_dev || nb_triangles == 0 || nb_meshes == 0)
        return rtr_set_error(ctx, RTR_E_INVALID, "bvh_adopt: NULL or empty argument");
    rtr_bvh* b = *out;
    const bool fresh = (b == nullptr);
    if (fresh) {
        b = new (std::nothrow) rtr_bvh();
        if (!b) return rtr_set_error(ctx, RTR_E_NOMEM, "bvh_adopt: host allocation failed");
        b->ctx = ctx;
    }
    auto body = [&]() -> int {
        if (!b->tparams) RTR_CHECK(dev_alloc(ctx, &b->tparams, 1));
        b->built = false;
        b->n = nb_triangles; b->array_len = nb_triangles; b->nb_meshes = nb_meshes;
        b->tris = tris_dev; b->meshes = meshes_dev; b->flat_view = nodes_dev;
        b->adopted = true; b->trav_only = false;
        RTR_CHECK(rtr_bvh_compute_trace_params(b));
        b->built = true;
        return RTR_OK;
    };
    const int r = body();
    if (r != RTR_OK && fresh) { rtr_bvh_destroy(b); return r; }
    *out = b;
    return r;
}

int rtr_bvh_destroy(rtr_bvh* b) {
    if (!b) return RTR_OK;
    if (b->ctx) { cudaSetDevice(b->ctx->device); cudaStreamSynchronize(b->ctx->stream); }
    for (auto& e : b->ev) if (e) cudaEventDestroy(e);
    bvh_free_arrays(b);
    delete b;
    return RTR_OK;
}

uint32_t rtr_bvh_nb_triangles(const rtr_bvh* b) { return b ? b->n : 0; }
uint32_t rtr_bvh_nb_nodes(const rtr_bvh* b) { return (b && b->n) ? 2 * b->n - 1 : 0; }

int rtr_bvh_enable_stage_timing(rtr_bvh* b, int enable) {
    if (!b) return RTR_E_INVALID;
    if (enable)
        for (auto& e : b->ev)
            if (!e) RTR_CUDA(b->ctx, cudaEventCreate(&e));
    b->timing = enable != 0;
    return RTR_OK;
}

int rtr_bvh_stage_ms(rtr_bvh* b, float out[6]) {
    if (!b || !out) return RTR_E_INVALID;
    RTR_CHECK(rtr_bvh_finish(b));
    if (!b->timing || !b->built || b->adopted) return rtr_set_error(b->ctx, RTR_E_STATE, "stage timing not enabled for the last build");
    RTR_CUDA(b->ctx, cudaEventSynchronize(b->ev[5]));
    for (int i = 0; i < 5; ++i) RTR_CUDA(b->ctx, cudaEventElapsedTime(&out[i], b->ev[i], b->ev[i + 1]));
    RTR_CUDA(b->ctx, cudaEventElapsedTime(&out[5], b->ev[0], b->ev[5]));
    return RTR_OK;
}

int rtr_bvh_iteration_trace(rtr_bvh* b, uint32_t* active, uint32_t* merges, uint32_t capacity, uint32_t* count) {
    return with_bvh(b, true, [&]() -> int {
        rtr_ctx* ctx = b->ctx;
        if (count) *count = b->iterations;
        const uint32_t m = b->iterations < capacity ? b->iterations : capacity;
        if (m && active) RTR_CUDA(ctx, cudaMemcpyAsync(active, b->trace_active, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (m && merges) RTR_CUDA(ctx, cudaMemcpyAsync(merges, b->trace_merges, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    });
}

// diagnostics: %globaltimer (ns) at the start of every PLOC iteration of the last build, count + 1 values
int rtr_bvh_iteration_times(rtr_bvh* b, uint64_t* start_ns, uint32_t capacity, uint32_t* count) {
    return with_bvh(b, true, [&]() -> int {
        rtr_ctx* ctx = b->ctx;
        if (count) *count = b->iterations;
        const uint32_t m = (b->iterations + 1) < capacity ? (b->iterations + 1) : capacity;
        if (m && start_ns) RTR_CUDA(ctx, cudaMemcpyAsync(start_ns, b->iter_ns, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    });
}

int rtr_bvh_morton_codes(rtr_bvh* b, uint32_t* out) {
    return with_bvh(b, true, [&]() -> int {
        if (!out) return rtr_set_error(b->ctx, RTR_E_INVALID, "NULL out");
        if (b->key_bits == 64) return rtr_set_error(b->ctx, RTR_E_STATE, "built over 64-bit keys: use rtr_bvh_morton_codes64");
        return rtr_dev_download(b->ctx, out, b->codes, (size_t)b->n * 4);
    });
}
int rtr_bvh_morton_codes64(rtr_bvh* b, uint64_t* out) {
    return with_bvh(b, true, [&]() -> int {
        if (!out) return rtr_set_error(b->ctx, RTR_E_INVALID, "NULL out");
        if (b->key_bits != 64) return rtr_set_error(b->ctx, RTR_E_STATE, "built over 32-bit keys: use rtr_bvh_morton_codes");
        return rtr_dev_download(b->ctx, out, b->codes64, (size_t)b->n * 8);
    });
}
int rtr_bvh_triangle_indices(rtr_bvh* b, uint32_t* out) {
    return with_bvh(b, true, [&]() -> int {
        if (!out) return rtr_set_error(b->ctx, RTR_E_INVALID, "NULL out");
        return rtr_dev_download(b->ctx, out, b->tri_idx, (size_t)b->n * 4);
    });
}

int rtr_bvh_clusters(rtr_bvh* b, rtr_node* clusters, uint32_t* parent, uint32_t* left, uint32_t* right, uint8_t* is_leaf) {
    return with_bvh(b, true, [&]() -> int {
        rtr_ctx* ctx = b->ctx;
        const size_t nc = 2 * (size_t)b->n - 1;
        Staging st;
        RTR_CHECK(staging_begin(ctx, nc * (sizeof(rtr_node) + 13) + 4096, 0, &st));
        rtr_node* d_cl = static_cast<rtr_node*>(staging_take(&st, nc * sizeof(rtr_node)));
        uint32_t* d_par = static_cast<uint32_t*>(staging_take(&st, nc * 4));
        uint32_t* d_l = static_cast<uint32_t*>(staging_take(&st, nc * 4));
        uint32_t* d_r = static_cast<uint32_t*>(staging_take(&st, nc * 4));
        uint8_t* d_leaf = static_cast<uint8_t*>(staging_take(&st, nc));
        auto body = [&]() -> int {
            RTR_CHECK(rtr_bvh_export_clusters(b, d_cl, d_par, d_l, d_r, d_leaf));
            if (clusters) RTR_CUDA(ctx, cudaMemcpyAsync(clusters, d_cl, nc * sizeof(rtr_node), cudaMemcpyDeviceToHost, ctx->stream));
            if (parent) RTR_CUDA(ctx, cudaMemcpyAsync(parent, d_par, nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (left) RTR_CUDA(ctx, cudaMemcpyAsync(left, d_l, nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (right) RTR_CUDA(ctx, cudaMemcpyAsync(right, d_r, nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (is_leaf) RTR_CUDA(ctx, cudaMemcpyAsync(is_leaf, d_leaf, nc, cudaMemcpyDeviceToHost, ctx->stream));
            RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return RTR_OK;
        };
        const int r = body();
        staging_end(&st);
        return r;
    });
}

int rtr_bvh_flat_nodes(rtr_bvh* b, rtr_node* out) {
    return with_bvh(b, false, [&]() -> int {
        if (!out) return rtr_set_error(b->ctx, RTR_E_INVALID, "NULL out");
        if (b->trav_only) return rtr_set_error(b->ctx, RTR_E_STATE, "this BVH holds traversal records only (rtr_bvh_broadcast_traversal)");
        return rtr_dev_download(b->ctx, out, b->flat_view, (2 * (size_t)b->n - 1) * sizeof(rtr_node));
    });
}
const rtr_node* rtr_bvh_device_nodes(const rtr_bvh* b) { return (b && b->built && !b->trav_only) ? b->flat_view : nullptr; }
const rtr_triangle* rtr_bvh_device_triangles(const rtr_bvh* b) { return (b && b->built) ? b->tris : nullptr; }
const rtr_mesh* rtr_bvh_device_meshes(const rtr_bvh* b) { return (b && b->built) ? b->meshes : nullptr; }

// ---- traversal ----
static int trace_args_ok(rtr_ctx* ctx, const rtr_bvh* b, const void* cam_or_rays) {
    if (!ctx) return RTR_E_INVALID;
    if (!b || !b->built) return rtr_set_error(ctx, RTR_E_STATE, "trace: BVH not built");
    if (b->ctx != ctx) return rtr_set_error(ctx, RTR_E_INVALID, "trace: BVH belongs to another ctx");
    if (!cam_or_rays) return rtr_set_error(ctx, RTR_E_INVALID, "trace: NULL camera/rays");
    return RTR_OK;
}
// the by-the-letter order walks the 48-byte nodes, which a traversal-only replica does not have
static int order_ok(rtr_ctx* ctx, const rtr_bvh* b, uint32_t flags) {
    if ((flags & (RTR_TRACE_REFERENCE_ORDER | RTR_TRACE_DEEP_STACK)) && b->trav_only)
        return rtr_set_error(ctx, RTR_E_STATE, "trace: RTR_TRACE_REFERENCE_ORDER / RTR_TRACE_DEEP_STACK need the flat nodes; this BVH holds traversal records only");
    return RTR_OK;
}

// A ray that runs out of traversal stack (128 entries; the shader's is 1024 deep, raytracer.glsl:251) drops subtrees:
// the kernels count such rays in TraceParams::stack_overflows.  The host-pointer entry points synchronise anyway and
// refuse to hand such a frame back; callers of the asynchronous _dev forms ask with rtr_bvh_stack_overflows.
static int fetch_overflows(rtr_ctx* ctx, const rtr_bvh* b, uint32_t* host_count) {
    RTR_CUDA(ctx, cudaMemcpyAsync(host_count, &b->tparams->stack_overflows, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return RTR_OK;
}
static int refuse_overflows(rtr_ctx* ctx, const rtr_bvh* b, uint32_t count, const char* what, uint32_t flags) {
    if (count == 0u) return RTR_OK;
    cudaMemsetAsync(&b->tparams->stack_overflows, 0, sizeof(uint32_t), ctx->stream);
    return rtr_set_error(ctx, RTR_E_UNSUPPORTED, "%s: %u ray(s) ran out of traversal stack (%d entries%s); the frame is incomplete", what,
                         count, (flags & RTR_TRACE_DEEP_STACK) ? 1024 : 128,
                         (flags & RTR_TRACE_DEEP_STACK) ? ", the depth of the shader's own stack, raytracer.glsl:251" : "");
}
// The host-pointer calls run `attempt(flags, &overflows)` (launch + copies + stream sync).  If rays ran out of the 128
// lane-stack entries, the call is traced once more with the shader's own 1024-entry stack (same records by
// construction); a traversal-only replica has no flat nodes to do that with and fails right away.
extern "C++" {
template <typename F>
static int trace_with_fallback(rtr_ctx* ctx, const rtr_bvh* b, uint32_t flags, const char* what, F&& attempt) {
    uint32_t overflows = 0;
    RTR_CHECK(attempt(flags, &overflows));
    if (overflows != 0u && !(flags & RTR_TRACE_DEEP_STACK) && !b->trav_only) {
        RTR_CUDA(ctx, cudaMemsetAsync(&b->tparams->stack_overflows, 0, sizeof(uint32_t), ctx->stream));
        flags |= RTR_TRACE_DEEP_STACK;
        overflows = 0;
        RTR_CHECK(attempt(flags, &overflows));
    }
    return refuse_overflows(ctx, b, overflows, what, flags);
}
}  // extern "C++"

int rtr_bvh_stack_overflows(const rtr_bvh* b, uint32_t* count_out) {
    if (!b) return RTR_E_INVALID;
    rtr_ctx* ctx = b->ctx;
    if (!count_out) return rtr_set_error(ctx, RTR_E_INVALID, "stack_overflows: NULL output");
    RTR_CHECK(rtr_bvh_finish(const_cast<rtr_bvh*>(b)));
    if (!b->built || !b->tparams) return rtr_set_error(ctx, RTR_E_STATE, "stack_overflows: the BVH has not been built");
    RTR_CHECK(fetch_overflows(ctx, b, count_out));
    RTR_CUDA(ctx, cudaMemsetAsync(&b->tparams->stack_overflows, 0, sizeof(uint32_t), ctx->stream));
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RTR_OK;
}

int rtr_trace_primary_dev(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera* cam, uint32_t width, uint32_t height,
                          uint32_t denom_w, uint32_t denom_h, uint32_t row0, uint32_t row1, uint32_t flags, rtr_hit* hits_dev) {
    RTR_CHECK(trace_args_ok(ctx, b, cam));
    RTR_CHECK(order_ok(ctx, b, flags));
    if (!hits_dev) return rtr_set_error(ctx, RTR_E_INVALID, "trace_primary: NULL output");
    return rtr_trace_primary_launch(ctx, b, *cam, width, height, denom_w, denom_h, row0, row1, flags, hits_dev);
}

int rtr_trace_primary(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera* cam, uint32_t width, uint32_t height,
                      uint32_t denom_w, uint32_t denom_h, uint32_t flags, rtr_hit* hits_out) {
    if (b) RTR_CHECK(rtr_bvh_finish(const_cast<rtr_bvh*>(b)));  // a build submitted with rtr_bvh_build_dev may still owe its verdict
    RTR_CHECK(trace_args_ok(ctx, b, cam));
    RTR_CHECK(order_ok(ctx, b, flags));
    if (!hits_out) return rtr_set_error(ctx, RTR_E_INVALID, "trace_primary: NULL output");
    const size_t bytes = (size_t)width * height * sizeof(rtr_hit);
    Staging st;
    RTR_CHECK(staging_begin(ctx, bytes + 512, 0, &st));
    rtr_hit* d_hits = static_cast<rtr_hit*>(staging_take(&st, bytes));
    auto attempt = [&](uint32_t f, uint32_t* overflows) -> int {
        RTR_CHECK(rtr_trace_primary_launch(ctx, b, *cam, width, height, denom_w, denom_h, 0, height, f, d_hits));
        RTR_CUDA(ctx, cudaMemcpyAsync(hits_out, d_hits, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CHECK(fetch_overflows(ctx, b, overflows));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    const int r = trace_with_fallback(ctx, b, flags, "trace_primary", attempt);
    staging_end(&st);
    return r;
}

int rtr_trace_rays_dev(rtr_ctx* ctx, const rtr_bvh* b, const rtr_ray* rays_dev, uint64_t n_rays, int any_hit,
                       const float* t_max_dev, uint32_t flags, rtr_hit* hits_dev) {
    RTR_CHECK(trace_args_ok(ctx, b, n_rays ? (const void*)rays_dev : (const void*)b));
    RTR_CHECK(order_ok(ctx, b, flags));
    if (n_rays && !hits_dev) return rtr_set_error(ctx, RTR_E_INVALID, "trace_rays: NULL output");
    return rtr_trace_rays_launch(ctx, b, rays_dev, n_rays, any_hit, t_max_dev, flags, hits_dev);
}

int rtr_trace_rays(rtr_ctx* ctx, const rtr_bvh* b, const rtr_ray* rays, uint64_t n_rays, int any_hit, const float* t_max,
                   uint32_t flags, rtr_hit* hits_out) {
    if (b) RTR_CHECK(rtr_bvh_finish(const_cast<rtr_bvh*>(b)));  // a build submitted with rtr_bvh_build_dev may still owe its verdict
    RTR_CHECK(trace_args_ok(ctx, b, n_rays ? (const void*)rays : (const void*)b));
    RTR_CHECK(order_ok(ctx, b, flags));
    if (n_rays == 0) return RTR_OK;
    if (!hits_out) return rtr_set_error(ctx, RTR_E_INVALID, "trace_rays: NULL output");
    if (!(flags & RTR_TRACE_REFERENCE_ORDER)) {
        // The default order prunes with a bound on the rounding error of a computed hit distance that is derived for
        // the shader's rays, which are unit vectors (getRay normalises, raytracer.glsl:92-100); the error grows with |d|
        // (measured: 3 % of the bound at |d| = 1, 25 % at 10, beyond it near 40; tests/test_prune_bound_cpu.py).
        for (uint64_t i = 0; i < n_rays; ++i) {
            const float* d = rays[i].direction;
            const float d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            if (d2 > 64.f && d2 < INFINITY) {  // the batch goes through the shader's own order: any length, same records
                if (b->trav_only)
                    return rtr_set_error(ctx, RTR_E_UNSUPPORTED,
                                         "trace_rays: ray %llu has |direction| = %g; a traversal-only replica traces in the default order, "
                                         "which needs |direction| <= 8 (normalise it)", (unsigned long long)i, sqrtf(d2));
                flags |= RTR_TRACE_REFERENCE_ORDER;
                break;
            }
        }
    }
    const size_t rb = n_rays * sizeof(rtr_ray), hb = n_rays * sizeof(rtr_hit), tb = t_max ? n_rays * 4 : 0;
    Staging st;
    RTR_CHECK(staging_begin(ctx, rb + hb + tb + 1024, 0, &st));
    rtr_ray* d_rays = static_cast<rtr_ray*>(staging_take(&st, rb));
    rtr_hit* d_hits = static_cast<rtr_hit*>(staging_take(&st, hb));
    float* d_tmax = t_max ? static_cast<float*>(staging_take(&st, tb)) : nullptr;
    auto attempt = [&](uint32_t f, uint32_t* overflows) -> int {
        RTR_CUDA(ctx, cudaMemcpyAsync(d_rays, rays, rb, cudaMemcpyHostToDevice, ctx->stream));
        if (t_max) RTR_CUDA(ctx, cudaMemcpyAsync(d_tmax, t_max, tb, cudaMemcpyHostToDevice, ctx->stream));
        RTR_CHECK(rtr_trace_rays_launch(ctx, b, d_rays, n_rays, any_hit, d_tmax, f, d_hits));
        RTR_CUDA(ctx, cudaMemcpyAsync(hits_out, d_hits, hb, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CHECK(fetch_overflows(ctx, b, overflows));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    const int r = trace_with_fallback(ctx, b, flags, "trace_rays", attempt);
    staging_end(&st);
    return r;
}

int rtr_render_dev(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera* cam, uint32_t width, uint32_t height,
                   uint32_t denom_w, uint32_t denom_h, uint32_t row0, uint32_t row1, uint32_t bounces, int shadow,
                   const float light_pos[3], uint32_t flags, float* rgba_dev, rtr_hit* hits_dev, uint64_t* rays_dev) {
    RTR_CHECK(trace_args_ok(ctx, b, cam));
    RTR_CHECK(order_ok(ctx, b, flags));
    if (shadow && !light_pos) return rtr_set_error(ctx, RTR_E_INVALID, "render: shadow rays need a light position");
    return rtr_render_launch(ctx, b, *cam, width, height, denom_w, denom_h, row0, row1, bounces, shadow, light_pos, flags,
                             rgba_dev, hits_dev, rays_dev);
}

int rtr_render_sharded_dev(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera* cam, uint32_t width, uint32_t height,
                           uint32_t denom_w, uint32_t denom_h, uint32_t rows_per_block, uint32_t shard_rank,
                           uint32_t shard_count, uint32_t bounces, int shadow, const float light_pos[3], uint32_t flags,
                           float* rgba_dev, rtr_hit* hits_dev, uint64_t* rays_dev) {
    RTR_CHECK(trace_args_ok(ctx, b, cam));
    RTR_CHECK(order_ok(ctx, b, flags));
    if (shadow && !light_pos) return rtr_set_error(ctx, RTR_E_INVALID, "render: shadow rays need a light position");
    if (shard_count == 0 || shard_rank >= shard_count || shard_count > 255)
        return rtr_set_error(ctx, RTR_E_INVALID, "render: bad shard %u of %u", shard_rank, shard_count);
    const uint8_t off = (uint8_t)shard_rank;
    return rtr_render_launch(ctx, b, *cam, width, height, denom_w, denom_h, 0, height, bounces, shadow, light_pos, flags,
                             rgba_dev, hits_dev, rays_dev, rows_per_block, shard_count, 1u, &off);
}

int rtr_render_stripes_dev(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera* cam, uint32_t width, uint32_t height,
                           uint32_t denom_w, uint32_t denom_h, uint32_t rows_per_block, const uint32_t* stripes_of_rank,
                           uint32_t nranks, uint32_t rank, uint32_t bounces, int shadow, const float light_pos[3],
                           uint32_t flags, float* rgba_dev, rtr_hit* hits_dev, uint64_t* rays_dev) {
    RTR_CHECK(trace_args_ok(ctx, b, cam));
    RTR_CHECK(order_ok(ctx, b, flags));
    if (shadow && !light_pos) return rtr_set_error(ctx, RTR_E_INVALID, "render: shadow rays need a light position");
    if (!stripes_of_rank || nranks == 0 || rank >= nranks)
        return rtr_set_error(ctx, RTR_E_INVALID, "render: bad stripe layout (rank %u of %u)", rank, nranks);
    const std::vector<int> owner = rtr_stripe_owners(stripes_of_rank, (int)nranks);
    if (owner.empty() || owner.size() > 255) return rtr_set_error(ctx, RTR_E_INVALID, "render: %zu stripes (1..255)", owner.size());
    if (stripes_of_rank[rank] > RTR_MAX_STRIPES_PER_RANK)
        return rtr_set_error(ctx, RTR_E_INVALID, "render: more than %d stripes for one rank", RTR_MAX_STRIPES_PER_RANK);
    if (owner.size() == 1)  // a single stripe: the whole image (or nothing)
        return stripes_of_rank[rank] ? rtr_render_launch(ctx, b, *cam, width, height, denom_w, denom_h, 0, height, bounces, shadow,
                                                         light_pos, flags, rgba_dev, hits_dev, rays_dev)
                                     : RTR_OK;
    uint8_t off[RTR_MAX_STRIPES_PER_RANK];
    uint32_t mine = 0;
    for (size_t v = 0; v < owner.size(); ++v)
        if (owner[v] == (int)rank) off[mine++] = (uint8_t)v;
    return rtr_render_launch(ctx, b, *cam, width, height, denom_w, denom_h, 0, height, bounces, shadow, light_pos, flags,
                             rgba_dev, hits_dev, rays_dev, rows_per_block, (uint32_t)owner.size(), mine, off);
}

int rtr_shade_dev(rtr_ctx* ctx, const rtr_hit* hits_dev, uint64_t n, const rtr_triangle* tris_dev, const rtr_mesh* meshes_dev,
                  const rtr_material* materials_dev, uint32_t flags, const float* bvh_rgba_dev, float* rgba_dev) {
    if (!ctx) return RTR_E_INVALID;
    if (n && (!hits_dev || !tris_dev || !meshes_dev || !materials_dev || !rgba_dev))
        return rtr_set_error(ctx, RTR_E_INVALID, "shade: NULL argument");
    if (n && (flags & RTR_SHADE_BVH) && !bvh_rgba_dev) return rtr_set_error(ctx, RTR_E_INVALID, "shade: RTR_SHADE_BVH without overlay colours");
    return rtr_shade_launch(ctx, hits_dev, n, tris_dev, meshes_dev, materials_dev, flags, bvh_rgba_dev, rgba_dev);
}

int rtr_shade(rtr_ctx* ctx, const rtr_hit* hits, uint64_t n, const rtr_triangle* tris, uint32_t nb_triangles,
              const rtr_mesh* meshes, uint32_t nb_meshes, const rtr_material* materials, uint32_t nb_materials,
              uint32_t flags, const float* bvh_rgba, float* rgba_out) {
    if (!ctx) return RTR_E_INVALID;
    if (n == 0) return RTR_OK;
    if (!hits || !tris || !meshes || !materials || !rgba_out || !nb_triangles || !nb_meshes || !nb_materials)
        return rtr_set_error(ctx, RTR_E_INVALID, "shade: NULL or empty argument");
    const bool overlay = (flags & RTR_SHADE_BVH) != 0;
    if (overlay && !bvh_rgba) return rtr_set_error(ctx, RTR_E_INVALID, "shade: RTR_SHADE_BVH without overlay colours");
    // the records index triangles, models and materials: reject what the shader would read out of bounds
    for (uint64_t i = 0; i < n; ++i)
        if (hits[i].did_hit && hits[i].triangle_id >= nb_triangles)
            return rtr_set_error(ctx, RTR_E_INVALID, "shade: hit %llu names triangle %u of %u", (unsigned long long)i,
                                 hits[i].triangle_id, nb_triangles);
    for (uint32_t t = 0; t < nb_triangles; ++t)
        if (tris[t].model_id >= nb_meshes) return rtr_set_error(ctx, RTR_E_INVALID, "shade: triangle %u names model %u of %u", t, tris[t].model_id, nb_meshes);
    for (uint32_t m = 0; m < nb_meshes; ++m)
        if (meshes[m].material_id >= nb_materials) return rtr_set_error(ctx, RTR_E_INVALID, "shade: model %u names material %u of %u", m, meshes[m].material_id, nb_materials);
    const size_t hb = n * sizeof(rtr_hit), tb = (size_t)nb_triangles * sizeof(rtr_triangle), mb = (size_t)nb_meshes * sizeof(rtr_mesh);
    const size_t ab = (size_t)nb_materials * sizeof(rtr_material), cb = n * 16;
    Staging st;
    RTR_CHECK(staging_begin(ctx, hb + tb + mb + ab + 2 * cb + 2048, 0, &st));
    rtr_hit* d_hits = static_cast<rtr_hit*>(staging_take(&st, hb));
    rtr_triangle* d_tris = static_cast<rtr_triangle*>(staging_take(&st, tb));
    rtr_mesh* d_meshes = static_cast<rtr_mesh*>(staging_take(&st, mb));
    rtr_material* d_mat = static_cast<rtr_material*>(staging_take(&st, ab));
    float* d_rgba = static_cast<float*>(staging_take(&st, cb));
    float* d_bvh = static_cast<float*>(staging_take(&st, cb));
    auto body = [&]() -> int {
        RTR_CUDA(ctx, cudaMemcpyAsync(d_hits, hits, hb, cudaMemcpyHostToDevice, ctx->stream));
        RTR_CUDA(ctx, cudaMemcpyAsync(d_tris, tris, tb, cudaMemcpyHostToDevice, ctx->stream));
        RTR_CUDA(ctx, cudaMemcpyAsync(d_meshes, meshes, mb, cudaMemcpyHostToDevice, ctx->stream));
        RTR_CUDA(ctx, cudaMemcpyAsync(d_mat, materials, ab, cudaMemcpyHostToDevice, ctx->stream));
        if (overlay) RTR_CUDA(ctx, cudaMemcpyAsync(d_bvh, bvh_rgba, cb, cudaMemcpyHostToDevice, ctx->stream));
        RTR_CHECK(rtr_shade_launch(ctx, d_hits, n, d_tris, d_meshes, d_mat, flags, overlay ? d_bvh : nullptr, d_rgba));
        RTR_CUDA(ctx, cudaMemcpyAsync(rgba_out, d_rgba, cb, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    const int r = body();
    staging_end(&st);
    return r;
}

int rtr_bvh_depth_overlay_dev(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera* cam, uint32_t width, uint32_t height,
                              uint32_t denom_w, uint32_t denom_h, int display_depth, float* bvh_rgba_dev) {
    RTR_CHECK(trace_args_ok(ctx, b, cam));
    RTR_CHECK(order_ok(ctx, b, RTR_TRACE_REFERENCE_ORDER));  // walks the 48-byte nodes
    if (!bvh_rgba_dev) return rtr_set_error(ctx, RTR_E_INVALID, "depth_overlay: NULL output");
    return rtr_depth_overlay_launch(ctx, b, *cam, width, height, denom_w, denom_h, display_depth, bvh_rgba_dev);
}

int rtr_bvh_depth_overlay(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera* cam, uint32_t width, uint32_t height,
                          uint32_t denom_w, uint32_t denom_h, int display_depth, float* bvh_rgba_out) {
    RTR_CHECK(trace_args_ok(ctx, b, cam));
    RTR_CHECK(order_ok(ctx, b, RTR_TRACE_REFERENCE_ORDER));
    if (!bvh_rgba_out) return rtr_set_error(ctx, RTR_E_INVALID, "depth_overlay: NULL output");
    const size_t cb = (size_t)width * height * 16;
    Staging st;
    RTR_CHECK(staging_begin(ctx, cb + 512, 0, &st));
    float* d_out = static_cast<float*>(staging_take(&st, cb));
    auto body = [&]() -> int {
        RTR_CHECK(rtr_depth_overlay_launch(ctx, b, *cam, width, height, denom_w, denom_h, display_depth, d_out));
        RTR_CUDA(ctx, cudaMemcpyAsync(bvh_rgba_out, d_out, cb, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    const int r = body();
    staging_end(&st);
    return r;
}

int rtr_render(rtr_ctx* ctx, const rtr_bvh* b, const rtr_camera* cam, uint32_t width, uint32_t height, uint32_t denom_w,
               uint32_t denom_h, uint32_t row0, uint32_t row1, uint32_t bounces, int shadow, const float light_pos[3],
               uint32_t flags, float* rgba_out, rtr_hit* hits_out, uint64_t* rays_traced) {
    if (b) RTR_CHECK(rtr_bvh_finish(const_cast<rtr_bvh*>(b)));  // a build submitted with rtr_bvh_build_dev may still owe its verdict
    RTR_CHECK(trace_args_ok(ctx, b, cam));
    RTR_CHECK(order_ok(ctx, b, flags));
    if (shadow && !light_pos) return rtr_set_error(ctx, RTR_E_INVALID, "render: shadow rays need a light position");
    if (row1 == 0) row1 = height;
    if (row0 >= row1 || row1 > height) return rtr_set_error(ctx, RTR_E_INVALID, "render: bad rows [%u,%u)", row0, row1);
    const size_t px = (size_t)width * (row1 - row0);
    const size_t cb = px * 16, hb = px * sizeof(rtr_hit);
    Staging st;
    RTR_CHECK(staging_begin(ctx, cb + hb + 1024, 0, &st));
    float* d_rgba = static_cast<float*>(staging_take(&st, cb));
    rtr_hit* d_hits = static_cast<rtr_hit*>(staging_take(&st, hb));
    uint64_t* d_rays = static_cast<uint64_t*>(staging_take(&st, 8));
    auto attempt = [&](uint32_t f, uint32_t* overflows) -> int {
        RTR_CUDA(ctx, cudaMemsetAsync(d_rays, 0, 8, ctx->stream));
        RTR_CHECK(rtr_render_launch(ctx, b, *cam, width, height, denom_w, denom_h, row0, row1, bounces, shadow, light_pos,
                                    f, d_rgba, hits_out ? d_hits : nullptr, d_rays));
        if (rgba_out) RTR_CUDA(ctx, cudaMemcpyAsync(rgba_out, d_rgba, cb, cudaMemcpyDeviceToHost, ctx->stream));
        if (hits_out) RTR_CUDA(ctx, cudaMemcpyAsync(hits_out, d_hits, hb, cudaMemcpyDeviceToHost, ctx->stream));
        if (rays_traced) RTR_CUDA(ctx, cudaMemcpyAsync(rays_traced, d_rays, 8, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CHECK(fetch_overflows(ctx, b, overflows));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return RTR_OK;
    };
    const int r = trace_with_fallback(ctx, b, flags, "render", attempt);
    staging_end(&st);
    return r;
}

}  // extern "C"
