// comm.cu -- multi-GPU plumbing of the C ABI: one process per GPU, the BVH is built on one rank and
// broadcast over NVLink, image rows are dealt to ranks in blocks and all-gathered.
//
// The reference has no multi-GPU path at all; this is new surface.
// PLOC iterations are globally dependent, so the build itself is NOT sharded -- only its result is
// replicated -- while rays are independent and shard without any data-path collective.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): inside a torch process that is the
// library torch.distributed already loaded (same soname), in a plain C++ host it is the system one.
#include <dlfcn.h>

#include "bvh.cuh"
#include <vector>

namespace {

typedef struct { char internal[RTR_NCCL_UNIQUE_ID_BYTES]; } NcclUniqueId;
typedef void* NcclComm;
constexpr int kNcclUint8 = 1;

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*);
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
    int (*CommDestroy)(NcclComm);
    int (*CommAbort)(NcclComm);
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char* (*GetErrorString)(int);
    void* lib;
};

NcclApi g_nccl = {};

int load_nccl(rtr_ctx* ctx) {
    if (g_nccl.lib) return RTR_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return rtr_set_error(ctx, RTR_E_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
    NcclApi api = {};
    api.lib = lib;
#define RTR_SYM(field, name)                                                              \
    *reinterpret_cast<void**>(&api.field) = dlsym(lib, name);                             \
    if (!api.field) return rtr_set_error(ctx, RTR_E_COMM, "libnccl lacks symbol %s", name)
    RTR_SYM(GetUniqueId, "ncclGetUniqueId");
    RTR_SYM(CommInitRank, "ncclCommInitRank");
    RTR_SYM(CommDestroy, "ncclCommDestroy");
    RTR_SYM(CommAbort, "ncclCommAbort");
    RTR_SYM(Broadcast, "ncclBroadcast");
    RTR_SYM(Send, "ncclSend");
    RTR_SYM(Recv, "ncclRecv");
    RTR_SYM(GroupStart, "ncclGroupStart");
    RTR_SYM(GroupEnd, "ncclGroupEnd");
    RTR_SYM(GetErrorString, "ncclGetErrorString");
#undef RTR_SYM
    g_nccl = api;
    return RTR_OK;
}

#define RTR_NCCL(ctx, call)                                                                          \
    do {                                                                                             \
        int _e = (call);                                                                             \
        if (_e != 0)                                                                                 \
            return rtr_set_error((ctx), RTR_E_COMM, "%s:%d %s -> %s", __FILE__, __LINE__, #call,     \
                                 g_nccl.GetErrorString ? g_nccl.GetErrorString(_e) : "nccl error"); \
    } while (0)

// A rank that finds, before it joins a collective, that it cannot take part (the root has no BVH to send, ...) must not
// just return: its peers have enqueued their side and would wait forever.  The communicator is aborted instead, which
// makes every rank's pending and later calls on it fail (RTR_E_COMM) -- the ranks fail together.
int comm_fail(rtr_ctx* ctx, int code, const char* what) {
    if (ctx->nccl_comm && ctx->nranks > 1 && g_nccl.CommAbort) {
        g_nccl.CommAbort(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return rtr_set_error(ctx, code, "%s (the communicator was aborted so that the other ranks do not wait)", what);
}

// Inside ncclGroupStart/End an error must not return before the group is closed (the thread would stay in group
// mode): RTR_NCCL_G remembers the first failure and carries on, rtr_group_end closes the group and reports it.
struct GroupGuard {
    int first_error = 0;
    const char* what = nullptr;
};
#define RTR_NCCL_G(g, call)                                            \
    do {                                                               \
        if ((g).first_error == 0) {                                    \
            int _e = (call);                                           \
            if (_e != 0) { (g).first_error = _e; (g).what = #call; }   \
        }                                                              \
    } while (0)
int rtr_group_end(rtr_ctx* ctx, const GroupGuard& g) {
    const int e = g_nccl.GroupEnd();
    if (g.first_error != 0)
        return rtr_set_error(ctx, RTR_E_COMM, "%s -> %s", g.what, g_nccl.GetErrorString ? g_nccl.GetErrorString(g.first_error) : "nccl error");
    if (e != 0) return rtr_set_error(ctx, RTR_E_COMM, "ncclGroupEnd -> %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "nccl error");
    return RTR_OK;
}

}  // namespace

extern "C" {

int rtr_comm_unique_id(void* id_out) {
    if (!id_out) return RTR_E_INVALID;
    RTR_CHECK(load_nccl(nullptr));
    NcclUniqueId id;
    const int e = g_nccl.GetUniqueId(&id);
    if (e != 0) return rtr_set_error(nullptr, RTR_E_COMM, "ncclGetUniqueId -> %s", g_nccl.GetErrorString(e));
    memcpy(id_out, &id, sizeof(id));
    return RTR_OK;
}

int rtr_comm_init(rtr_ctx* ctx, const void* unique_id, int rank, int nranks) {
    if (!ctx) return RTR_E_INVALID;
    if (!unique_id || nranks < 1 || rank < 0 || rank >= nranks)
        return rtr_set_error(ctx, RTR_E_INVALID, "comm_init: bad rank %d / nranks %d", rank, nranks);
    if (ctx->nccl_comm) return rtr_set_error(ctx, RTR_E_STATE, "comm_init: communicator already initialised");
    RTR_CHECK(load_nccl(ctx));
    RTR_CUDA(ctx, cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    NcclComm comm = nullptr;
    RTR_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return RTR_OK;
}

int rtr_comm_destroy(rtr_ctx* ctx) {
    if (!ctx) return RTR_E_INVALID;
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy(ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->nranks = 1;
    return RTR_OK;
}

int rtr_bvh_broadcast(rtr_ctx* ctx, rtr_bvh** bvh, int root) {
    if (!ctx || !bvh) return RTR_E_INVALID;
    if (!ctx->nccl_comm) return rtr_set_error(ctx, RTR_E_STATE, "bvh_broadcast: rtr_comm_init has not been called");
    if (root < 0 || root >= ctx->nranks) return rtr_set_error(ctx, RTR_E_INVALID, "bvh_broadcast: bad root %d", root);
    const bool is_root = ctx->rank == root;
    if (is_root && (!*bvh || !(*bvh)->built)) return comm_fail(ctx, RTR_E_STATE, "bvh_broadcast: root has no built BVH");
    if (is_root && (*bvh)->trav_only) return comm_fail(ctx, RTR_E_STATE, "bvh_broadcast: the root holds traversal records only");
    NcclComm comm = ctx->nccl_comm;

    // 1. header: triangle and mesh counts
    RTR_CHECK(rtr_ws_reserve(ctx, 256));
    uint32_t* h_hdr = static_cast<uint32_t*>(ctx->pinned) + 64;
    uint32_t* d_hdr = static_cast<uint32_t*>(ctx->ws);
    if (is_root) {
        // [2]: the root's triangle-cache indexing (1 = by leaf rank, a BVH built here; 0 = by triangle id, adopted)
        h_hdr[0] = (*bvh)->n; h_hdr[1] = (*bvh)->nb_meshes; h_hdr[2] = (*bvh)->wtri_by_rank ? 1u : 0u; h_hdr[3] = 0;
        RTR_CUDA(ctx, cudaMemcpyAsync(d_hdr, h_hdr, 16, cudaMemcpyHostToDevice, ctx->stream));
    }
    RTR_NCCL(ctx, g_nccl.Broadcast(d_hdr, d_hdr, 16, kNcclUint8, root, comm, ctx->stream));
    RTR_CUDA(ctx, cudaMemcpyAsync(h_hdr, d_hdr, 16, cudaMemcpyDeviceToHost, ctx->stream));
    RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t n = h_hdr[0], nb_meshes = h_hdr[1];
    const bool by_rank = h_hdr[2] != 0;
    if (n == 0 || nb_meshes == 0) return rtr_set_error(ctx, RTR_E_COMM, "bvh_broadcast: empty header from root");
    const size_t nc = 2 * (size_t)n - 1;

    // 2. receivers own their copies
    rtr_bvh* b = *bvh;
    if (!is_root) {
        if (!b) {
            b = new (std::nothrow) rtr_bvh();
            if (!b) return rtr_set_error(ctx, RTR_E_NOMEM, "bvh_broadcast: host allocation failed");
            b->ctx = ctx;
            *bvh = b;
        }
        b->built = false;
        if (b->recv_cap < n) {
            RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (b->flat_recv) cudaFree(b->flat_recv);
            if (b->tris_own) cudaFree(b->tris_own);
            b->flat_recv = nullptr; b->tris_own = nullptr; b->tris_own_cap = 0; b->recv_cap = 0;
            RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->flat_recv), nc * sizeof(rtr_node)));
            RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->tris_own), (size_t)n * sizeof(rtr_triangle)));
            b->tris_own_cap = n; b->recv_cap = n;
        }
        if (b->meshes_own_cap < nb_meshes) {
            RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (b->meshes_own) cudaFree(b->meshes_own);
            b->meshes_own = nullptr; b->meshes_own_cap = 0;
            RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->meshes_own), (size_t)nb_meshes * sizeof(rtr_mesh)));
            b->meshes_own_cap = nb_meshes;
        }
        if (!b->tparams) RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->tparams), sizeof(TraceParams)));
        if (b->wtri_own_cap < n) {
            RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (b->wtri_own) cudaFree(b->wtri_own);
            b->wtri_own = nullptr; b->wtri_own_cap = 0;
            RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->wtri_own), (size_t)n * 3 * sizeof(float4)));
            b->wtri_own_cap = n;
        }
    }
    void* nodes = is_root ? const_cast<rtr_node*>(b->flat_view) : b->flat_recv;
    void* tris = is_root ? const_cast<rtr_triangle*>(b->tris) : b->tris_own;
    void* meshes = is_root ? const_cast<rtr_mesh*>(b->meshes) : b->meshes_own;
    void* wtri = is_root ? const_cast<float4*>(b->wtri_view) : b->wtri_own;

    // 3. one grouped broadcast: flat nodes (48 B*(2n-1)) + triangles (64 B*n) + meshes + world-space
    //    triangle cache (48 B*n) + trace constants
    RTR_NCCL(ctx, g_nccl.GroupStart());
    GroupGuard grp;
    RTR_NCCL_G(grp, g_nccl.Broadcast(nodes, nodes, nc * sizeof(rtr_node), kNcclUint8, root, comm, ctx->stream));
    RTR_NCCL_G(grp, g_nccl.Broadcast(tris, tris, (size_t)n * sizeof(rtr_triangle), kNcclUint8, root, comm, ctx->stream));
    RTR_NCCL_G(grp, g_nccl.Broadcast(meshes, meshes, (size_t)nb_meshes * sizeof(rtr_mesh), kNcclUint8, root, comm, ctx->stream));
    RTR_NCCL_G(grp, g_nccl.Broadcast(wtri, wtri, (size_t)n * 3 * sizeof(float4), kNcclUint8, root, comm, ctx->stream));
    RTR_NCCL_G(grp, g_nccl.Broadcast(b->tparams, b->tparams, sizeof(TraceParams), kNcclUint8, root, comm, ctx->stream));
    RTR_CHECK(rtr_group_end(ctx, grp));

    if (!is_root) {
        b->n = n; b->array_len = n; b->nb_meshes = nb_meshes;
        b->tris = b->tris_own; b->meshes = b->meshes_own; b->flat_view = b->flat_recv;
        b->wtri_view = b->wtri_own; b->wtri_by_rank = by_rank;
        b->adopted = true; b->trav_only = false;
        RTR_CHECK(rtr_bvh_pack_pairs_own(b));  // derived locally: cheaper than 64 B*(2n-1) more on the wire
        b->built = true;
    }
    return RTR_OK;
}

int rtr_bvh_broadcast_traversal(rtr_ctx* ctx, rtr_bvh** bvh, int root, uint32_t expected_triangles) {
    if (!ctx || !bvh) return RTR_E_INVALID;
    if (!ctx->nccl_comm) return rtr_set_error(ctx, RTR_E_STATE, "bvh_broadcast: rtr_comm_init has not been called");
    if (root < 0 || root >= ctx->nranks) return rtr_set_error(ctx, RTR_E_INVALID, "bvh_broadcast: bad root %d", root);
    const bool is_root = ctx->rank == root;
    if (is_root && (!*bvh || !(*bvh)->built)) return comm_fail(ctx, RTR_E_STATE, "bvh_broadcast: root has no built BVH");
    NcclComm comm = ctx->nccl_comm;

    // 1. header: triangle count -- unless every rank already knows it (a frame loop over one scene): then no
    //    rank has to wait on the host for the root, and the receive can be enqueued ahead of the root's rebuild
    uint32_t n = expected_triangles;
    if (n == 0) {
        RTR_CHECK(rtr_ws_reserve(ctx, 256));
        uint32_t* h_hdr = static_cast<uint32_t*>(ctx->pinned) + 64;
        uint32_t* d_hdr = static_cast<uint32_t*>(ctx->ws);
        if (is_root) {
            h_hdr[0] = (*bvh)->n; h_hdr[1] = h_hdr[2] = h_hdr[3] = 0;
            RTR_CUDA(ctx, cudaMemcpyAsync(d_hdr, h_hdr, 16, cudaMemcpyHostToDevice, ctx->stream));
        }
        RTR_NCCL(ctx, g_nccl.Broadcast(d_hdr, d_hdr, 16, kNcclUint8, root, comm, ctx->stream));
        RTR_CUDA(ctx, cudaMemcpyAsync(h_hdr, d_hdr, 16, cudaMemcpyDeviceToHost, ctx->stream));
        RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        n = h_hdr[0];
        if (n == 0) return rtr_set_error(ctx, RTR_E_COMM, "bvh_broadcast: empty header from root");
    } else if (is_root && (*bvh)->n != n) {
        return comm_fail(ctx, RTR_E_INVALID, "bvh_broadcast: the root's BVH does not have the triangle count the ranks expect");
    }
    const size_t nc = 2 * (size_t)n - 1;

    // 2. receivers own their copy of the traversal records and of node 0 (the root box every ray starts with)
    rtr_bvh* b = *bvh;
    if (!is_root) {
        if (!b) {
            b = new (std::nothrow) rtr_bvh();
            if (!b) return rtr_set_error(ctx, RTR_E_NOMEM, "bvh_broadcast: host allocation failed");
            b->ctx = ctx;
            *bvh = b;
        }
        b->built = false;
        if (!b->flat_recv) {
            RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->flat_recv), sizeof(rtr_node)));
            b->recv_cap = 0;  // one node only: a later full broadcast reallocates
        }
        if (b->pairs_own_cap < n) {
            RTR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (b->pairs_own) cudaFree(b->pairs_own);
            b->pairs_own = nullptr; b->pairs_own_cap = 0;
            RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->pairs_own), nc * 4 * sizeof(uint4)));
            b->pairs_own_cap = n;
        }
        if (!b->tparams) RTR_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->tparams), sizeof(TraceParams)));
    }
    void* pairs = is_root ? const_cast<uint4*>(b->pairs_view) : b->pairs_own;
    void* node0 = is_root ? const_cast<rtr_node*>(b->flat_view) : b->flat_recv;

    // 3. one grouped broadcast: 64 B*(2n-1) + 48 B + trace constants
    RTR_NCCL(ctx, g_nccl.GroupStart());
    GroupGuard grp;
    RTR_NCCL_G(grp, g_nccl.Broadcast(pairs, pairs, nc * 4 * sizeof(uint4), kNcclUint8, root, comm, ctx->stream));
    RTR_NCCL_G(grp, g_nccl.Broadcast(node0, node0, sizeof(rtr_node), kNcclUint8, root, comm, ctx->stream));
    RTR_NCCL_G(grp, g_nccl.Broadcast(b->tparams, b->tparams, sizeof(TraceParams), kNcclUint8, root, comm, ctx->stream));
    RTR_CHECK(rtr_group_end(ctx, grp));

    if (!is_root) {
        b->n = n; b->array_len = n; b->nb_meshes = 0;
        b->tris = nullptr; b->meshes = nullptr; b->flat_view = b->flat_recv; b->wtri_view = nullptr;
        b->pairs_view = b->pairs_own;
        b->adopted = true; b->trav_only = true;
        b->built = true;
    }
    return RTR_OK;
}

int rtr_allgather_rows(rtr_ctx* ctx, void* image_dev, uint32_t width, uint32_t height, uint32_t bytes_per_pixel,
                       uint32_t rows_per_block) {
    if (!ctx) return RTR_E_INVALID;
    if (!image_dev || width == 0 || height == 0 || bytes_per_pixel == 0 || rows_per_block == 0)
        return rtr_set_error(ctx, RTR_E_INVALID, "allgather_rows: bad argument");
    if (ctx->nranks == 1) return RTR_OK;
    if (!ctx->nccl_comm) return rtr_set_error(ctx, RTR_E_STATE, "allgather_rows: rtr_comm_init has not been called");
    const size_t row_bytes = (size_t)width * bytes_per_pixel;
    const uint32_t blocks = (height + rows_per_block - 1) / rows_per_block;
    RTR_NCCL(ctx, g_nccl.GroupStart());
    GroupGuard grp;
    for (uint32_t blk = 0; blk < blocks; ++blk) {
        const uint32_t r0 = blk * rows_per_block;
        const uint32_t r1 = (r0 + rows_per_block < height) ? r0 + rows_per_block : height;
        char* p = static_cast<char*>(image_dev) + (size_t)r0 * row_bytes;
        RTR_NCCL_G(grp, g_nccl.Broadcast(p, p, (size_t)(r1 - r0) * row_bytes, kNcclUint8, (int)(blk % (uint32_t)ctx->nranks),
                                       ctx->nccl_comm, ctx->stream));
    }
    RTR_CHECK(rtr_group_end(ctx, grp));
    return RTR_OK;
}

int rtr_allgather_stripes(rtr_ctx* ctx, void* image_dev, uint32_t width, uint32_t height, uint32_t bytes_per_pixel,
                          uint32_t rows_per_block, const uint32_t* stripes_of_rank) {
    if (!ctx) return RTR_E_INVALID;
    if (!image_dev || width == 0 || height == 0 || bytes_per_pixel == 0 || rows_per_block == 0 || !stripes_of_rank)
        return rtr_set_error(ctx, RTR_E_INVALID, "allgather_stripes: bad argument");
    if (ctx->nranks == 1) return RTR_OK;
    if (!ctx->nccl_comm) return rtr_set_error(ctx, RTR_E_STATE, "allgather_stripes: rtr_comm_init has not been called");
    const std::vector<int> owner = rtr_stripe_owners(stripes_of_rank, ctx->nranks);  // stripe -> rank
    if (owner.empty()) return rtr_set_error(ctx, RTR_E_INVALID, "allgather_stripes: no stripes");
    const size_t row_bytes = (size_t)width * bytes_per_pixel;
    const uint32_t blocks = (height + rows_per_block - 1) / rows_per_block;
    RTR_NCCL(ctx, g_nccl.GroupStart());
    GroupGuard grp;
    for (uint32_t blk = 0; blk < blocks; ++blk) {
        const uint32_t r0 = blk * rows_per_block;
        const uint32_t r1 = (r0 + rows_per_block < height) ? r0 + rows_per_block : height;
        char* p = static_cast<char*>(image_dev) + (size_t)r0 * row_bytes;
        RTR_NCCL_G(grp, g_nccl.Broadcast(p, p, (size_t)(r1 - r0) * row_bytes, kNcclUint8, owner[blk % owner.size()],
                                       ctx->nccl_comm, ctx->stream));
    }
    RTR_CHECK(rtr_group_end(ctx, grp));
    return RTR_OK;
}

int rtr_gather_stripes(rtr_ctx* ctx, void* image_dev, uint32_t width, uint32_t height, uint32_t bytes_per_pixel,
                       uint32_t rows_per_block, const uint32_t* stripes_of_rank, int root) {
    if (!ctx) return RTR_E_INVALID;
    if (!image_dev || width == 0 || height == 0 || bytes_per_pixel == 0 || rows_per_block == 0 || !stripes_of_rank)
        return rtr_set_error(ctx, RTR_E_INVALID, "gather_stripes: bad argument");
    if (ctx->nranks == 1) return RTR_OK;
    if (!ctx->nccl_comm) return rtr_set_error(ctx, RTR_E_STATE, "gather_stripes: rtr_comm_init has not been called");
    if (root < 0 || root >= ctx->nranks) return rtr_set_error(ctx, RTR_E_INVALID, "gather_stripes: bad root %d", root);
    const std::vector<int> owner = rtr_stripe_owners(stripes_of_rank, ctx->nranks);  // stripe -> rank
    if (owner.empty()) return rtr_set_error(ctx, RTR_E_INVALID, "gather_stripes: no stripes");
    const size_t row_bytes = (size_t)width * bytes_per_pixel;
    const uint32_t blocks = (height + rows_per_block - 1) / rows_per_block;
    // point to point: every block travels once, from its owner to the root (an all-gather moves nranks - 1 copies)
    RTR_NCCL(ctx, g_nccl.GroupStart());
    GroupGuard grp;
    for (uint32_t blk = 0; blk < blocks; ++blk) {
        const int own = owner[blk % owner.size()];
        if (own == root || (ctx->rank != root && ctx->rank != own)) continue;
        const uint32_t r0 = blk * rows_per_block;
        const uint32_t r1 = (r0 + rows_per_block < height) ? r0 + rows_per_block : height;
        char* p = static_cast<char*>(image_dev) + (size_t)r0 * row_bytes;
        const size_t bytes = (size_t)(r1 - r0) * row_bytes;
        if (ctx->rank == root) RTR_NCCL_G(grp, g_nccl.Recv(p, bytes, kNcclUint8, own, ctx->nccl_comm, ctx->stream));
        else RTR_NCCL_G(grp, g_nccl.Send(p, bytes, kNcclUint8, root, ctx->nccl_comm, ctx->stream));
    }
    RTR_CHECK(rtr_group_end(ctx, grp));
    return RTR_OK;
}

// Sharded host input: an array of n_elems elements is cut into nranks contiguous slices, rank r holding elements
// [n_elems * r / nranks, n_elems * (r + 1) / nranks) in slice_dev (each rank uploaded its slice over its own PCIe
// link); the slices are assembled in full_dev on `root` over NVLink.  Every slice travels once.
int rtr_gather_slices(rtr_ctx* ctx, const void* slice_dev, void* full_dev, uint64_t n_elems, uint32_t elem_bytes, int root) {
    if (!ctx) return RTR_E_INVALID;
    if (elem_bytes == 0 || (n_elems && !slice_dev)) return rtr_set_error(ctx, RTR_E_INVALID, "gather_slices: bad argument");
    if (root < 0 || root >= ctx->nranks) return rtr_set_error(ctx, RTR_E_INVALID, "gather_slices: bad root %d", root);
    const bool is_root = ctx->rank == root;
    if (is_root && !full_dev && n_elems) return rtr_set_error(ctx, RTR_E_INVALID, "gather_slices: the root needs full_dev");
    auto lo_of = [&](int r) { return n_elems * (uint64_t)r / (uint64_t)ctx->nranks; };
    const uint64_t my_lo = lo_of(ctx->rank), my_hi = lo_of(ctx->rank + 1);
    if (is_root && my_hi > my_lo) {
        char* dst = static_cast<char*>(full_dev) + my_lo * elem_bytes;
        if (dst != slice_dev)
            RTR_CUDA(ctx, cudaMemcpyAsync(dst, slice_dev, (my_hi - my_lo) * elem_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (ctx->nranks == 1) return RTR_OK;
    if (!ctx->nccl_comm) return rtr_set_error(ctx, RTR_E_STATE, "gather_slices: rtr_comm_init has not been called");
    RTR_NCCL(ctx, g_nccl.GroupStart());
    GroupGuard grp;
    if (is_root) {
        for (int r = 0; r < ctx->nranks; ++r) {
            if (r == root) continue;
            const uint64_t lo = lo_of(r), hi = lo_of(r + 1);
            if (hi > lo)
                RTR_NCCL_G(grp, g_nccl.Recv(static_cast<char*>(full_dev) + lo * elem_bytes, (hi - lo) * elem_bytes, kNcclUint8, r,
                                            ctx->nccl_comm, ctx->stream));
        }
    } else if (my_hi > my_lo) {
        RTR_NCCL_G(grp, g_nccl.Send(slice_dev, (my_hi - my_lo) * elem_bytes, kNcclUint8, root, ctx->nccl_comm, ctx->stream));
    }
    return rtr_group_end(ctx, grp);
}

// Sharded host output: this rank's row blocks of a striped frame go from image_dev straight to the same rows of
// image_host (a full-size host image, e.g. one pinned buffer shared by all ranks), over this rank's own PCIe link; one
// asynchronous copy per block on the context's stream.  No collective.
int rtr_download_stripes_async(rtr_ctx* ctx, void* image_host, const void* image_dev, uint32_t width, uint32_t height,
                               uint32_t bytes_per_pixel, uint32_t rows_per_block, const uint32_t* stripes_of_rank) {
    if (!ctx) return RTR_E_INVALID;
    if (!image_host || !image_dev || width == 0 || height == 0 || bytes_per_pixel == 0 || rows_per_block == 0 || !stripes_of_rank)
        return rtr_set_error(ctx, RTR_E_INVALID, "download_stripes: bad argument");
    const std::vector<int> owner = rtr_stripe_owners(stripes_of_rank, ctx->nranks);  // stripe -> rank
    if (owner.empty()) return rtr_set_error(ctx, RTR_E_INVALID, "download_stripes: no stripes");
    const size_t row_bytes = (size_t)width * bytes_per_pixel;
    const uint32_t blocks = (height + rows_per_block - 1) / rows_per_block;
    for (uint32_t blk = 0; blk < blocks; ++blk) {
        if (owner[blk % owner.size()] != ctx->rank) continue;
        const uint32_t r0 = blk * rows_per_block;
        const uint32_t r1 = (r0 + rows_per_block < height) ? r0 + rows_per_block : height;
        const size_t off = (size_t)r0 * row_bytes;
        RTR_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(image_host) + off, static_cast<const char*>(image_dev) + off,
                                      (size_t)(r1 - r0) * row_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return RTR_OK;
}

}  // extern "C"
