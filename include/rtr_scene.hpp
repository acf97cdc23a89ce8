// rtr_scene.hpp -- C++ host shim: the reference's srcCommon scene/BVH interface on top of the C ABI.
//
// A maintainer of MrBigoudi/RealTimeRaytracing who wants the B200 path replaces
//     #include "bvh.hpp"            (srcCommon/scene/geometry/bvh.hpp)
// by
//     #include "rtr_scene.hpp"
// and links librtr_b200.so: glr::Scene::bindSSBO (srcOpenGL/scene/scene.cpp:148) keeps constructing
//     _BVH = cr::BVH_Ptr(new cr::BVH(nbTriangles, trianglesGPU, modelsGPU));
// and glr::Scene::recursiveTopDownTraversalBVH (scene.cpp:189-201) keeps reading
//     bvh->_InternalStruct._Clusters[id].value(), _IsLeaf[id], _LeftChild[id].value(), _RightChild[id].value()
// unchanged, because this header declares the same types with the same member names:
//
//   cr::TriangleGPU   <- srcCommon/scene/geometry/triangle.hpp:9-14   (64 B, _ModelId at 48)
//   cr::MeshModelGPU  <- srcCommon/scene/geometry/mesh.hpp:12-15      (68 B)
//   cr::AABB_GPU, cr::BVH_NodeGPU, cr::BVH_Params, cr::BVH  <- srcCommon/scene/geometry/bvh.hpp:22-91
//
// plus what the reference does on the host after the build and what it does per frame:
//
//   cr::BVH::getFlatNodes()          = glr::Scene::getBVH_NodesToGPUData (scene.cpp:203-208), computed on the GPU
//   cr::BVH::tracePrimary(...)       = glDispatchCompute on raytracer.glsl (srcOpenGL/application.cpp:245)
//
// With glm on the include path define RTR_SCENE_USE_GLM before including this file and the vector
// members are glm::vec4 / glm::mat4 / glm::vec3 exactly as in the reference; without it, layout-identical
// PODs are used.  Errors follow the reference's convention (srcCommon/core/errorHandler.cpp:13-29: print
// file:line + message to stderr and exit(EXIT_FAILURE) for fatal errors) -- the C ABI below never exits.
#ifndef RTR_SCENE_HPP
#define RTR_SCENE_HPP

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <optional>
#include <vector>

#include "rtr.h"

#ifdef RTR_SCENE_USE_GLM
#include <glm/glm.hpp>
#endif

namespace cr {

#ifdef RTR_SCENE_USE_GLM
using vec3 = glm::vec3;
using vec4 = glm::vec4;
using mat4 = glm::mat4;
#else
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct mat4 {  // column-major like glm::mat4
    float m[16];
    static mat4 identity() { mat4 r; std::memset(&r, 0, sizeof(r)); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
};
#endif

// triangle.hpp:9-14
struct TriangleGPU {
    vec4 _P0;
    vec4 _P1;
    vec4 _P2;
    alignas(16) uint32_t _ModelId;
};
// mesh.hpp:12-15
struct MeshModelGPU {
#ifdef RTR_SCENE_USE_GLM
    mat4 _Model = mat4(1.f);
#else
    mat4 _Model = mat4::identity();
#endif
    uint32_t _MaterialId = 0;
};
// bvh.hpp:22-26
struct AABB_GPU {
    vec3 _Min = {INFINITY, INFINITY, INFINITY};
    alignas(16) vec3 _Max = {-INFINITY, -INFINITY, -INFINITY};
};
// bvh.hpp:36-42 ; leaf <=> _LeftChild == 0 && _RightChild == 0
struct BVH_NodeGPU {
    AABB_GPU _BoundingBox;
    uint32_t _TriangleId;
    uint32_t _LeftChild;
    uint32_t _RightChild;
};
static_assert(sizeof(TriangleGPU) == sizeof(rtr_triangle) && offsetof(TriangleGPU, _ModelId) == 48, "TriangleGPU layout");
static_assert(sizeof(MeshModelGPU) == sizeof(rtr_mesh), "MeshModelGPU layout");
static_assert(sizeof(AABB_GPU) == 32 && offsetof(AABB_GPU, _Max) == 16, "AABB_GPU layout");
static_assert(sizeof(BVH_NodeGPU) == sizeof(rtr_node) && offsetof(BVH_NodeGPU, _TriangleId) == 32 &&
              offsetof(BVH_NodeGPU, _LeftChild) == 36 && offsetof(BVH_NodeGPU, _RightChild) == 40, "BVH_NodeGPU layout");

// camera.hpp:21-30
struct CameraGPU {
    mat4 _View, _Proj, _InvView, _InvProj;
    vec4 _Eye;
    float _PlaneWidth, _PlaneHeight, _PlaneNear;
};
static_assert(sizeof(CameraGPU) == sizeof(rtr_camera), "CameraGPU layout");

// raytracer.glsl:36-40
struct Hit {
    vec4 _Coords;  // (b0, b1, b2, t)
    uint32_t _DidHit;
    uint32_t _TriangleId;
};
static_assert(sizeof(Hit) == sizeof(rtr_hit), "Hit layout");

// errorHandler.cpp:13-29 behaviour for a failed C-ABI call
inline void rtrFatal(rtr_ctx* ctx, int code, const char* what, const char* file, int line) {
    std::fprintf(stderr, "Error triggered in %s at line %d\n\t%s failed with code %d: %s\n", file, line, what, code,
                 rtr_last_error(ctx));
    std::exit(EXIT_FAILURE);
}
#define RTR_SCENE_CHECK(ctx, call)                                         \
    do {                                                                   \
        const int _rc = (call);                                            \
        if (_rc != RTR_OK) ::cr::rtrFatal((ctx), _rc, #call, __FILE__, __LINE__); \
    } while (0)

// one device context per host thread, created on first use (the reference has one GL context per app)
inline rtr_ctx* defaultContext(int device = 0) {
    static thread_local rtr_ctx* ctx = nullptr;
    if (!ctx) RTR_SCENE_CHECK(nullptr, rtr_ctx_create(device, &ctx));
    return ctx;
}

// bvh.hpp:44-63 ; sized by the actual triangle count instead of Triangle::MAX_NB_TRIANGLES
struct BVH_Params {
    size_t _NbTriangles = 0;
    std::vector<TriangleGPU> _UnsortedTriangles;
    std::vector<MeshModelGPU> _MeshesInTheScene;
    std::vector<std::optional<BVH_NodeGPU>> _Clusters;
    std::vector<std::optional<bool>> _IsLeaf;  // set for leaves only (bvh.cpp:41)
    std::vector<std::optional<uint32_t>> _Parent;
    std::vector<std::optional<uint32_t>> _LeftChild;
    std::vector<std::optional<uint32_t>> _RightChild;
    std::vector<uint32_t> _TriangleIndices;
    std::vector<uint32_t> _MortonCodes;  // PlocParams::_MortonCodes (bvh.hpp:70), sorted
};

class BVH;
using BVH_Ptr = std::shared_ptr<BVH>;

class BVH {
public:
    BVH_Params _InternalStruct = {};

    // bvh.hpp:89-91.  Builds synchronously on the GPU (Morton -> Onesweep -> PLOC -> flatten).
    // `fillInternalStruct = false` skips the download of the by-cluster-id view when only the flat
    // array / traversal are needed (e.g. a per-frame rebuild).
    BVH(uint32_t nbTriangles, const std::vector<TriangleGPU>& unsortedTriangles,
        const std::vector<MeshModelGPU>& meshesInTheScene, rtr_ctx* ctx = nullptr, bool fillInternalStruct = true,
        uint32_t searchRadius = RTR_DEFAULT_SEARCH_RADIUS)
        : _ctx(ctx ? ctx : defaultContext()) {
        RTR_SCENE_CHECK(_ctx, rtr_bvh_build(_ctx, reinterpret_cast<const rtr_triangle*>(unsortedTriangles.data()), nbTriangles,
                                           static_cast<uint32_t>(unsortedTriangles.size()),
                                           reinterpret_cast<const rtr_mesh*>(meshesInTheScene.data()),
                                           static_cast<uint32_t>(meshesInTheScene.size()), searchRadius, &_bvh));
        _InternalStruct._NbTriangles = nbTriangles;
        if (!fillInternalStruct) return;
        _InternalStruct._UnsortedTriangles = unsortedTriangles;  // the reference ctor copies both (bvh.cpp:16-17)
        _InternalStruct._MeshesInTheScene = meshesInTheScene;
        const size_t nc = 2 * static_cast<size_t>(nbTriangles) - 1;
        std::vector<rtr_node> clusters(nc);
        std::vector<uint32_t> parent(nc), left(nc), right(nc);
        std::vector<uint8_t> isLeaf(nc);
        RTR_SCENE_CHECK(_ctx, rtr_bvh_clusters(_bvh, clusters.data(), parent.data(), left.data(), right.data(), isLeaf.data()));
        BVH_Params& p = _InternalStruct;
        p._Clusters.assign(nc, std::nullopt); p._IsLeaf.assign(nc, std::nullopt); p._Parent.assign(nc, std::nullopt);
        p._LeftChild.assign(nc, std::nullopt); p._RightChild.assign(nc, std::nullopt);
        for (size_t c = 0; c < nc; ++c) {
            BVH_NodeGPU nd;
            std::memcpy(static_cast<void*>(&nd), &clusters[c], sizeof(nd));
            p._Clusters[c] = nd;
            if (isLeaf[c]) p._IsLeaf[c] = true;
            if (parent[c] != RTR_NONE) p._Parent[c] = parent[c];
            if (left[c] != RTR_NONE) p._LeftChild[c] = left[c];
            if (right[c] != RTR_NONE) p._RightChild[c] = right[c];
        }
        p._TriangleIndices.resize(nbTriangles);
        p._MortonCodes.resize(nbTriangles);
        RTR_SCENE_CHECK(_ctx, rtr_bvh_triangle_indices(_bvh, p._TriangleIndices.data()));
        RTR_SCENE_CHECK(_ctx, rtr_bvh_morton_codes(_bvh, p._MortonCodes.data()));
    }
    ~BVH() { rtr_bvh_destroy(_bvh); }
    BVH(const BVH&) = delete;
    BVH& operator=(const BVH&) = delete;

    // glr::Scene::getBVH_NodesToGPUData (scene.cpp:203-208): DFS pre-order array, root first
    std::vector<BVH_NodeGPU> getFlatNodes() const {
        std::vector<BVH_NodeGPU> out(rtr_bvh_nb_nodes(_bvh));
        RTR_SCENE_CHECK(_ctx, rtr_bvh_flat_nodes(_bvh, reinterpret_cast<rtr_node*>(out.data())));
        return out;
    }
    // the per-frame dispatch of raytracer.glsl (application.cpp:225-245): one closest hit per pixel;
    // pixels beyond floor(W/16)*16 x floor(H/16)*16 are not traced
    std::vector<Hit> tracePrimary(const CameraGPU& camera, uint32_t width, uint32_t height) const {
        std::vector<Hit> hits(static_cast<size_t>(width) * height);
        RTR_SCENE_CHECK(_ctx, rtr_trace_primary(_ctx, _bvh, reinterpret_cast<const rtr_camera*>(&camera), width, height, 0, 0,
                                               RTR_TRACE_DEFAULT, reinterpret_cast<rtr_hit*>(hits.data())));
        return hits;
    }
    rtr_bvh* handle() const { return _bvh; }
    rtr_ctx* context() const { return _ctx; }

private:
    rtr_ctx* _ctx = nullptr;
    rtr_bvh* _bvh = nullptr;
};

}  // namespace cr

#endif  // RTR_SCENE_HPP
