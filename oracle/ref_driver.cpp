// ref_driver.cpp -- thin C-ABI driver around the REFERENCE's own cr::BVH
// (srcCommon/scene/geometry/bvh.cpp + triangle.cpp compiled from where they lie
// under /root/reference; see oracle/Makefile).  TEST INFRASTRUCTURE ONLY: it is
// linked into oracle/_ref/libref_bvh_<capacity>.so, which is used to pin the C
// restatement (rtr_oracle.c), to generate tests/golden/, and as the CPU baseline
// ("kind": "reference") of bench.py.  No reference source is copied into the repo.
//
// `private` is opened so the stage-level getMortonCodes() (bvh.hpp:100) can be
// compared; everything else is read from the public BVH::_InternalStruct (bvh.hpp:86).
#define private public
#include "bvh.hpp"
#undef private

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include <omp.h>

namespace {
struct RefNode {  // same 48-byte layout as cr::BVH_NodeGPU, padding zeroed
    float bmin[3]; uint32_t pad0;
    float bmax[3]; uint32_t pad1;
    uint32_t tri, left, right, pad2;
};
static_assert(sizeof(cr::TriangleGPU) == 64, "TriangleGPU layout");
static_assert(sizeof(cr::MeshModelGPU) == 68, "MeshModelGPU layout");
static_assert(sizeof(cr::BVH_NodeGPU) == 48, "BVH_NodeGPU layout");
static_assert(sizeof(RefNode) == 48, "RefNode layout");

struct RefBvh {
    cr::BVH* bvh;
    uint32_t n;
    double build_ms;
};

void copy_node(const cr::BVH_NodeGPU& s, RefNode& d) {
    std::memset(&d, 0, sizeof(d));
    d.bmin[0] = s._BoundingBox._Min.x; d.bmin[1] = s._BoundingBox._Min.y; d.bmin[2] = s._BoundingBox._Min.z;
    d.bmax[0] = s._BoundingBox._Max.x; d.bmax[1] = s._BoundingBox._Max.y; d.bmax[2] = s._BoundingBox._Max.z;
    d.tri = s._TriangleId; d.left = s._LeftChild; d.right = s._RightChild;
}
}  // namespace

extern "C" {

uint64_t ref_max_triangles() { return cr::Triangle::MAX_NB_TRIANGLES; }
uint64_t ref_max_meshes() { return cr::Mesh::MAX_NB_MESHES; }

// Runs the reference constructor (bvh.cpp:11-24) exactly as glr::Scene::bindSSBO
// does (scene.cpp:148): vectors of array_len triangles / nb_meshes models.
void* ref_bvh_build(const void* tris, uint32_t n, uint32_t array_len,
                    const void* meshes, uint32_t nb_meshes) {
    if (n == 0 || n > cr::Triangle::MAX_NB_TRIANGLES || array_len < n) return nullptr;
    std::vector<cr::TriangleGPU> tv(array_len);
    std::memcpy(tv.data(), tris, sizeof(cr::TriangleGPU) * (size_t)array_len);
    std::vector<cr::MeshModelGPU> mv(nb_meshes);
    std::memcpy(mv.data(), meshes, sizeof(cr::MeshModelGPU) * (size_t)nb_meshes);
    auto t0 = std::chrono::steady_clock::now();
    cr::BVH* b = new cr::BVH(n, tv, mv);
    auto t1 = std::chrono::steady_clock::now();
    RefBvh* r = new RefBvh{b, n, std::chrono::duration<double, std::milli>(t1 - t0).count()};
    return r;
}

double ref_bvh_build_ms(const void* h) { return static_cast<const RefBvh*>(h)->build_ms; }

void ref_bvh_destroy(void* h) {
    RefBvh* r = static_cast<RefBvh*>(h);
    if (!r) return;
    delete r->bvh;
    delete r;
}

// unsorted Morton codes, bvh.cpp:330-348
void ref_bvh_morton_codes(const void* h, uint32_t* out) {
    const RefBvh* r = static_cast<const RefBvh*>(h);
    std::vector<uint32_t> codes = r->bvh->getMortonCodes();
    std::memcpy(out, codes.data(), sizeof(uint32_t) * (size_t)r->n);
}

void ref_bvh_triangle_indices(const void* h, uint32_t* out) {
    const RefBvh* r = static_cast<const RefBvh*>(h);
    std::memcpy(out, r->bvh->_InternalStruct._TriangleIndices.data(), sizeof(uint32_t) * (size_t)r->n);
}

// cluster-id indexed arrays; absent optionals -> 0xFFFFFFFF. returns #clusters that have a value
uint32_t ref_bvh_clusters(const void* h, void* nodes_out, uint32_t* parent, uint32_t* left,
                          uint32_t* right, uint8_t* is_leaf) {
    const RefBvh* r = static_cast<const RefBvh*>(h);
    const cr::BVH_Params& p = r->bvh->_InternalStruct;
    RefNode* nodes = static_cast<RefNode*>(nodes_out);
    uint32_t nc = 2 * r->n - 1, present = 0;
    for (uint32_t i = 0; i < nc; ++i) {
        if (p._Clusters[i].has_value()) { copy_node(p._Clusters[i].value(), nodes[i]); present++; }
        else std::memset(&nodes[i], 0xFF, sizeof(RefNode));
        parent[i] = p._Parent[i].has_value() ? p._Parent[i].value() : 0xFFFFFFFFu;
        left[i] = p._LeftChild[i].has_value() ? p._LeftChild[i].value() : 0xFFFFFFFFu;
        right[i] = p._RightChild[i].has_value() ? p._RightChild[i].value() : 0xFFFFFFFFu;
        is_leaf[i] = p._IsLeaf[i].has_value() ? 1 : 0;  // Q9: has_value() is what scene.cpp:193 tests
    }
    return present;
}

// OpenMP threads of the NEXT ref_bvh_build calls (bvh.cpp:67-147 has the reference's 8 parallel regions).  Cluster ids
// are only reproducible with 1 thread (Q3); the tree itself is the same for any count.
void ref_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
int ref_max_threads() { return omp_get_max_threads(); }

// The reference's pair sort, BVH::sortMortonCodesAndTriangleIndices (bvh.cpp:214-232), as a stage of its own for the
// per-stage CPU baseline: the method is private and only reachable through the constructor (which runs the whole
// build), so its five lines are restated -- pairs of (code, iota index), std::sort, unzip.  Returns milliseconds.
double ref_sort_pairs_ms(const uint32_t* codes, uint32_t n, uint32_t* codes_out, uint32_t* indices_out) {
    std::vector<std::pair<uint32_t, uint32_t>> v(n);
    for (uint32_t i = 0; i < n; ++i) v[i] = {codes[i], i};
    auto t0 = std::chrono::steady_clock::now();
    std::sort(v.begin(), v.end());
    auto t1 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < n; ++i) { codes_out[i] = v[i].first; indices_out[i] = v[i].second; }
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

}  // extern "C"
