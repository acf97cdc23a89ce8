// testScene -- glr::Scene of include/rtr_scene.hpp (the mirror of srcOpenGL/scene/scene.hpp:22-62, scene.cpp), the parts
// that need no device: the container (materials, meshes, counters, the reference's caps), the three *ToGPUData views with
// and without the reference's fixed-size padding, and Scene::flattenTopDown -- the reference's recursiveTopDownTraversalBVH
// (scene.cpp:189-201) -- against the oracle's flatten on a tree the oracle's PLOC builds (oracle/librtr_oracle.so, test
// infrastructure; argv[1]).  The reference's own scene.cpp needs a GL context and cannot be compiled here.  CPU only.
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../oracle/rtr_oracle.h"
#include "rtr_scene.hpp"

static int bad = 0;
#define EXPECT(cond)                                                                   \
    do {                                                                               \
        if (!(cond)) { std::fprintf(stderr, "testScene:%d: %s\n", __LINE__, #cond); ++bad; } \
    } while (0)

static bool sameBytes(const void* a, const void* b, size_t n) { return std::memcmp(a, b, n) == 0; }

int main(int argc, char** argv) {
    int cases = 0;
    // ---- container, unpadded: what goes in comes out, in order ----
    {
        std::srand(11);
        glr::Scene scene;
        std::srand(11);
        const cr::Material sameDraws;  // the scene's default material took the same three rand() draws (material.cpp:17-27)
        EXPECT(scene.getNbMaterials() == 1 && scene.getNbMeshes() == 0 && scene.getNbTriangles() == 0);
        EXPECT(scene.getMaterialToGPUData().size() == 1);
        EXPECT(sameBytes(&scene.getMaterialToGPUData()[0], &sameDraws._InternalStruct, sizeof(cr::MaterialGPU)));
        scene.addMaterial(cr::vec4{0.2f, 0.3f, 0.1f, 1.f});
        scene.addRandomMaterial();
        EXPECT(scene.getNbMaterials() == 3 && scene.getMaterialToGPUData().size() == 3);
        EXPECT(scene.getMaterialToGPUData()[1]._Color.y == 0.3f);
        cr::MeshPtr cube = cr::Mesh::primitiveCube(), square = cr::Mesh::primitiveSquare();
        cube->setMaterial(1);
        square->setPosition(cr::vec3{1.f, 2.f, 3.f});
        scene.addMesh(cube);
        scene.addMesh(square);
        EXPECT(scene.getNbMeshes() == 2 && scene.getNbTriangles() == cube->_Triangles.size() + square->_Triangles.size());
        const std::vector<cr::TriangleGPU> tris = scene.getTriangleToGPUData();
        EXPECT(tris.size() == scene.getNbTriangles());
        for (size_t i = 0; i < cube->_Triangles.size(); ++i) EXPECT(sameBytes(&tris[i], &cube->_Triangles[i]._InternalStruct, 52));
        for (size_t i = 0; i < square->_Triangles.size(); ++i)
            EXPECT(sameBytes(&tris[cube->_Triangles.size() + i], &square->_Triangles[i]._InternalStruct, 52));
        const std::vector<cr::MeshModelGPU> models = scene.getMeshModelToGPUData();
        EXPECT(models.size() == 2 && models[0]._MaterialId == 1);
        EXPECT(sameBytes(&models[1], &square->_InternalStruct, sizeof(cr::MeshModelGPU)));
        ++cases;
    }
    // ---- the reference's fixed-size views and caps (scene.cpp:17-55) ----
    {
        glr::Scene scene(nullptr, /*referencePadding=*/true);
        EXPECT(scene.getMaterialToGPUData().size() == cr::Material::MAX_NB_MATERIALS);
        EXPECT(scene.getMaterialToGPUData()[5]._Color.x == 1.f && scene.getMaterialToGPUData()[5]._Color.w == 1.f);  // material.hpp:10
        EXPECT(scene.getTriangleToGPUData().size() == cr::Triangle::MAX_NB_TRIANGLES);
        EXPECT(scene.getMeshModelToGPUData().size() == cr::Mesh::MAX_NB_MESHES);
        for (size_t m = 0; m < cr::Mesh::MAX_NB_MESHES + 3; ++m) scene.addMesh(cr::Mesh::primitiveSquare());
        EXPECT(scene.getNbMeshes() == cr::Mesh::MAX_NB_MESHES);          // the 65th .. 67th are dropped (scene.cpp:51)
        EXPECT(scene.getNbTriangles() == 2 * cr::Mesh::MAX_NB_MESHES);
        const std::vector<cr::TriangleGPU> tris = scene.getTriangleToGPUData();
        EXPECT(tris.size() == cr::Triangle::MAX_NB_TRIANGLES);
        EXPECT(tris[2 * cr::Mesh::MAX_NB_MESHES - 1]._P0.w == 1.f && tris[2 * cr::Mesh::MAX_NB_MESHES]._P0.w == 0.f);  // zero padding
        // a mesh larger than the triangle cap is cut there (scene.cpp:30-33), and counted as the cap (scene.cpp:54)
        glr::Scene big(nullptr, true);
        cr::MeshPtr huge = cr::Mesh::primitiveSquare();
        huge->_Triangles.resize(cr::Triangle::MAX_NB_TRIANGLES + 10, huge->_Triangles[0]);
        big.addMesh(huge);
        EXPECT(big.getNbTriangles() == cr::Triangle::MAX_NB_TRIANGLES && big.getTriangleToGPUData().size() == cr::Triangle::MAX_NB_TRIANGLES);
        glr::Scene all;                                                    // without the padding nothing is dropped
        all.addMesh(huge);
        EXPECT(all.getNbTriangles() == cr::Triangle::MAX_NB_TRIANGLES + 10 && all.getTriangleToGPUData().size() == all.getNbTriangles());
        ++cases;
    }
    // ---- flattenTopDown on hand-made trees: one leaf, and ((A B) C) ----
    {
        cr::BVH_Params p;
        cr::BVH_NodeGPU leaf{};
        leaf._TriangleId = 7;
        p._Clusters = {leaf};
        p._IsLeaf = {true};
        p._LeftChild = {std::nullopt};
        p._RightChild = {std::nullopt};
        const auto one = glr::Scene::flattenTopDown(p, 1);
        EXPECT(one.size() == 1 && one[0]._TriangleId == 7 && one[0]._LeftChild == 0 && one[0]._RightChild == 0);
        cr::BVH_Params q;  // clusters 0,1,2 leaves A,B,C; 3 = (A B); 4 = root (3 C)
        q._Clusters.assign(5, cr::BVH_NodeGPU{});
        for (uint32_t c = 0; c < 5; ++c) q._Clusters[c].value()._TriangleId = 100 + c;
        q._IsLeaf = {true, true, true, std::nullopt, std::nullopt};
        q._LeftChild = {std::nullopt, std::nullopt, std::nullopt, 0u, 3u};
        q._RightChild = {std::nullopt, std::nullopt, std::nullopt, 1u, 2u};
        const auto f = glr::Scene::flattenTopDown(q, 3);
        EXPECT(f.size() == 5);
        EXPECT(f[0]._TriangleId == 104 && f[0]._LeftChild == 1 && f[0]._RightChild == 4);
        EXPECT(f[1]._TriangleId == 103 && f[1]._LeftChild == 2 && f[1]._RightChild == 3);
        EXPECT(f[2]._TriangleId == 100 && f[3]._TriangleId == 101 && f[4]._TriangleId == 102);
        EXPECT(glr::Scene::flattenTopDown(q, 0).empty());
        ++cases;
    }
    // ---- flattenTopDown against the oracle's flatten on PLOC trees ----
    const char* path = argc > 1 ? argv[1] : "oracle/librtr_oracle.so";
    void* lib = dlopen(path, RTLD_NOW);
    if (!lib) {
        std::fprintf(stderr, "oracle library not found (%s): PLOC-tree comparison skipped\n", path);
    } else {
        auto build = reinterpret_cast<decltype(&orc_bvh_build)>(dlsym(lib, "orc_bvh_build"));
        auto destroy = reinterpret_cast<decltype(&orc_bvh_destroy)>(dlsym(lib, "orc_bvh_destroy"));
        auto flatten = reinterpret_cast<decltype(&orc_flatten)>(dlsym(lib, "orc_flatten"));
        if (!build || !destroy || !flatten) { std::fprintf(stderr, "oracle symbols missing\n"); return 2; }
        std::mt19937 rng(5);
        std::uniform_real_distribution<float> pos(-10.f, 10.f), edge(-0.4f, 0.4f);
        for (uint32_t n : {1u, 2u, 3u, 17u, 1000u, 20000u}) {
            std::vector<cr::TriangleGPU> tris(n);
            for (auto& t : tris) {
                std::memset(static_cast<void*>(&t), 0, sizeof(t));
                const float x = pos(rng), y = pos(rng), z = pos(rng);
                t._P0 = cr::vec4{x, y, z, 1.f};
                t._P1 = cr::vec4{x + edge(rng), y + edge(rng), z + edge(rng), 1.f};
                t._P2 = cr::vec4{x + edge(rng), y + edge(rng), z + edge(rng), 1.f};
            }
            const cr::MeshModelGPU model{};
            orc_bvh* b = build(reinterpret_cast<const orc_triangle*>(tris.data()), n, n, reinterpret_cast<const orc_mesh*>(&model), 1, 16);
            if (!b) { std::fprintf(stderr, "orc_bvh_build failed for n = %u\n", n); ++bad; continue; }
            const uint32_t nc = 2 * n - 1;
            cr::BVH_Params p;
            p._Clusters.assign(nc, std::nullopt); p._IsLeaf.assign(nc, std::nullopt);
            p._LeftChild.assign(nc, std::nullopt); p._RightChild.assign(nc, std::nullopt);
            for (uint32_t c = 0; c < nc; ++c) {
                cr::BVH_NodeGPU nd;
                std::memcpy(static_cast<void*>(&nd), &b->clusters[c], sizeof(nd));
                p._Clusters[c] = nd;
                if (b->left[c] == 0xFFFFFFFFu) p._IsLeaf[c] = true;
                else { p._LeftChild[c] = b->left[c]; p._RightChild[c] = b->right[c]; }
            }
            std::vector<orc_node> expect(nc);
            flatten(b->clusters, b->left, b->right, n, expect.data());
            const std::vector<cr::BVH_NodeGPU> got = glr::Scene::flattenTopDown(p, n);
            EXPECT(got.size() == nc);
            if (got.size() == nc) {
                for (uint32_t i = 0; i < nc; ++i) {
                    const cr::BVH_NodeGPU& g = got[i];
                    const orc_node& e = expect[i];
                    if (!(sameBytes(&g._BoundingBox._Min, e.bmin, 12) && sameBytes(&g._BoundingBox._Max, e.bmax, 12) &&
                          g._TriangleId == e.triangle_id && g._LeftChild == e.left && g._RightChild == e.right)) {
                        std::fprintf(stderr, "n = %u: flat node %u differs from the oracle's\n", n, i);
                        ++bad;
                        break;
                    }
                }
            }
            destroy(b);
            ++cases;
        }
    }
    std::printf("testScene: %d cases, %d mismatches\n", cases, bad);
    return bad ? 1 : 0;
}
