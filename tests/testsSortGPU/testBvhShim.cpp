// The C++ shim include/rtr_scene.hpp driven the way srcOpenGL/scene/scene.cpp drives cr::BVH:
// construct cr::BVH(nbTriangles, triangles, meshes) (scene.cpp:148), walk _InternalStruct from the root
// cluster 2n-2 in DFS pre-order relabelling nodes (the consumer loop of scene.cpp:189-208), and compare
// with the flat array the library produced on the GPU.  Then one primary-ray dispatch.
#include <cmath>
#include <random>

#include "harness.hpp"
#include "rtr_scene.hpp"

// host-side consumer of BVH_Params, same contract as glr::Scene::getBVH_NodesToGPUData
static std::vector<cr::BVH_NodeGPU> flattenOnHost(const cr::BVH& bvh) {
    const cr::BVH_Params& p = bvh._InternalStruct;
    std::vector<cr::BVH_NodeGPU> out;
    struct Item { uint32_t cluster; int parentSlot; bool isRight; };
    std::vector<Item> stack{{static_cast<uint32_t>(2 * p._NbTriangles - 2), -1, false}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const uint32_t slot = static_cast<uint32_t>(out.size());
        out.push_back(p._Clusters[it.cluster].value());
        if (it.parentSlot >= 0) (it.isRight ? out[it.parentSlot]._RightChild : out[it.parentSlot]._LeftChild) = slot;
        if (!p._IsLeaf[it.cluster].has_value()) {  // inner cluster (Q9)
            stack.push_back({p._RightChild[it.cluster].value(), static_cast<int>(slot), true});
            stack.push_back({p._LeftChild[it.cluster].value(), static_cast<int>(slot), false});  // left is emitted first
        }
    }
    return out;
}

int main() {
    const uint32_t n = 5000;
    std::mt19937 gen(11);
    std::uniform_real_distribution<float> centre(-4.f, 4.f), offset(-0.25f, 0.25f);
    std::vector<cr::TriangleGPU> tris(n);
    for (auto& t : tris) {
        const float cx = centre(gen), cy = centre(gen), cz = centre(gen);
        cr::vec4* v[3] = {&t._P0, &t._P1, &t._P2};
        for (auto* p : v) { p->x = cx + offset(gen); p->y = cy + offset(gen); p->z = cz + offset(gen); p->w = 1.f; }
        t._ModelId = 0;
    }
    std::vector<cr::MeshModelGPU> meshes(1);

    cr::BVH_Ptr bvh(new cr::BVH(n, tris, meshes));
    const cr::BVH_Params& p = bvh->_InternalStruct;
    assert(p._Clusters.size() == 2 * n - 1 && p._TriangleIndices.size() == n);
    for (uint32_t i = 0; i < n; ++i) {  // leaves: Morton order, codes ascending
        assert(p._IsLeaf[i].has_value() && p._Clusters[i]->_TriangleId == p._TriangleIndices[i]);
        if (i) assert(p._MortonCodes[i - 1] <= p._MortonCodes[i]);
    }
    assert(!p._Parent[2 * n - 2].has_value());  // the root has no parent

    const std::vector<cr::BVH_NodeGPU> expected = flattenOnHost(*bvh), flat = bvh->getFlatNodes();
    assert(flat.size() == expected.size());
    for (size_t i = 0; i < flat.size(); ++i) {
        assert(std::memcmp(&flat[i]._BoundingBox._Min, &expected[i]._BoundingBox._Min, 12) == 0);
        assert(std::memcmp(&flat[i]._BoundingBox._Max, &expected[i]._BoundingBox._Max, 12) == 0);
        assert(flat[i]._TriangleId == expected[i]._TriangleId && flat[i]._LeftChild == expected[i]._LeftChild &&
               flat[i]._RightChild == expected[i]._RightChild);
    }
    std::fprintf(stderr, "flat array of %zu nodes equals the host-side walk of _InternalStruct\n", flat.size());

    // one frame of primary rays from (0,0,-15) looking down +z
    cr::CameraGPU cam;
    std::memset(&cam, 0, sizeof(cam));
    cam._View = cr::mat4::identity(); cam._Proj = cr::mat4::identity(); cam._InvProj = cr::mat4::identity();
    cam._InvView = cr::mat4::identity();
    cam._InvView.m[14] = -15.f;  // camera-to-world translation (column-major)
    cam._Eye = {0.f, 0.f, -15.f, 1.f};
    const float near = 0.1f, fov = 45.f * 3.14159265f / 180.f;
    const uint32_t W = 128, H = 96;
    cam._PlaneHeight = 2.f * near * std::tan(0.5f * fov);
    cam._PlaneWidth = cam._PlaneHeight * float(W) / float(H);
    cam._PlaneNear = near;
    const std::vector<cr::Hit> hits = bvh->tracePrimary(cam, W, H);
    size_t nbHits = 0;
    for (const cr::Hit& h : hits)
        if (h._DidHit) { ++nbHits; assert(h._TriangleId < n && h._Coords.w > 0.f); }
    std::fprintf(stderr, "%zu of %zu primary rays hit\n", nbHits, hits.size());
    assert(nbHits > hits.size() / 20);

    // glr::Scene (scene.cpp): the same triangles as the one mesh of a scene -> sendDataToGpu (the reference's bindSSBO
    // builds the BVH there, scene.cpp:148) -> the same tree through getBVH_NodesToGPUData and through the reference's
    // own top-down walk of _InternalStruct, and the frame of drawOneFrame = getColor over the primary hits
    {
        glr::ScenePtr scene(new glr::Scene());
        scene->addMaterial(cr::vec4{0.2f, 0.3f, 0.1f, 1.f});
        cr::MeshPtr mesh(new cr::Mesh());
        for (const cr::TriangleGPU& t : tris) mesh->_Triangles.emplace_back(t);
        mesh->setMaterial(1);
        scene->addMesh(mesh);
        assert(scene->getNbTriangles() == n && scene->getNbMeshes() == 1 && scene->getNbMaterials() == 2);
        scene->sendDataToGpu();
        const std::vector<cr::BVH_NodeGPU> sent = scene->getBVH_NodesToGPUData(scene->getBVH());
        const std::vector<cr::BVH_NodeGPU> walked = glr::Scene::flattenTopDown(scene->getBVH()->_InternalStruct, n);
        assert(sent.size() == flat.size() && walked.size() == flat.size());
        for (size_t i = 0; i < flat.size(); ++i) {
            assert(std::memcmp(&sent[i]._BoundingBox._Min, &flat[i]._BoundingBox._Min, 12) == 0 &&
                   std::memcmp(&sent[i]._BoundingBox._Max, &flat[i]._BoundingBox._Max, 12) == 0 &&
                   sent[i]._TriangleId == flat[i]._TriangleId && sent[i]._LeftChild == flat[i]._LeftChild &&
                   sent[i]._RightChild == flat[i]._RightChild);
            assert(std::memcmp(&walked[i]._BoundingBox._Min, &flat[i]._BoundingBox._Min, 12) == 0 &&
                   std::memcmp(&walked[i]._BoundingBox._Max, &flat[i]._BoundingBox._Max, 12) == 0 &&
                   walked[i]._TriangleId == flat[i]._TriangleId && walked[i]._LeftChild == flat[i]._LeftChild &&
                   walked[i]._RightChild == flat[i]._RightChild);
        }
        const std::vector<float> frame = scene->drawOneFrame(cam, W, H);
        const std::vector<cr::MaterialGPU> materials = scene->getMaterialToGPUData();
        const std::vector<cr::MeshModelGPU> models = scene->getMeshModelToGPUData();
        std::vector<float> expect(frame.size(), 0.f);
        HARNESS_CHECK(bvh->context(), rtr_shade(bvh->context(), reinterpret_cast<const rtr_hit*>(hits.data()), hits.size(),
                                                reinterpret_cast<const rtr_triangle*>(tris.data()), n,
                                                reinterpret_cast<const rtr_mesh*>(models.data()), 1,
                                                reinterpret_cast<const rtr_material*>(materials.data()), 2, 0u, nullptr, expect.data()));
        assert(frame.size() == static_cast<size_t>(W) * H * 4 && std::memcmp(frame.data(), expect.data(), frame.size() * 4) == 0);
        size_t lit = 0;
        for (size_t px = 0; px < frame.size(); px += 4) lit += frame[px] != 0.f || frame[px + 1] != 0.f || frame[px + 2] != 0.f;
        assert(lit > hits.size() / 20);
        std::fprintf(stderr, "glr::Scene: same %zu nodes, frame of %zu lit pixels equals getColor over the hits\n", sent.size(), lit);
    }
    return EXIT_SUCCESS;
}
