// cr::Mesh / cr::Triangle of include/rtr_scene.hpp (on rtr_obj_load, rtr_mesh_*, rtr_triangle_centroid) against the
// REFERENCE's own mesh.cpp + triangle.cpp (oracle/_ref/libref_mesh.so, built by oracle/Makefile from /root/reference
// against the tinyobjloader 1.2.0 of its tree): Mesh::load on every tests/golden/obj/*.obj given on the command line,
// the primitives, and replayed setPosition / setScale / setRotation / setMaterial calls -- TriangleGPU and MeshModelGPU
// byte for byte (but for _ModelId, a process-wide counter on both sides).  Host only: no device is touched.
//   testMesh <libref_mesh.so> <file.obj>...        exits 77 when the reference library is not there.
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "rtr_scene.hpp"

using LoadFn = int64_t (*)(const char*, void*, int64_t, uint32_t*);
using PrimFn = int64_t (*)(int, void*, int64_t, uint32_t*);
using XformFn = void (*)(int, const int*, const float*, const float*, const float*, const float*, void*);
using CentroidFn = void (*)(const void*, const float*, float*);

static int compare_triangles(const char* what, const cr::MeshPtr& mesh, const std::vector<cr::TriangleGPU>& ref, int64_t n) {
    if ((int64_t)mesh->_Triangles.size() != n) {
        std::fprintf(stderr, "%s: %zu triangles, reference %lld\n", what, mesh->_Triangles.size(), (long long)n);
        return 1;
    }
    for (int64_t i = 0; i < n; ++i) {
        const cr::TriangleGPU& g = mesh->_Triangles[(size_t)i]._InternalStruct;
        if (g._ModelId != mesh->getId() || std::memcmp(&g, &ref[(size_t)i], 48) != 0) {
            std::fprintf(stderr, "%s: triangle %lld differs\n", what, (long long)i);
            return 1;
        }
    }
    return 0;
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "oracle/_ref/libref_mesh.so";
    void* lib = dlopen(path, RTLD_NOW);
    if (!lib) { std::fprintf(stderr, "reference mesh library not found (%s): skipped\n", path); return 77; }
    LoadFn ref_load = reinterpret_cast<LoadFn>(dlsym(lib, "ref_mesh_load"));
    PrimFn ref_prim = reinterpret_cast<PrimFn>(dlsym(lib, "ref_mesh_primitive"));
    XformFn ref_xform = reinterpret_cast<XformFn>(dlsym(lib, "ref_mesh_transform"));
    CentroidFn ref_centroid = reinterpret_cast<CentroidFn>(dlsym(lib, "ref_triangle_centroid"));
    if (!ref_load || !ref_prim || !ref_xform || !ref_centroid) { std::fprintf(stderr, "reference entry points missing\n"); return 2; }
    int bad = 0, cases = 0;
    std::vector<cr::TriangleGPU> ref(1 << 19);

    for (int i = 2; i < argc; ++i, ++cases) {  // Mesh::load
        uint32_t id = 0;
        const int64_t n = ref_load(argv[i], ref.data(), (int64_t)ref.size(), &id);
        bad += compare_triangles(argv[i], cr::Mesh::load(argv[i]), ref, n);
    }
    for (int which = 0; which < 4; ++which, ++cases) {  // primitives
        uint32_t id = 0;
        const int64_t n = ref_prim(which, ref.data(), (int64_t)ref.size(), &id);
        cr::MeshPtr m = which == 0 ? cr::Mesh::primitiveTriangle() : which == 1 ? cr::Mesh::primitiveSquare()
                      : which == 2 ? cr::Mesh::primitiveCube() : cr::Mesh::primitiveSphere();
        bad += compare_triangles("primitive", m, ref, n);
    }
    std::mt19937 rng(11);
    std::uniform_real_distribution<float> ang(-7.f, 7.f), pos(-100.f, 100.f), scl(0.01f, 20.f);
    for (int c = 0; c < 2000; ++c, ++cases) {  // setters, then the centroid under the resulting matrix
        std::vector<int> kind; std::vector<float> a, b, d;
        const int events = c % 9;
        cr::Mesh mesh;
        for (int i = 0; i < events; ++i) {
            const int k = (int)(rng() % 4);
            kind.push_back(k);
            if (k == 0) { a.push_back(pos(rng)); b.push_back(pos(rng)); d.push_back(pos(rng)); mesh.setPosition(cr::vec3{a.back(), b.back(), d.back()}); }
            else if (k == 1) { a.push_back(scl(rng)); b.push_back(0.f); d.push_back(0.f); mesh.setScale(a.back()); }
            else if (k == 2) { a.push_back(ang(rng)); b.push_back(ang(rng)); d.push_back(ang(rng)); mesh.setRotation(a.back(), b.back(), d.back()); }
            else { a.push_back((float)(rng() % 64)); b.push_back(0.f); d.push_back(0.f); mesh.setMaterial((uint32_t)a.back()); }
        }
        unsigned char expect[68];
        ref_xform(events, kind.data(), a.data(), b.data(), d.data(), nullptr, expect);
        if (std::memcmp(&mesh._InternalStruct, expect, 68) != 0) {
            std::fprintf(stderr, "MeshModelGPU differs after %d setter calls (case %d)\n", events, c);
            ++bad;
        }
        const cr::Triangle t(cr::vec3{pos(rng), pos(rng), pos(rng)}, cr::vec3{pos(rng), pos(rng), pos(rng)}, cr::vec3{pos(rng), pos(rng), pos(rng)}, 0);
        float want[3];
        ref_centroid(&t._InternalStruct, reinterpret_cast<const float*>(&mesh._InternalStruct._ModelMatrix), want);
        const cr::vec3 got = cr::Triangle::getCentroid(t._InternalStruct, mesh._InternalStruct._ModelMatrix);
        if (std::memcmp(&got, want, 12) != 0) {
            std::fprintf(stderr, "centroid differs (case %d): %.9g %.9g %.9g vs %.9g %.9g %.9g\n", c, got.x, got.y, got.z, want[0], want[1], want[2]);
            ++bad;
        }
        if (bad > 5) break;
    }
    std::printf("testMesh: %d cases, %d mismatches\n", cases, bad);
    return bad ? 1 : 0;
}
