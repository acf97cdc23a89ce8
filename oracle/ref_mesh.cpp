// ref_mesh.cpp -- thin C-ABI driver around the REFERENCE's own cr::Mesh (srcCommon/scene/geometry/mesh.cpp +
// triangle.cpp + core/errorHandler.cpp compiled from where they lie under /root/reference, with the vendored GLM and
// the tinyobjloader 1.2.0 of the reference's tree behind oracle/shim/tiny_obj_loader.h; see oracle/Makefile).
// TEST INFRASTRUCTURE ONLY: linked into oracle/_ref/libref_mesh.so, which pins rtr_obj_load / rtr_mesh_* and the
// cr::Mesh mirror of include/rtr_scene.hpp.  No reference source is copied into the repo.
#define TINYOBJLOADER_IMPLEMENTATION
#include <tiny_obj_loader.h>

#include "mesh.hpp"

#include <cstdint>
#include <cstring>

static_assert(sizeof(cr::TriangleGPU) == 64, "TriangleGPU layout");
static_assert(sizeof(cr::MeshModelGPU) == 68, "MeshModelGPU layout");

// errorHandler.cpp calls glfwTerminate() in a branch this path never takes; the symbol only has to exist.
extern "C" void glfwTerminate(void) {}

static int64_t copy_out(const cr::MeshPtr& m, void* out, int64_t cap, uint32_t* model_id) {
    const int64_t n = static_cast<int64_t>(m->_Triangles.size());
    if (model_id) *model_id = n ? m->_Triangles[0]._InternalStruct._ModelId : 0xffffffffu;
    for (int64_t i = 0; i < n && i < cap; ++i)
        std::memcpy(static_cast<char*>(out) + 64 * i, &m->_Triangles[size_t(i)]._InternalStruct, 64);
    return n;
}

extern "C" {

// Mesh::load(path) (mesh.cpp:186-263): number of triangles; the first min(n, cap) TriangleGPU are written to `out`.
// A load error is FATAL in the reference (ErrorHandler exits the process): call it from a child process to see that.
int64_t ref_mesh_load(const char* path, void* out, int64_t cap, uint32_t* model_id) {
    return copy_out(cr::Mesh::load(path), out, cap, model_id);
}

// which: 0 primitiveTriangle, 1 primitiveSquare, 2 primitiveCube, 3 primitiveSphere (mesh.cpp:64-183)
int64_t ref_mesh_primitive(int which, void* out, int64_t cap, uint32_t* model_id) {
    cr::MeshPtr m = which == 0 ? cr::Mesh::primitiveTriangle() : which == 1 ? cr::Mesh::primitiveSquare()
                  : which == 2 ? cr::Mesh::primitiveCube() : cr::Mesh::primitiveSphere();
    return copy_out(m, out, cap, model_id);
}

// Replays transform calls on a fresh Mesh and writes its MeshModelGPU (68 bytes).
// kind 0: setPosition(a,b,c); 1: setScale(a); 2: setRotation(a,b,c); 3: setMaterial(uint(a)); 4: setModel(mat16 at a-index
// into `mats`, i.e. mats + 16*int(a)).
void ref_mesh_transform(int n_ops, const int* kind, const float* a, const float* b, const float* c, const float* mats,
                        void* out) {
    cr::Mesh m;
    for (int i = 0; i < n_ops; ++i) {
        switch (kind[i]) {
            case 0: m.setPosition(glm::vec3(a[i], b[i], c[i])); break;
            case 1: m.setScale(a[i]); break;
            case 2: m.setRotation(a[i], b[i], c[i]); break;
            case 3: m.setMaterial(static_cast<uint32_t>(a[i])); break;
            case 4: { glm::mat4 mm; std::memcpy(&mm, mats + 16 * static_cast<int>(a[i]), 64); m.setModel(mm); break; }
        }
    }
    std::memcpy(out, &m._InternalStruct, 68);
}

// Triangle::getCentroid(triangle, model) (triangle.cpp:30-32) -- the centroid the builder feeds the Morton codes with.
void ref_triangle_centroid(const void* tri, const float* model16, float out[3]) {
    cr::TriangleGPU t; std::memcpy(&t, tri, 64);
    glm::mat4 m; std::memcpy(&m, model16, 64);
    const glm::vec3 c = cr::Triangle::getCentroid(t, m);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

}  // extern "C"
