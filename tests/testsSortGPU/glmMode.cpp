// include/rtr_scene.hpp with RTR_SCENE_USE_GLM: the vector members are glm::vec3 / glm::vec4 / glm::mat4 exactly as in the
// reference's headers.  Compiled by tests/test_cpp_harness.py against the GLM of the reference's tree when it is there
// (not part of the Makefile: GLM is not in this repository).  Host-only calls; exits 0.
#define RTR_SCENE_USE_GLM
#include "rtr_scene.hpp"
int main() {
    cr::MeshPtr m = cr::Mesh::primitiveCube();
    m->setScale(2.f); m->setRotation(0.1f, 0.2f, 0.3f); m->setPosition(glm::vec3(1, 2, 3)); m->setModel(glm::mat4(1.f));
    cr::Material mat(glm::vec4(0.2f, 0.3f, 0.1f, 1.f));
    cr::Triangle t(glm::vec3(0, 1, 0), glm::vec3(-1, -1, 0), glm::vec3(1, -1, 0), 0);
    glm::vec3 c = cr::Triangle::getCentroid(t._InternalStruct, m->_InternalStruct._ModelMatrix);
    const float eye[3] = {0, 0, -5};
    cr::Camera cam(eye, 1.5f);
    cr::CameraGPU g = cam.getGpuData();
    return (int)(c.x + g._PlaneNear + mat._InternalStruct._Color.x + m->_Triangles.size()) == 12345;
}
