// ref_camera.cpp -- thin C-ABI driver around the REFERENCE's own cr::Camera (srcCommon/scene/camera.cpp compiled from
// where it lies under /root/reference with the vendored GLM; see oracle/Makefile).  TEST INFRASTRUCTURE ONLY: it is
// linked into oracle/_ref/libref_camera.so, which pins the cr::Camera mirror of include/rtr_scene.hpp
// (tests/testsSortGPU/testCamera.cpp).  No reference source is copied into the repo.
#include "camera.hpp"

#include <cstdint>
#include <cstring>

static_assert(sizeof(cr::CameraGPU) == 284, "CameraGPU layout");

extern "C" {

// Builds a camera like application.cpp:16-20 does, replays `n_ops` input events -- kind 0: ProcessMouseMovement(a, b),
// kind 1..6: processKeyboard(FORWARD..DOWN, a) with _Accelerate = (b != 0) -- and writes getGpuData() (284 bytes).
void ref_camera_gpu_data(const float eye[3], float aspect, float fov, float near_, float far_, int n_ops,
                         const int* kind, const float* a, const float* b, void* out) {
    cr::Camera cam(glm::vec3(eye[0], eye[1], eye[2]), aspect, fov, near_, far_);
    for (int i = 0; i < n_ops; ++i) {
        if (kind[i] == 0) cam.ProcessMouseMovement(a[i], b[i]);
        else {
            cam._Accelerate = b[i] != 0.f;
            cam.processKeyboard(static_cast<cr::CameraMovement>(kind[i] - 1), a[i]);
        }
    }
    const cr::CameraGPU g = cam.getGpuData();
    std::memcpy(out, &g, sizeof(g));
}

}  // extern "C"
