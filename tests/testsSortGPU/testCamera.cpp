// cr::Camera of include/rtr_scene.hpp against the REFERENCE's own camera.cpp (oracle/_ref/libref_camera.so, built by
// oracle/Makefile from /root/reference): CameraGPU must be bit-identical -- view (glm::lookAt), projection
// (glm::perspective), both inverses (glm::inverse), eye and the image-plane sizes -- for the application's start-up
// camera (application.cpp:16-20), a sweep of positions / aspect ratios / fields of view, and replayed mouse and
// keyboard input.  CPU only; exits 77 when the reference library is not there.
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "rtr_scene.hpp"

using RefFn = void (*)(const float*, float, float, float, float, int, const int*, const float*, const float*, void*);

static int compare(RefFn ref, const float eye[3], float aspect, float fov, float near_, float far_,
                   const std::vector<int>& kind, const std::vector<float>& a, const std::vector<float>& b) {
    unsigned char expect[sizeof(cr::CameraGPU)];
    ref(eye, aspect, fov, near_, far_, (int)kind.size(), kind.data(), a.data(), b.data(), expect);
    cr::Camera cam(eye, aspect, fov, near_, far_);
    for (size_t i = 0; i < kind.size(); ++i) {
        if (kind[i] == 0) cam.ProcessMouseMovement(a[i], b[i]);
        else { cam._Accelerate = b[i] != 0.f; cam.processKeyboard(static_cast<cr::CameraMovement>(kind[i] - 1), a[i]); }
    }
    const cr::CameraGPU got = cam.getGpuData();
    if (std::memcmp(&got, expect, sizeof(got)) == 0) return 0;
    const float* g = reinterpret_cast<const float*>(&got);
    const float* e = reinterpret_cast<const float*>(expect);
    for (size_t i = 0; i < sizeof(got) / 4; ++i)
        if (std::memcmp(g + i, e + i, 4) != 0)
            std::fprintf(stderr, "CameraGPU word %zu: got %.9g, reference %.9g (eye %g %g %g aspect %g fov %g, %zu events)\n", i,
                         g[i], e[i], eye[0], eye[1], eye[2], aspect, fov, kind.size());
    return 1;
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "oracle/_ref/libref_camera.so";
    void* lib = dlopen(path, RTLD_NOW);
    if (!lib) { std::fprintf(stderr, "reference camera library not found (%s): skipped\n", path); return 77; }
    RefFn ref = reinterpret_cast<RefFn>(dlsym(lib, "ref_camera_gpu_data"));
    if (!ref) { std::fprintf(stderr, "ref_camera_gpu_data missing\n"); return 2; }
    int bad = 0, cases = 0;
    {   // the application's camera: eye (0,0,-5), 1280/720, fov 45, near 0.1, far 200
        const float eye[3] = {0.f, 0.f, -5.f};
        bad += compare(ref, eye, 1280.f / 720.f, 45.f, 0.1f, 200.f, {}, {}, {}); ++cases;
    }
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> pos(-300.f, 300.f), asp(0.5f, 2.5f), fov(20.f, 100.f), mouse(-400.f, 400.f), dt(0.001f, 0.05f);
    for (int c = 0; c < 2000; ++c) {
        const float eye[3] = {pos(rng), pos(rng), pos(rng)};
        const float aspect = asp(rng), f = fov(rng), near_ = 0.05f + 0.5f * dt(rng), far_ = 100.f + std::fabs(pos(rng));
        std::vector<int> kind; std::vector<float> a, b;
        const int events = c % 7;
        for (int i = 0; i < events; ++i) {
            const int k = (int)(rng() % 7);
            kind.push_back(k);
            if (k == 0) { a.push_back(mouse(rng)); b.push_back(mouse(rng)); }
            else { a.push_back(dt(rng)); b.push_back((rng() & 1) ? 1.f : 0.f); }
        }
        bad += compare(ref, eye, aspect, f, near_, far_, kind, a, b); ++cases;
        if (bad > 5) break;
    }
    std::printf("testCamera: %d cases, %d mismatches\n", cases, bad);
    return bad ? 1 : 0;
}
