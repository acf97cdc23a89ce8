// render_obj -- the reference's start-up frame without a window: Application::initScene + initCamera + one
// drawOneFrame (srcOpenGL/application.cpp:16-20, 176-186, 225-245) written against include/rtr_scene.hpp.
//
//   render_obj <model.obj> <out.ppm> [--size W H] [--scale s] [--rotate x y z] [--eye x y z] [--wireframe]
//              [--bvh-depth d] [--dump-camera file] [--dump-rgba file] [--frames n [--rebuild]]
//
// --frames n runs the reference's main loop n times first (mainLoop -> render -> drawOneFrame, application.cpp:220-313:
// dispatch the shader, _FPS.increment(), the statistics every 10 frames) with the camera stepping forward like a held
// W key (Camera::processKeyboard); --rebuild also rebuilds the BVH every frame, the call the reference keeps
// commented out at application.cpp:224.
//
// cr::Mesh::load -> cr::BVH (Morton, sort, PLOC, flatten on the GPU) -> tracePrimary (the compute shader's dispatch)
// -> rtr_shade (getColor) -> 8-bit PPM.  Everything the reference does between reading the OBJ file and
// imageStore, through the drop-in types; the image is what its window would show (rows top to bottom as the shader
// numbers them).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rtr_scene.hpp"

static void die(const char* msg) { std::fprintf(stderr, "render_obj: %s\n", msg); std::exit(2); }

int main(int argc, char** argv) {
    if (argc < 3) die("usage: render_obj <model.obj> <out.ppm> [--size W H] [--scale s] [--rotate x y z] [--eye x y z] "
                      "[--wireframe] [--bvh-depth d] [--dump-camera file] [--dump-rgba file]");
    uint32_t W = 1280, H = 720;  // ApplicationParameters defaults
    float scale = 1.f, rot[3] = {0.f, 0.f, 0.f}, eye[3] = {0.f, 0.f, -5.f};
    bool wireframe = false, rebuild = false;
    int bvhDepth = -1, frames = 0;
    const char* dumpCamera = nullptr;
    const char* dumpRgba = nullptr;
    for (int i = 3; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](int k) { if (i + k >= argc) die(("missing value after " + a).c_str()); };
        if (a == "--size") { need(2); W = (uint32_t)std::atoi(argv[i + 1]); H = (uint32_t)std::atoi(argv[i + 2]); i += 2; }
        else if (a == "--scale") { need(1); scale = (float)std::atof(argv[++i]); }
        else if (a == "--rotate") { need(3); for (int k = 0; k < 3; ++k) rot[k] = (float)std::atof(argv[++i]); }
        else if (a == "--eye") { need(3); for (int k = 0; k < 3; ++k) eye[k] = (float)std::atof(argv[++i]); }
        else if (a == "--wireframe") wireframe = true;
        else if (a == "--bvh-depth") { need(1); bvhDepth = std::atoi(argv[++i]); }
        else if (a == "--frames") { need(1); frames = std::atoi(argv[++i]); }
        else if (a == "--rebuild") rebuild = true;
        else if (a == "--dump-camera") { need(1); dumpCamera = argv[++i]; }
        else if (a == "--dump-rgba") { need(1); dumpRgba = argv[++i]; }
        else die(("unknown option " + a).c_str());
    }
    if (W == 0 || H == 0) die("empty image");

    // initScene: one material, one model that uses it
    std::vector<cr::MaterialGPU> materials = {cr::Material(cr::vec4{0.2f, 0.3f, 0.1f, 1.f})._InternalStruct};
    cr::MeshPtr model = cr::Mesh::load(argv[1]);
    model->setMaterial(0);
    if (rot[0] != 0.f || rot[1] != 0.f || rot[2] != 0.f) model->setRotation(rot[0], rot[1], rot[2]);
    if (scale != 1.f) model->setScale(scale);
    std::vector<cr::TriangleGPU> triangles;
    model->appendTo(triangles);
    for (cr::TriangleGPU& t : triangles) t._ModelId = 0;  // the scene's mesh slot (Mesh::_Id of the first mesh of a process)
    std::vector<cr::MeshModelGPU> meshes = {model->_InternalStruct};
    if (triangles.empty()) die("the model has no triangles");

    // initCamera
    const cr::Camera camera(eye, static_cast<float>(W) / static_cast<float>(H));
    const cr::CameraGPU cameraGPU = camera.getGpuData();
    if (dumpCamera) {
        FILE* f = std::fopen(dumpCamera, "wb");
        if (!f || std::fwrite(&cameraGPU, sizeof(cameraGPU), 1, f) != 1) die("cannot write the camera file");
        std::fclose(f);
    }

    // Scene::bindSSBO builds the BVH (scene.cpp:148); drawOneFrame dispatches the shader
    cr::BVH bvh(static_cast<uint32_t>(triangles.size()), triangles, meshes, nullptr, /*fillInternalStruct=*/false);
    if (frames > 0) {  // mainLoop: render() + _FPS.increment() + the statistics window (application.cpp:264-313)
        glr::ApplicationFPS fps;
        cr::Camera moving(eye, static_cast<float>(W) / static_cast<float>(H));
        fps.increment();
        for (int f = 0; f < frames; ++f) {
            if (rebuild) {
                cr::BVH again(static_cast<uint32_t>(triangles.size()), triangles, meshes, bvh.context(), /*fillInternalStruct=*/false);
                (void)again.tracePrimary(moving.getGpuData(), W, H);
            } else {
                (void)bvh.tracePrimary(moving.getGpuData(), W, H);
            }
            moving.processKeyboard(cr::FORWARD, 0.016f);
            fps.increment();
            fps.display(stdout);
        }
    }
    const std::vector<cr::Hit> hits = bvh.tracePrimary(cameraGPU, W, H);
    std::vector<float> rgba(static_cast<size_t>(W) * H * 4), overlay;
    uint32_t flags = wireframe ? RTR_SHADE_WIREFRAME : 0u;
    if (bvhDepth >= 0) {
        overlay.resize(rgba.size());
        RTR_SCENE_CHECK(bvh.context(), rtr_bvh_depth_overlay(bvh.context(), bvh.handle(), reinterpret_cast<const rtr_camera*>(&cameraGPU),
                                                             W, H, 0, 0, bvhDepth, overlay.data()));
        flags |= RTR_SHADE_BVH;
    }
    RTR_SCENE_CHECK(bvh.context(), rtr_shade(bvh.context(), reinterpret_cast<const rtr_hit*>(hits.data()), hits.size(),
                                             reinterpret_cast<const rtr_triangle*>(triangles.data()), static_cast<uint32_t>(triangles.size()),
                                             reinterpret_cast<const rtr_mesh*>(meshes.data()), 1u,
                                             reinterpret_cast<const rtr_material*>(materials.data()), 1u, flags,
                                             overlay.empty() ? nullptr : overlay.data(), rgba.data()));
    if (dumpRgba) {
        FILE* f = std::fopen(dumpRgba, "wb");
        if (!f || std::fwrite(rgba.data(), sizeof(float), rgba.size(), f) != rgba.size()) die("cannot write the rgba file");
        std::fclose(f);
    }
    FILE* f = std::fopen(argv[2], "wb");
    if (!f) die("cannot open the output file");
    std::fprintf(f, "P6\n%u %u\n255\n", W, H);
    std::vector<unsigned char> row(static_cast<size_t>(W) * 3);
    size_t hit = 0;
    for (uint32_t y = 0; y < H; ++y) {
        for (uint32_t x = 0; x < W; ++x) {
            const float* px = &rgba[(static_cast<size_t>(y) * W + x) * 4];
            for (int c = 0; c < 3; ++c) {
                const float v = px[c] < 0.f ? 0.f : px[c] > 1.f ? 1.f : px[c];
                row[3 * x + c] = static_cast<unsigned char>(v * 255.f + 0.5f);
            }
            hit += hits[static_cast<size_t>(y) * W + x]._DidHit ? 1u : 0u;
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
    std::printf("render_obj: %zu triangles, %u nodes, %ux%u, %zu pixels hit -> %s\n", triangles.size(),
                2u * static_cast<uint32_t>(triangles.size()) - 1u, W, H, hit, argv[2]);
    return 0;
}
