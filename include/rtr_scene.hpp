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
//   cr::Triangle, cr::Mesh (load / primitive* / set*)       <- triangle.hpp:16-34, mesh.hpp:17-44
//   cr::Camera                                              <- srcCommon/scene/camera.hpp:35-106
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

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <optional>
#include <string>
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
    mat4 _ModelMatrix = mat4(1.f);
#else
    mat4 _ModelMatrix = mat4::identity();
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

// cr::Camera (srcCommon/scene/camera.hpp:35-106, camera.cpp): produces the CameraGPU the traversal consumes.
// Host-side scalar code like the reference's; the arithmetic follows GLM 0.9.9.9 operation by operation
// (lookAtRH matrix_transform.inl:153-173, perspectiveRH_NO matrix_clip_space.inl:249-262, inverse
// func_matrix.inl:347-405, normalize = v * (1 / sqrt(dot)), dot = (x + y) + z, radians = deg * 0.0174532925...f)
// so that getGpuData() is bit-identical to the reference's on the same libm
// (tests/testsSortGPU/testCamera.cpp and tests/test_camera_cpu.py compare against the reference's own camera.cpp,
// oracle/_ref/libref_camera.so).
// Compile without FMA contraction (-ffp-contract=off, or no -march that enables FMA).
enum CameraMovement { FORWARD, BACKWARD, LEFT, RIGHT, UP, DOWN };

class Camera {
public:
    bool _Accelerate = false;

    Camera(const float position[3], float aspectRatio, float fov = 45.f, float near_ = 0.1f, float far_ = 200.f,
           const float* worldUp = nullptr) {
        _AspectRatio = aspectRatio; _Fov = fov; _Near = near_; _Far = far_;
        _WorldUp[0] = worldUp ? worldUp[0] : 0.f; _WorldUp[1] = worldUp ? worldUp[1] : 1.f; _WorldUp[2] = worldUp ? worldUp[2] : 0.f;
        for (int k = 0; k < 3; ++k) _Eye[k] = position[k];
        updateCameraVectors();
    }

    CameraGPU getGpuData() const {  // camera.cpp:23-34
        CameraGPU out;
        std::memset(&out, 0, sizeof(out));
        float view[16], proj[16], inv[16];
        getView(view); getPerspective(proj);
        std::memcpy(&out._View, view, 64);
        std::memcpy(&out._Proj, proj, 64);
        inverse4(view, inv); std::memcpy(&out._InvView, inv, 64);
        inverse4(proj, inv); std::memcpy(&out._InvProj, inv, 64);
        const float eye[4] = {_Eye[0], _Eye[1], _Eye[2], 1.f};
        std::memcpy(&out._Eye, eye, 16);
        out._PlaneHeight = getPlaneHeight();
        out._PlaneWidth = out._PlaneHeight * _AspectRatio;  // getPlaneWidth(planeHeight), camera.cpp:64-66
        out._PlaneNear = _Near;
        return out;
    }
    void getView(float m[16]) const {  // glm::lookAt(_Eye, _Eye + _At, _Up), right-handed
        const float center[3] = {_Eye[0] + _At[0], _Eye[1] + _At[1], _Eye[2] + _At[2]};
        float f[3] = {center[0] - _Eye[0], center[1] - _Eye[1], center[2] - _Eye[2]}, s[3], u[3];
        normalize3(f);
        cross3(f, _Up, s); normalize3(s);
        cross3(s, f, u);
        for (int i = 0; i < 16; ++i) m[i] = 0.f;
        m[15] = 1.f;
        m[0] = s[0]; m[4] = s[1]; m[8] = s[2];
        m[1] = u[0]; m[5] = u[1]; m[9] = u[2];
        m[2] = -f[0]; m[6] = -f[1]; m[10] = -f[2];
        m[12] = -dot3(s, _Eye); m[13] = -dot3(u, _Eye); m[14] = dot3(f, _Eye);
    }
    void getPerspective(float m[16]) const {  // glm::perspective(radians(_Fov), aspect, near, far), RH, depth -1..1
        const float fovy = radians(_Fov);
        const float tanHalfFovy = std::tan(fovy / 2.f);
        for (int i = 0; i < 16; ++i) m[i] = 0.f;
        m[0] = 1.f / (_AspectRatio * tanHalfFovy);
        m[5] = 1.f / tanHalfFovy;
        m[10] = -(_Far + _Near) / (_Far - _Near);
        m[11] = -1.f;
        m[14] = -(2.f * _Far * _Near) / (_Far - _Near);
    }
    float getPlaneHeight() const { return 2.f * _Near * std::tan(0.5f * radians(_Fov)); }  // camera.cpp:52-54

    void processKeyboard(CameraMovement direction, float deltaTime) {  // camera.cpp:68-93
        float velocity = 20.f * deltaTime;
        if (_Accelerate) velocity *= 5.f;
        const float* v = direction <= BACKWARD ? _At : (direction <= RIGHT ? _Right : _WorldUp);
        const bool minus = direction == FORWARD || direction == LEFT || direction == DOWN;
        for (int k = 0; k < 3; ++k) { const float d = v[k] * velocity; _Eye[k] = minus ? _Eye[k] - d : _Eye[k] + d; }
    }
    void ProcessMouseMovement(float xoffset, float yoffset, bool constrainPitch = true) {  // camera.cpp:95-109
        xoffset *= 0.1f; yoffset *= 0.1f;
        _Yaw += xoffset; _Pitch += yoffset;
        if (constrainPitch) { if (_Pitch > 89.0f) _Pitch = 89.0f; if (_Pitch < -89.0f) _Pitch = -89.0f; }
        updateCameraVectors();
    }
    const float* getPosition() const { return _Eye; }
    const float* getAt() const { return _At; }

private:
    float _Eye[3], _At[3], _WorldUp[3], _Up[3], _Right[3];
    float _Fov = 0.f, _AspectRatio = 0.f, _Near = 0.f, _Far = 0.f;
    float _Yaw = -90.f, _Pitch = 0.f;

    static float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
    static float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
    static void cross3(const float x[3], const float y[3], float o[3]) {
        o[0] = x[1] * y[2] - y[1] * x[2]; o[1] = x[2] * y[0] - y[2] * x[0]; o[2] = x[0] * y[1] - y[0] * x[1];
    }
    static void normalize3(float v[3]) {
        const float inv = 1.f / std::sqrt(dot3(v, v));
        v[0] *= inv; v[1] *= inv; v[2] *= inv;
    }
    void updateCameraVectors() {  // camera.cpp:111-123: unqualified cos/sin on a float are the double functions
        const double yaw = radians(_Yaw), pitch = radians(_Pitch);
        float front[3];
        front[0] = (float)(std::cos(yaw) * std::cos(pitch));
        front[1] = (float)std::sin(pitch);
        front[2] = (float)(std::sin(yaw) * std::cos(pitch));
        normalize3(front);
        for (int k = 0; k < 3; ++k) _At[k] = front[k];
        cross3(_At, _WorldUp, _Right); normalize3(_Right);
        cross3(_Right, _At, _Up); normalize3(_Up);
    }
    // glm::inverse(mat4), func_matrix.inl:347-405; m[c * 4 + r]
    static void inverse4(const float m[16], float out[16]) {
#define RTR_M(c, r) m[(c) * 4 + (r)]
        const float Coef00 = RTR_M(2, 2) * RTR_M(3, 3) - RTR_M(3, 2) * RTR_M(2, 3);
        const float Coef02 = RTR_M(1, 2) * RTR_M(3, 3) - RTR_M(3, 2) * RTR_M(1, 3);
        const float Coef03 = RTR_M(1, 2) * RTR_M(2, 3) - RTR_M(2, 2) * RTR_M(1, 3);
        const float Coef04 = RTR_M(2, 1) * RTR_M(3, 3) - RTR_M(3, 1) * RTR_M(2, 3);
        const float Coef06 = RTR_M(1, 1) * RTR_M(3, 3) - RTR_M(3, 1) * RTR_M(1, 3);
        const float Coef07 = RTR_M(1, 1) * RTR_M(2, 3) - RTR_M(2, 1) * RTR_M(1, 3);
        const float Coef08 = RTR_M(2, 1) * RTR_M(3, 2) - RTR_M(3, 1) * RTR_M(2, 2);
        const float Coef10 = RTR_M(1, 1) * RTR_M(3, 2) - RTR_M(3, 1) * RTR_M(1, 2);
        const float Coef11 = RTR_M(1, 1) * RTR_M(2, 2) - RTR_M(2, 1) * RTR_M(1, 2);
        const float Coef12 = RTR_M(2, 0) * RTR_M(3, 3) - RTR_M(3, 0) * RTR_M(2, 3);
        const float Coef14 = RTR_M(1, 0) * RTR_M(3, 3) - RTR_M(3, 0) * RTR_M(1, 3);
        const float Coef15 = RTR_M(1, 0) * RTR_M(2, 3) - RTR_M(2, 0) * RTR_M(1, 3);
        const float Coef16 = RTR_M(2, 0) * RTR_M(3, 2) - RTR_M(3, 0) * RTR_M(2, 2);
        const float Coef18 = RTR_M(1, 0) * RTR_M(3, 2) - RTR_M(3, 0) * RTR_M(1, 2);
        const float Coef19 = RTR_M(1, 0) * RTR_M(2, 2) - RTR_M(2, 0) * RTR_M(1, 2);
        const float Coef20 = RTR_M(2, 0) * RTR_M(3, 1) - RTR_M(3, 0) * RTR_M(2, 1);
        const float Coef22 = RTR_M(1, 0) * RTR_M(3, 1) - RTR_M(3, 0) * RTR_M(1, 1);
        const float Coef23 = RTR_M(1, 0) * RTR_M(2, 1) - RTR_M(2, 0) * RTR_M(1, 1);
        const float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
        const float Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
        const float Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
        const float Vec0[4] = {RTR_M(1, 0), RTR_M(0, 0), RTR_M(0, 0), RTR_M(0, 0)};
        const float Vec1[4] = {RTR_M(1, 1), RTR_M(0, 1), RTR_M(0, 1), RTR_M(0, 1)};
        const float Vec2[4] = {RTR_M(1, 2), RTR_M(0, 2), RTR_M(0, 2), RTR_M(0, 2)};
        const float Vec3[4] = {RTR_M(1, 3), RTR_M(0, 3), RTR_M(0, 3), RTR_M(0, 3)};
        const float SignA[4] = {+1.f, -1.f, +1.f, -1.f}, SignB[4] = {-1.f, +1.f, -1.f, +1.f};
        float Inverse[16];
        for (int k = 0; k < 4; ++k) {
            const float Inv0 = (Vec1[k] * Fac0[k] - Vec2[k] * Fac1[k]) + Vec3[k] * Fac2[k];
            const float Inv1 = (Vec0[k] * Fac0[k] - Vec2[k] * Fac3[k]) + Vec3[k] * Fac4[k];
            const float Inv2 = (Vec0[k] * Fac1[k] - Vec1[k] * Fac3[k]) + Vec3[k] * Fac5[k];
            const float Inv3 = (Vec0[k] * Fac2[k] - Vec1[k] * Fac4[k]) + Vec2[k] * Fac5[k];
            Inverse[0 * 4 + k] = Inv0 * SignA[k]; Inverse[1 * 4 + k] = Inv1 * SignB[k];
            Inverse[2 * 4 + k] = Inv2 * SignA[k]; Inverse[3 * 4 + k] = Inv3 * SignB[k];
        }
        const float Dot0[4] = {RTR_M(0, 0) * Inverse[0], RTR_M(0, 1) * Inverse[4], RTR_M(0, 2) * Inverse[8], RTR_M(0, 3) * Inverse[12]};
        const float Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
        const float OneOverDeterminant = 1.f / Dot1;
        for (int i = 0; i < 16; ++i) out[i] = Inverse[i] * OneOverDeterminant;
#undef RTR_M
    }
};

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

// cr::MaterialGPU / cr::Material (srcCommon/scene/pbr/material.hpp:9-29, material.cpp): a colour per material
struct MaterialGPU {
    vec4 _Color = {1.f, 1.f, 1.f, 1.f};
};
static_assert(sizeof(MaterialGPU) == sizeof(rtr_material), "MaterialGPU layout");
class Material {
    static uint32_t& idGenerator() { static uint32_t next = 0; return next; }
    uint32_t _Id = 0;

public:
    static const size_t MAX_NB_MATERIALS = 2 << 15;
    MaterialGPU _InternalStruct{};
    explicit Material(const vec4& color) : _Id(idGenerator()++) { _InternalStruct._Color = color; }
    Material() : _Id(idGenerator()++) {  // material.cpp:17-27: three rand() draws
        const float r = static_cast<float>(std::rand()) / static_cast<float>(RAND_MAX);
        const float g = static_cast<float>(std::rand()) / static_cast<float>(RAND_MAX);
        const float b = static_cast<float>(std::rand()) / static_cast<float>(RAND_MAX);
        _InternalStruct._Color = vec4{r, g, b, 1.f};
    }
    uint32_t getId() const { return _Id; }
};

// cr::Triangle (triangle.hpp:16-34, triangle.cpp)
class Triangle {
public:
    static const size_t MAX_NB_TRIANGLES = 2 << 15;  // the reference's SSBO cap; nothing in this library is bound by it
    TriangleGPU _InternalStruct{};

    Triangle() = default;
    explicit Triangle(const TriangleGPU& t) : _InternalStruct(t) {}
    Triangle(const vec4& p0, const vec4& p1, const vec4& p2, uint32_t modelId) {
        std::memset(&_InternalStruct, 0, sizeof(_InternalStruct));
        _InternalStruct._P0 = p0; _InternalStruct._P1 = p1; _InternalStruct._P2 = p2;
        _InternalStruct._ModelId = modelId;
    }
    Triangle(const vec3& p0, const vec3& p1, const vec3& p2, uint32_t modelId)
        : Triangle(vec4{p0.x, p0.y, p0.z, 1.f}, vec4{p1.x, p1.y, p1.z, 1.f}, vec4{p2.x, p2.y, p2.z, 1.f}, modelId) {}

    static vec3 getCentroid(const TriangleGPU& triangle, const mat4& model) {  // triangle.cpp:30-32
        float c[3];
        rtr_triangle_centroid(reinterpret_cast<const rtr_triangle*>(&triangle), reinterpret_cast<const float*>(&model), c);
        return vec3{c[0], c[1], c[2]};
    }
};

// cr::Mesh (mesh.hpp:17-44, mesh.cpp).  Loading and the matrix setters run in librtr_b200 (rtr_obj_load, rtr_mesh_*),
// which follow the reference's tinyobjloader / GLM arithmetic bit for bit; a load error is fatal like the
// reference's (errorHandler.cpp:13-29).
class Mesh;
using MeshPtr = std::shared_ptr<Mesh>;

class Mesh {
    static uint32_t& idGenerator() { static uint32_t next = 0; return next; }  // Mesh::_IdGenerator, mesh.cpp:9
    uint32_t _Id = 0;

public:
    std::vector<Triangle> _Triangles{};
    MeshModelGPU _InternalStruct;
    static const size_t MAX_NB_MESHES = 2 << 5;

    Mesh() { _Id = idGenerator()++; }
    uint32_t getId() const { return _Id; }

    void setModel(const mat4& model) { rtr_mesh_set_model(raw(), reinterpret_cast<const float*>(&model)); }
    void setPosition(const vec3& position) { rtr_mesh_set_position(raw(), position.x, position.y, position.z); }
    void setScale(float scale) { rtr_mesh_set_scale(raw(), scale); }
    void setRotation(float thetaX, float thetaY, float thetaZ) { rtr_mesh_set_rotation(raw(), thetaX, thetaY, thetaZ); }
    void setMaterial(uint32_t materialId) { rtr_mesh_set_material(raw(), materialId); }

    static MeshPtr primitiveTriangle() { return primitive(RTR_PRIMITIVE_TRIANGLE); }
    static MeshPtr primitiveSquare() { return primitive(RTR_PRIMITIVE_SQUARE); }
    static MeshPtr primitiveCube() { return primitive(RTR_PRIMITIVE_CUBE); }
    static MeshPtr primitiveSphere() { return primitive(RTR_PRIMITIVE_SPHERE); }

    static MeshPtr load(const std::string& path) {  // mesh.cpp:186-263
        MeshPtr mesh(new Mesh());
        rtr_triangle* tris = nullptr;
        uint64_t n = 0;
        const int rc = rtr_obj_load(path.c_str(), mesh->_Id, &tris, &n);
        if (rc != RTR_OK) {
            std::fprintf(stderr, "Error triggered in %s:%d\n\tError loading object `%s': %s\nExiting the program!\n", __FILE__,
                         __LINE__, path.c_str(), rtr_last_error(nullptr));
            std::exit(EXIT_FAILURE);
        }
        mesh->adopt(tris, n);
        rtr_obj_free(tris);
        return mesh;
    }

    // the TriangleGPU records of the mesh, back to back (what glr::Scene::getTriangleToGPUData copies, scene.cpp:26-40)
    void appendTo(std::vector<TriangleGPU>& out) const {
        for (const Triangle& t : _Triangles) out.push_back(t._InternalStruct);
    }

private:
    rtr_mesh* raw() { return reinterpret_cast<rtr_mesh*>(&_InternalStruct); }
    void adopt(const rtr_triangle* tris, uint64_t n) {
        _Triangles.resize(static_cast<size_t>(n));
        for (size_t i = 0; i < _Triangles.size(); ++i) std::memcpy(&_Triangles[i]._InternalStruct, tris + i, sizeof(rtr_triangle));
    }
    static MeshPtr primitive(int which) {
        MeshPtr mesh(new Mesh());
        rtr_triangle tris[12];
        uint64_t n = 0;
        if (rtr_mesh_primitive(which, mesh->_Id, tris, 12, &n) == RTR_OK) mesh->adopt(tris, n);
        return mesh;
    }
};

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

// ---------------------------------------------------------------------------------------------------------------
// glr::Scene (srcOpenGL/scene/scene.hpp:22-62, scene.cpp): the container the reference's application fills and hands
// to the shader -- materials (always one default one), meshes, the three *ToGPUData views, the DFS flatten of the BVH
// (getBVH_NodesToGPUData) and sendDataToGpu.  Same members and method names.  What the reference does with GL there
// (createSSBO / bindSSBO / the three uniforms) is the build of the cr::BVH from the views (scene.cpp:148) and keeping the
// arrays for the rays and the shading; `referencePadding` reproduces the fixed-size SSBO views of the reference --
// vectors of MAX_NB_* records, extra input dropped (65 536 triangles, 64 meshes) -- without it the views have the
// size of their content and nothing is dropped.
// ---------------------------------------------------------------------------------------------------------------
namespace glr {
class Scene;
using ScenePtr = std::shared_ptr<Scene>;

class Scene {
private:
    std::vector<cr::Material> _Materials = {cr::Material()};  // always one default material (scene.hpp:24)
    std::vector<cr::MeshPtr> _Meshes = {};
    uint32_t _NbTriangles = 0;
    uint32_t _NbMaterials = 1;  // the default one
    uint32_t _NbMeshes = 0;
    cr::BVH_Ptr _BVH = nullptr;
    rtr_ctx* _ctx = nullptr;
    bool _referencePadding = false;
    // what bindSSBO sent (scene.cpp:112-176), kept for the shading of the hit records
    std::vector<cr::TriangleGPU> _TrianglesSent;
    std::vector<cr::MeshModelGPU> _ModelsSent;
    std::vector<cr::MaterialGPU> _MaterialsSent;

public:
    explicit Scene(rtr_ctx* ctx = nullptr, bool referencePadding = false) : _ctx(ctx), _referencePadding(referencePadding) {}

    std::vector<cr::TriangleGPU> getTriangleToGPUData() const {  // scene.cpp:26-40
        std::vector<cr::TriangleGPU> trianglesGPU;
        if (_referencePadding) {
            cr::TriangleGPU zero;
            std::memset(static_cast<void*>(&zero), 0, sizeof(zero));
            trianglesGPU.assign(cr::Triangle::MAX_NB_TRIANGLES, zero);
        }
        size_t i = 0;
        for (const cr::MeshPtr& mesh : _Meshes)
            for (const cr::Triangle& triangle : mesh->_Triangles) {
                if (_referencePadding) {
                    if (i == cr::Triangle::MAX_NB_TRIANGLES) break;
                    trianglesGPU[i] = triangle._InternalStruct;
                } else {
                    trianglesGPU.push_back(triangle._InternalStruct);
                }
                i++;
            }
        return trianglesGPU;
    }
    std::vector<cr::MaterialGPU> getMaterialToGPUData() const {  // scene.cpp:42-48
        const size_t n = _referencePadding ? std::min(_Materials.size(), cr::Material::MAX_NB_MATERIALS) : _Materials.size();
        std::vector<cr::MaterialGPU> materialGPU(_referencePadding ? cr::Material::MAX_NB_MATERIALS : n);
        for (size_t i = 0; i < n; i++) materialGPU[i] = _Materials[i]._InternalStruct;
        return materialGPU;
    }
    std::vector<cr::MeshModelGPU> getMeshModelToGPUData() const {  // scene.cpp:17-23
        const size_t n = _referencePadding ? std::min(_Meshes.size(), cr::Mesh::MAX_NB_MESHES) : _Meshes.size();
        std::vector<cr::MeshModelGPU> modelsGPU(_referencePadding ? cr::Mesh::MAX_NB_MESHES : n);
        for (size_t i = 0; i < n; i++) modelsGPU[i] = _Meshes[i]->_InternalStruct;
        return modelsGPU;
    }
    // scene.cpp:203-208: the DFS pre-order array the shader walks, root first.  Computed by the library's flatten
    // kernels; flattenTopDown below is the reference's own recursion over _InternalStruct, for comparison.
    std::vector<cr::BVH_NodeGPU> getBVH_NodesToGPUData(cr::BVH_Ptr bvh) const { return bvh->getFlatNodes(); }

    // recursiveTopDownTraversalBVH (scene.cpp:189-201) over the by-cluster-id view, with an explicit stack instead of
    // the recursion (a 10 M-leaf PLOC tree is deep enough to matter): a node is appended when it is first met, its
    // _LeftChild / _RightChild become the positions its children are appended at; leaves keep the 0 / 0 links.
    static std::vector<cr::BVH_NodeGPU> flattenTopDown(const cr::BVH_Params& params, uint32_t nbTriangles) {
        std::vector<cr::BVH_NodeGPU> out;
        if (nbTriangles == 0) return out;
        out.reserve(2 * static_cast<size_t>(nbTriangles) - 1);
        struct Pending { uint32_t node; uint32_t parentPosition; bool isRight; };
        std::vector<Pending> stack;
        stack.push_back({2 * nbTriangles - 2, 0xFFFFFFFFu, false});  // rootId (scene.cpp:205)
        while (!stack.empty()) {
            const Pending cur = stack.back();
            stack.pop_back();
            const uint32_t position = static_cast<uint32_t>(out.size());
            out.push_back(params._Clusters[cur.node].value());
            if (cur.parentPosition != 0xFFFFFFFFu) {
                if (cur.isRight) out[cur.parentPosition]._RightChild = position;
                else out[cur.parentPosition]._LeftChild = position;
            }
            if (!params._IsLeaf[cur.node]) {  // has_value(): set for leaves only (scene.cpp:193)
                stack.push_back({params._RightChild[cur.node].value(), position, true});   // popped second: after the whole
                stack.push_back({params._LeftChild[cur.node].value(), position, false});   // left subtree, as in the recursion
            }
        }
        return out;
    }

    void addMesh(cr::MeshPtr mesh) {  // scene.cpp:50-55
        if (_referencePadding && _Meshes.size() == cr::Mesh::MAX_NB_MESHES) return;
        _Meshes.push_back(mesh);
        _NbMeshes++;
        _NbTriangles += static_cast<uint32_t>(_referencePadding ? std::min(mesh->_Triangles.size(), cr::Triangle::MAX_NB_TRIANGLES)
                                                                : mesh->_Triangles.size());
    }
    void addMaterial(const cr::vec4& color) {  // scene.cpp:57-61
        if (_Materials.size() == cr::Material::MAX_NB_MATERIALS) return;
        _Materials.emplace_back(color);
        _NbMaterials++;
    }
    void addRandomMaterial() {  // scene.cpp:63-67
        if (_Materials.size() == cr::Material::MAX_NB_MATERIALS) return;
        _Materials.emplace_back();
        _NbMaterials++;
    }

    // scene.cpp:210-222 + bindSSBO (:112-176): the views are taken, the BVH is built from them; the three uniforms
    // (uNbTriangles / uNbMaterials / uNbModels) are the getNb* accessors below
    void sendDataToGpu() {
        _MaterialsSent = getMaterialToGPUData();
        _TrianglesSent = getTriangleToGPUData();
        _ModelsSent = getMeshModelToGPUData();
        _BVH = cr::BVH_Ptr(new cr::BVH(_NbTriangles, _TrianglesSent, _ModelsSent, _ctx));
    }

    // one dispatch of raytracer.glsl over the scene that was sent (application.cpp:225-245): closest hits, then
    // getColor (raytracer.glsl:159-179,299-331) -- the rgba32f image, 4 floats per pixel
    std::vector<float> drawOneFrame(const cr::CameraGPU& camera, uint32_t width, uint32_t height, bool wireframe = false) const {
        std::vector<float> rgba(static_cast<size_t>(width) * height * 4, 0.f);
        if (!_BVH) return rgba;
        const std::vector<cr::Hit> hits = _BVH->tracePrimary(camera, width, height);
        RTR_SCENE_CHECK(_BVH->context(),
                        rtr_shade(_BVH->context(), reinterpret_cast<const rtr_hit*>(hits.data()), hits.size(),
                                  reinterpret_cast<const rtr_triangle*>(_TrianglesSent.data()), _NbTriangles,
                                  reinterpret_cast<const rtr_mesh*>(_ModelsSent.data()), _NbMeshes,
                                  reinterpret_cast<const rtr_material*>(_MaterialsSent.data()), _NbMaterials,
                                  wireframe ? RTR_SHADE_WIREFRAME : 0u, nullptr, rgba.data()));
        return rgba;
    }

    cr::BVH_Ptr getBVH() const { return _BVH; }
    uint32_t getNbTriangles() const { return _NbTriangles; }
    uint32_t getNbMaterials() const { return _NbMaterials; }
    uint32_t getNbMeshes() const { return _NbMeshes; }
};
}  // namespace glr

// ---------------------------------------------------------------------------------------------------------------
// glr::ApplicationFPS (srcOpenGL/application.hpp:15-34, application.cpp:345-373): the frame-time statistics of the
// reference's main loop (render() -> _FPS.increment(), then the "Statistics" window).  Same members, same float
// arithmetic -- increment() accumulates the time since the last frame, every _NbFramesBetweenDisplay frames display()
// turns the window's sum / longest / shortest frame into avg / min / max FPS and starts a new window -- with the two
// bindings to the outside made explicit: the clock is a parameter (the reference reads glfwGetTime(): seconds since
// start as a double, narrowed to float; increment() without argument reads a steady clock the same way) and display()
// prints the three lines of the ImGui text to a FILE* (nullptr: statistics only).
// ---------------------------------------------------------------------------------------------------------------
#include <chrono>
#include <cmath>
#include <cstdio>
namespace glr {
struct ApplicationFPS {
public:
    bool _DisplayFPS = true;
    uint32_t _NbFramesBetweenDisplay = 10;
    float _LastFrame = 0.f;

private:
    uint32_t _FrameCounter = 0;
    float _SumOfTimes = 0.f;
    float _MinTime = INFINITY;
    float _MaxTime = 0.f;
    float _AvgFPS = 0.f;
    float _MinFPS = 0.f;
    float _MaxFPS = 0.f;

public:
    void increment(double timeSeconds) {  // application.cpp:345-355, glfwGetTime() passed in
        const float currentFrame = static_cast<float>(timeSeconds);
        const float deltaTime = currentFrame - _LastFrame;
        _LastFrame = currentFrame;
        _SumOfTimes += deltaTime;
        _FrameCounter++;
        if (deltaTime > _MaxTime) _MaxTime = deltaTime;
        if (deltaTime < _MinTime) _MinTime = deltaTime;
    }
    void increment() {
        static const std::chrono::steady_clock::time_point start = std::chrono::steady_clock::now();
        increment(std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count());
    }
    void display(FILE* out = stdout) {  // application.cpp:357-373; returns after updating when out == nullptr
        const bool fresh = _FrameCounter >= _NbFramesBetweenDisplay;
        if (fresh) {
            _AvgFPS = 1.f / (_SumOfTimes / static_cast<float>(_FrameCounter));
            _MinFPS = 1.f / _MaxTime;
            _MaxFPS = 1.f / _MinTime;
            _FrameCounter = 0;
            _SumOfTimes = 0.f;
            _MinTime = INFINITY;
            _MaxTime = 0.f;
        }
        if (out && _DisplayFPS && fresh)
            std::fprintf(out, "avg FPS: %.2f\nmin FPS: %.2f\nmax FPS: %.2f\n\n", _AvgFPS, _MinFPS, _MaxFPS);
    }
    float avgFPS() const { return _AvgFPS; }
    float minFPS() const { return _MinFPS; }
    float maxFPS() const { return _MaxFPS; }
};
}  // namespace glr

#endif  // RTR_SCENE_HPP
