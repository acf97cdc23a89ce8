"""Deterministic synthetic inputs (SURVEY.md §8d "Synthetic generator").

Everything is derived from raw MT19937 32-bit outputs exactly the way libstdc++'s
``std::mt19937`` + ``std::uniform_real_distribution<float>`` /
``std::uniform_int_distribution<uint32_t>`` derive them, so the scenes equal the ones a
C++ host would generate with the recipe in SURVEY.md (and App. C's known-answer scene
reproduces the reference hashes recorded there).
"""
import math

import numpy as np

from .layouts import CAMERA, MESH, TRIANGLE


def mt19937_raw(seed: int, count: int) -> np.ndarray:
    """`count` successive outputs of std::mt19937(seed) as uint32."""
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed) & 0xFFFFFFFF)
    return bg.random_raw(count).astype(np.uint32)


def canonical_f32(raw_u32: np.ndarray) -> np.ndarray:
    """std::generate_canonical<float, 24>(mt19937): float(u) / 2^32, clamped below 1."""
    r = raw_u32.astype(np.float32) * np.float32(2.0 ** -32)
    one = np.float32(1.0)
    return np.where(r >= one, np.nextafter(one, np.float32(0.0)), r).astype(np.float32)


def uniform_f32(raw_u32: np.ndarray, a: float, b: float) -> np.ndarray:
    """std::uniform_real_distribution<float>(a, b): canonical * (b - a) + a, in fp32."""
    a32, b32 = np.float32(a), np.float32(b)
    return (canonical_f32(raw_u32) * np.float32(b32 - a32) + a32).astype(np.float32)


def identity_meshes(count: int = 1) -> np.ndarray:
    meshes = np.zeros(count, dtype=MESH)
    meshes["m"][:] = np.eye(4, dtype=np.float32).reshape(16)
    return meshes


def random_keys_u32(n: int, seed: int = 1, lo: int = 0, hi: int = 0xFFFFFFFF) -> np.ndarray:
    """Config-1 keys: mt19937(seed), uniform_int_distribution<uint32_t>(lo, hi).

    Full range uses the raw outputs (libstdc++ takes that shortcut); a narrower range is
    drawn with numpy's bounded generator instead (only the distribution matters there:
    the reference harness seeds from std::random_device, testHistogramCreation.cpp:19-20).
    """
    if lo == 0 and hi == 0xFFFFFFFF:
        return mt19937_raw(seed, n)
    rng = np.random.RandomState(seed)
    return rng.randint(lo, hi + 1, size=n, dtype=np.int64).astype(np.uint32)


def soup_extent(n: int, edge: float = 0.25) -> float:
    """Cube side L = 2*e*N^(1/3)."""
    return 2.0 * edge * (n ** (1.0 / 3.0))


def triangle_soup(n: int, seed=None, edge: float = 0.25, extent=None):
    """Uniform triangle soup: returns (triangles[n] TRIANGLE, meshes[1] MESH, L).

    mt19937(seed = n); per triangle 12 sequenced draws: centre c ~ U(-L/2, L/2)^3, then
    three vertices c + e * U(-1, 1)^3; w = 1; ModelId = 0; one identity mesh.
    """
    seed = n if seed is None else seed
    L = soup_extent(n, edge) if extent is None else float(extent)
    tris = np.zeros(n, dtype=TRIANGLE)
    chunk = 1 << 20
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed) & 0xFFFFFFFF)
    e32 = np.float32(edge)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        raw = bg.random_raw(12 * m).astype(np.uint32).reshape(m, 12)
        c = uniform_f32(raw[:, 0:3], -L / 2.0, L / 2.0)
        for k, name in enumerate(("p0", "p1", "p2")):
            off = uniform_f32(raw[:, 3 + 3 * k: 6 + 3 * k], -1.0, 1.0)
            tris[name][s:s + m, 0:3] = c + e32 * off
            tris[name][s:s + m, 3] = 1.0
    return tris, identity_meshes(1), L


def survey_known_answer_scene(n: int = 20000, seed: int = 7):
    """SURVEY.md App. C wiring scene: anisotropic slab of small triangles.

    r = v3(); c = (r.x*0.3f, r.y*2.f, r.z); o0, o1, o2 = v3(); P_k = (c + 0.05f*o_k, 1).
    Reference results: 2n-1 = 39999 flat nodes, root.left = 1, root.right = 20100,
    flat-node hash 6c2aabd40444c4ec, triangle-index hash cc1bb62a0ae92f47.
    """
    raw = mt19937_raw(seed, 12 * n).reshape(n, 12)
    u = uniform_f32(raw, -1.0, 1.0)
    c = np.empty((n, 3), dtype=np.float32)
    c[:, 0] = u[:, 0] * np.float32(0.3)
    c[:, 1] = u[:, 1] * np.float32(2.0)
    c[:, 2] = u[:, 2]
    tris = np.zeros(n, dtype=TRIANGLE)
    for k, name in enumerate(("p0", "p1", "p2")):
        tris[name][:, 0:3] = c + np.float32(0.05) * u[:, 3 + 3 * k: 6 + 3 * k]
        tris[name][:, 3] = 1.0
    return tris, identity_meshes(1)


def grid_mesh(nx: int, ny: int, size: float = 4.0, z_amp: float = 0.5):
    """A connected height-field mesh (shared edges/vertices => exact-tie stress case).

    2*nx*ny counter-clockwise triangles (front faces towards -z, i.e. towards the
    reference camera at z < 0).
    """
    xs = np.linspace(-size / 2, size / 2, nx + 1, dtype=np.float32)
    ys = np.linspace(-size / 2, size / 2, ny + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    Z = (np.float32(z_amp) * np.sin(X * np.float32(1.7)) * np.cos(Y * np.float32(1.3))).astype(np.float32)
    P = np.stack([X, Y, Z, np.ones_like(X)], axis=-1).astype(np.float32)
    a = P[:-1, :-1].reshape(-1, 4)
    b = P[:-1, 1:].reshape(-1, 4)
    c = P[1:, :-1].reshape(-1, 4)
    d = P[1:, 1:].reshape(-1, 4)
    tris = np.zeros(2 * nx * ny, dtype=TRIANGLE)
    # winding chosen so normalize(cross(P1-P0, P2-P0)) points to -z
    tris["p0"][0::2], tris["p1"][0::2], tris["p2"][0::2] = a, c, b
    tris["p0"][1::2], tris["p1"][1::2], tris["p2"][1::2] = b, c, d
    return tris, identity_meshes(1)


def _look_at(eye, center, up):
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, 0:3], m[1, 0:3], m[2, 0:3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s.dot(eye), -u.dot(eye), f.dot(eye)
    return m


def reference_camera(eye=(0.0, 0.0, -5.0), aspect: float = 1280.0 / 720.0, fov_deg: float = 45.0,
                     near: float = 0.1, far: float = 200.0, yaw_deg: float = -90.0,
                     pitch_deg: float = 0.0) -> np.ndarray:
    """cr::Camera::getGpuData (srcCommon/scene/camera.cpp:23-66,126-137) in float64,
    rounded once to fp32.  Producing CameraGPU is host work outside the accelerated
    path; the kernels only consume the struct."""
    eye = np.asarray(eye, dtype=np.float64)
    yaw, pitch = math.radians(yaw_deg), math.radians(pitch_deg)
    front = np.array([math.cos(yaw) * math.cos(pitch), math.sin(pitch), math.sin(yaw) * math.cos(pitch)])
    at = front / np.linalg.norm(front)
    world_up = np.array([0.0, 1.0, 0.0])
    right = np.cross(at, world_up)
    right /= np.linalg.norm(right)
    up = np.cross(right, at)
    up /= np.linalg.norm(up)
    view = _look_at(eye, eye + at, up)
    t = math.tan(math.radians(fov_deg) / 2.0)
    proj = np.zeros((4, 4))
    proj[0, 0] = 1.0 / (aspect * t)
    proj[1, 1] = 1.0 / t
    proj[2, 2] = -(far + near) / (far - near)
    proj[3, 2] = -1.0
    proj[2, 3] = -(2.0 * far * near) / (far - near)
    cam = np.zeros(1, dtype=CAMERA)
    # column-major storage: element [c*4 + r] = M[r, c]
    cam["view"][0] = view.T.reshape(16).astype(np.float32)
    cam["proj"][0] = proj.T.reshape(16).astype(np.float32)
    cam["inv_view"][0] = np.linalg.inv(view).T.reshape(16).astype(np.float32)
    cam["inv_proj"][0] = np.linalg.inv(proj).T.reshape(16).astype(np.float32)
    cam["eye"][0] = np.array([eye[0], eye[1], eye[2], 1.0], dtype=np.float32)
    plane_h = np.float32(2.0) * np.float32(near) * np.float32(math.tan(0.5 * math.radians(fov_deg)))
    cam["plane_height"][0] = plane_h
    cam["plane_width"][0] = np.float32(plane_h) * np.float32(aspect)
    cam["plane_near"][0] = np.float32(near)
    return cam


def soup_camera(L: float, width: int, height: int) -> np.ndarray:
    """Reference camera pulled back to (0, 0, -1.5 L) so the soup cube overfills the fov."""
    return reference_camera(eye=(0.0, 0.0, -1.5 * L), aspect=float(width) / float(height),
                            far=max(200.0, 4.0 * L))


def reference_denominators(width: int, height: int):
    """Q5: the reference traces floor(W/16)*16 x floor(H/16)*16 pixels and divides by that."""
    return (width // 16) * 16, (height // 16) * 16
