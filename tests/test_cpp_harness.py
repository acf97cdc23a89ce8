"""The C++ side of the boundary: the re-hosted tests/testsSortGPU executables and the cr::BVH shim
(include/rtr_scene.hpp) compile with plain g++ against the C ABI (CPU check), and pass on a B200 (GPU check)."""
import os
import subprocess

import pytest

from realtimeraytracing_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "testsSortGPU")
TESTS = ["testHistogramCreation", "testHistogramPrefixSum", "testRadixSort", "testBvhShim"]


@pytest.fixture(scope="module")
def executables():
    build.build()
    env = dict(os.environ)
    env.pop("CXX", None)
    out = subprocess.run(["make", "-C", HARNESS, "CXX=/usr/bin/g++"], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    return {t: os.path.join(HARNESS, "build", t) for t in TESTS + ["testCamera", "testMesh", "testFps", "testScene"]}


def test_camera_mirror_is_bit_identical_to_the_reference(executables):
    """cr::Camera of include/rtr_scene.hpp vs the reference's own camera.cpp (oracle/_ref/libref_camera.so, built from
    /root/reference by oracle/Makefile): 2001 cameras incl. replayed mouse / keyboard input, CameraGPU byte for byte."""
    from oracle import bindings as ob
    try:
        ob.build_reference()
    except Exception:
        pass  # no reference tree here: the prebuilt library travels with the repo snapshot
    lib = os.path.join(ROOT, "oracle", "_ref", "libref_camera.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref/libref_camera.so not built (needs /root/reference)")
    r = subprocess.run([executables["testCamera"], lib], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout


def test_mesh_mirror_is_bit_identical_to_the_reference(executables):
    """cr::Mesh / cr::Triangle of include/rtr_scene.hpp vs the reference's own mesh.cpp (oracle/_ref/libref_mesh.so):
    Mesh::load on the OBJ fixtures, primitives, 2000 replayed setter sequences and centroids."""
    import glob
    lib = os.path.join(ROOT, "oracle", "_ref", "libref_mesh.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref/libref_mesh.so not built (needs /root/reference)")
    objs = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "obj", "*.obj")))
    r = subprocess.run([executables["testMesh"], lib] + objs, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout and "%d cases" % (len(objs) + 4 + 2000) in r.stdout


def test_fps_statistics_mirror(executables):
    """glr::ApplicationFPS (application.cpp:345-373) in include/rtr_scene.hpp: 3000 replayed frames, three window
    lengths, against the restated float arithmetic; a known-answer window and its printed text."""
    r = subprocess.run([executables["testFps"]], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout


def test_scene_mirror_container_and_flatten(executables):
    """glr::Scene of include/rtr_scene.hpp (scene.hpp:22-62, scene.cpp): counters, caps and the fixed-size views of the
    reference, and its recursiveTopDownTraversalBVH restated over `_InternalStruct` against the oracle's flatten on PLOC
    trees of 1 ... 20 000 triangles (the oracle's flatten is pinned to the reference's fixtures)."""
    from oracle import bindings as ob
    ob.Oracle()     # builds oracle/librtr_oracle.so if it is not there
    lib = os.path.join(ROOT, "oracle", "librtr_oracle.so")
    assert os.path.exists(lib)
    r = subprocess.run([executables["testScene"], lib], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "9 cases, 0 mismatches" in r.stdout


def test_shim_compiles_with_glm_types(tmp_path):
    """RTR_SCENE_USE_GLM: cr::Mesh / cr::Triangle / cr::Material / cr::Camera with glm::vec / glm::mat members."""
    glm = "/root/reference/srcVulkan/dep/slang/external/glm"
    if not os.path.isdir(glm):
        pytest.skip("no GLM tree here")
    build.build()
    exe = str(tmp_path / "glmMode")
    libdir = os.path.join(ROOT, "realtimeraytracing_b200", "lib")
    out = subprocess.run(["/usr/bin/g++", "-std=c++20", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                          "-I" + glm, os.path.join(HARNESS, "glmMode.cpp"), "-o", exe, "-L" + libdir, "-lrtr_b200",
                          "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    assert subprocess.run([exe]).returncode == 0


def test_cpp_harness_compiles_and_links(executables):
    for t, path in executables.items():
        assert os.path.exists(path), t
        if t in ("testCamera", "testFps"):
            continue  # host-only: cr::Camera needs nothing from the library
        needed = subprocess.check_output(["readelf", "-d", path], text=True)
        assert "librtr_b200.so" in needed, t  # bound to the C ABI library, nothing else of ours


def test_cpp_harness_fails_loudly_without_a_device(executables):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    r = subprocess.run([executables["testHistogramCreation"]], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", TESTS)
def test_cpp_harness_passes_on_gpu(executables, name):
    r = subprocess.run([executables[name]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


EXAMPLES = os.path.join(ROOT, "examples")


@pytest.fixture(scope="module")
def render_obj():
    build.build()
    env = dict(os.environ)
    env.pop("CXX", None)
    out = subprocess.run(["make", "-C", EXAMPLES, "CXX=/usr/bin/g++"], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    return os.path.join(EXAMPLES, "build", "render_obj")


def test_render_obj_example_compiles_and_needs_a_device(render_obj):
    assert "librtr_b200.so" in subprocess.check_output(["readelf", "-d", render_obj], text=True)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    r = subprocess.run([render_obj, os.path.join(ROOT, "tests", "golden", "obj", "polygons.obj"), "/dev/null"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--wireframe"], ["--bvh-depth", "3"]])
def test_render_obj_example_equals_the_python_pipeline(render_obj, ctx, tmp_path, extra):
    """examples/render_obj.cpp (cr::Mesh::load -> cr::BVH -> tracePrimary -> rtr_shade through the C++ shim) produces
    the float frame the Python mirror produces from the same OBJ file, transform and camera, bit for bit."""
    import numpy as np
    from realtimeraytracing_b200 import capi, scene as rscene
    from realtimeraytracing_b200.layouts import CAMERA
    obj = os.path.join(ROOT, "tests", "golden", "obj", "random.obj")
    W, H = 320, 192
    ppm, cam_file, rgba_file = (str(tmp_path / n) for n in ("out.ppm", "cam.bin", "rgba.bin"))
    r = subprocess.run([render_obj, obj, ppm, "--size", str(W), str(H), "--scale", "2e-5", "--rotate", "0.3", "-0.2", "0.1",
                        "--dump-camera", cam_file, "--dump-rgba", rgba_file] + extra, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(rgba_file, dtype=np.float32).reshape(-1, 4)
    cam = np.fromfile(cam_file, dtype=CAMERA)
    assert cam.size == 1 and got.shape[0] == W * H

    mesh = rscene.Mesh.load(obj)
    mesh.setMaterial(0)
    mesh.setRotation(0.3, -0.2, 0.1)
    mesh.setScale(2e-5)
    tris = mesh._Triangles.copy()
    tris["model_id"] = 0
    bvh = capi.Bvh(ctx).build(tris, mesh._InternalStruct)
    try:
        hits = bvh.trace_primary(cam, W, H)
        overlay = bvh.depth_overlay(cam, W, H, 3) if "--bvh-depth" in extra else None
        exp = ctx.shade(hits, tris, mesh._InternalStruct, [[0.2, 0.3, 0.1, 1.0]], wireframe="--wireframe" in extra, bvh_rgba=overlay)
    finally:
        bvh.close()
    assert hits["did_hit"].sum() > 500
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    with open(ppm, "rb") as f:
        data = f.read()
    header = b"P6\n%d %d\n255\n" % (W, H)
    assert data.startswith(header) and len(data) == len(header) + 3 * W * H
    px = np.frombuffer(data[len(header):], dtype=np.uint8).reshape(-1, 3)
    assert np.array_equal(px, np.floor(np.clip(exp[:, :3], 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("rebuild", [False, True])
def test_render_obj_frame_loop_reports_fps(render_obj, tmp_path, rebuild):
    """--frames: the reference's mainLoop (application.cpp:220-313) over the shim -- 25 frames give two statistics
    windows of 10, and the final frame is the one the single-shot run writes."""
    obj = os.path.join(ROOT, "tests", "golden", "obj", "random.obj")
    base = [render_obj, obj, None, "--size", "320", "192", "--scale", "2e-5"]
    outs = []
    for extra in ([], ["--frames", "25"] + (["--rebuild"] if rebuild else [])):
        ppm = str(tmp_path / ("f%d.ppm" % len(outs)))
        cmd = list(base) + extra
        cmd[2] = ppm
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append((r.stdout, open(ppm, "rb").read()))
    assert outs[0][0].count("avg FPS") == 0 and outs[1][0].count("avg FPS") == 2
    fps = [float(l.split(":")[1]) for l in outs[1][0].splitlines() if l.startswith(("avg", "min", "max"))]
    assert all(f > 0 for f in fps) and fps[1] <= fps[0] <= fps[2]
    assert outs[0][1] == outs[1][1]
