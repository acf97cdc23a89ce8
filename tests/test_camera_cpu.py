"""scene.Camera / rtr_camera_gpu_data (the C++ shim's cr::Camera compiled into librtr_b200.so) against the reference's
own camera.cpp: CameraGPU (284 bytes) bit for bit.  Two checkers: tests/golden/camera_golden.npz (outputs of
oracle/_ref/libref_camera.so, written by `python tests/test_camera_cpu.py --regenerate`) and, when it is present,
that library itself on 600 random cameras with replayed mouse / keyboard input.  Host only."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from realtimeraytracing_b200 import build, capi, scene  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "libref_camera.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "camera_golden.npz")


def cases(seed, count):
    rng = np.random.default_rng(seed)
    out = [((0.0, 0.0, -5.0), 1280.0 / 720.0, 45.0, 0.1, 200.0, [])]  # application.cpp:16-20
    for c in range(count):
        eye = tuple(float(np.float32(v)) for v in rng.uniform(-300, 300, 3))
        aspect, fov = float(np.float32(rng.uniform(0.5, 2.5))), float(np.float32(rng.uniform(20, 100)))
        near, far = float(np.float32(rng.uniform(0.05, 0.5))), float(np.float32(rng.uniform(100, 400)))
        events = []
        for _ in range(c % 7):
            k = int(rng.integers(0, 7))
            if k == 0:
                events.append((0, float(np.float32(rng.uniform(-400, 400))), float(np.float32(rng.uniform(-400, 400)))))
            else:
                events.append((k, float(np.float32(rng.uniform(0.001, 0.05))), float(rng.integers(0, 2))))
        out.append((eye, aspect, fov, near, far, events))
    return out


def ours(case):
    eye, aspect, fov, near, far, events = case
    cam = scene.Camera(eye, aspect, fov, near, far)
    for kind, a, b in events:
        if kind == 0:
            cam.ProcessMouseMovement(a, b)
        else:
            cam._Accelerate = b != 0.0
            cam.processKeyboard(kind - 1, a)
    return cam.getGpuData().view(np.uint32).reshape(71).copy()


def reference(lib, case):
    eye, aspect, fov, near, far, events = case
    e = np.array(eye, dtype=np.float32)
    kind = np.array([ev[0] for ev in events], dtype=np.int32)
    a = np.array([ev[1] for ev in events], dtype=np.float32)
    b = np.array([ev[2] for ev in events], dtype=np.float32)
    out = np.zeros(71, dtype=np.uint32)
    lib.ref_camera_gpu_data(e.ctypes.data_as(C.c_void_p), C.c_float(aspect), C.c_float(fov), C.c_float(near), C.c_float(far),
                            len(events), kind.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                            b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def ref_lib():
    if not os.path.exists(REF):
        return None
    lib = C.CDLL(REF)
    lib.ref_camera_gpu_data.restype = None
    lib.ref_camera_gpu_data.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 4
    return lib


@pytest.fixture(scope="module", autouse=True)
def _lib():
    build.build()


def test_python_camera_equals_reference_golden():
    want = np.load(GOLDEN)["camera_gpu"]
    got = np.stack([ours(c) for c in cases(11, 60)])
    assert got.shape == want.shape and np.array_equal(got, want)


def test_python_camera_equals_reference_library():
    lib = ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libref_camera.so not built (needs /root/reference)")
    for i, case in enumerate(cases(23, 600)):
        assert np.array_equal(ours(case), reference(lib, case)), (i, case)
    assert np.array_equal(np.stack([reference(lib, c) for c in cases(11, 60)]), np.load(GOLDEN)["camera_gpu"])


def test_bad_event_is_refused():
    with pytest.raises(capi.RtrError):
        capi.camera_gpu_data((0, 0, -5), 1.5, events=[(9, 0.1, 0.0)])


if __name__ == "__main__" and "--regenerate" in sys.argv:
    lib = ref_lib()
    assert lib is not None, "run `make -C oracle ref` where /root/reference exists"
    np.savez_compressed(GOLDEN, camera_gpu=np.stack([reference(lib, c) for c in cases(11, 60)]))
    print("wrote", GOLDEN)
