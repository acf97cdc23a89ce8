"""The obligation the default traversal rests on, checked directly on the CPU: tests/host/trav_records_check.cu compiles
the encoder and the decoder of the traversal records (bvh.cuh -- the functions the kernels inline, __host__ __device__)
for the host and asserts, on 2.4 M random box tests, that a child box the reference's slab test accepts is accepted by
the compressed test with an entry distance that is not larger.  A deliberately shrunk encoding must be caught."""
import os
import shutil
import subprocess

import pytest

from realtimeraytracing_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "trav_records_check.cu")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    nvcc = build._nvcc()
    if not (os.path.exists(nvcc) or shutil.which(nvcc)):
        pytest.skip("nvcc not found")
    exe = str(tmp_path_factory.mktemp("trav") / "trav_records_check")
    cmd = [nvcc, "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-ffp-contract=off,-frounding-math",
           "-I" + os.path.join(ROOT, "include"), SRC, "-o", exe]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)
    out = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    return exe


def test_compressed_tests_never_reject_what_the_reference_accepts(checker):
    r = subprocess.run([checker, "400000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 violations" in r.stdout and "2400000 box tests" in r.stdout


def test_the_check_notices_a_shrunk_encoding(checker):
    r = subprocess.run([checker, "50000", "--shrink"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and " 0 violations" not in r.stdout


def test_extreme_magnitudes_are_flagged_not_misencoded(checker):
    """Scene scales from 1e-25 to 1e14: nodes the grid cannot represent carry the 'no culling here' flag (the check
    counts them as skipped) and every other node still satisfies the obligation."""
    r = subprocess.run([checker, "200000", "--wide"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 violations" in r.stdout and " 0 unencodable" not in r.stdout
