"""Builds librtr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m realtimeraytracing_b200.build [--force] [--verbose]

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX for other targets
  -fmad=false                               no FMA contraction: fp32 results bit-equal to the
                                            reference's x86-64 build
  -lineinfo                                 ncu source pages map to these files
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librtr_b200.so")
SOURCES = ["api.cu", "morton.cu", "sort.cu", "ploc.cu", "trace.cu", "comm.cu", "mesh.cu"]
HEADERS = ["common.cuh", "bvh.cuh", os.path.join("..", "..", "include", "rtr.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=default,-ffp-contract=off",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a compiler wrapper without OpenMP specs; nvcc must use the system g++
    extra = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    procs = []
    objs = []
    tuning = os.environ.get("RTR_NVCC_EXTRA", "").split()  # e.g. "-DRTR_LEAF_BATCH=8" for parameter sweeps
    only = os.environ.get("RTR_BUILD_ONLY", "").split()    # parameter sweeps: recompile just these sources (objects of the others are reused)
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        if only and src not in only and os.path.exists(obj):
            continue
        cmd = [nvcc] + NVCC_FLAGS + tuning + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write("[%s]\n%s\n" % (src, out))
    if failed:
        raise RuntimeError("librtr_b200 build failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + extra + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    subprocess.check_call(link, env=env)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
