#!/usr/bin/env python
"""Per-stage device times of the rebuild (events inside the library), for parameter sweeps:
RTR_NVCC_EXTRA="-DRTR_PLOC_WARPS=8" python profiles/time_build.py --tris 10000000 --force-build"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    p = argparse.ArgumentParser()
    p.add_argument("--tris", type=int, default=10_000_000)
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--force-build", action="store_true")
    args = p.parse_args()
    import numpy as np
    from realtimeraytracing_b200 import build as rbuild, capi, synth
    rbuild.build(force=args.force_build)
    with capi.Context(0) as ctx:
        n = args.tris
        tris, meshes, L = synth.triangle_soup(n)
        d_tris = ctx.dev_alloc(tris.nbytes); d_meshes = ctx.dev_alloc(meshes.nbytes)
        ctx.upload(d_tris, tris); ctx.upload(d_meshes, meshes)
        bvh = capi.Bvh(ctx)
        for _ in range(2):
            bvh.build_dev(d_tris, n, n, d_meshes, 1)
        bvh.enable_stage_timing(True)
        st = np.zeros(6)
        for _ in range(args.reps):
            bvh.build_dev(d_tris, n, n, d_meshes, 1)
            st += bvh.stage_ms()
        st /= args.reps
        bvh.enable_stage_timing(False)
        ctx.profile_enable(True)
        bvh.build_dev(d_tris, n, n, d_meshes, 1)
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        print("RTR_NVCC_EXTRA=%r  morton %.3f sort %.3f leaf %.3f ploc %.3f flatten %.3f total %.3f ms" % (
            (os.environ.get("RTR_NVCC_EXTRA", ""),) + tuple(st)))
        print("   " + "  ".join("%s %.3f/%d" % (k, v[0], v[1]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])))
        bvh.close()

if __name__ == "__main__":
    main()
