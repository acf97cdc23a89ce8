#!/usr/bin/env python
"""Times the traversal alone (BVH built once) with CUDA events inside the library's stream; used for
parameter sweeps:  RTR_NVCC_EXTRA="-DRTR_LEAF_BATCH=8" python profiles/time_render.py --tris 10000000"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--tris", type=int, default=10_000_000)
    p.add_argument("--width", type=int, default=3840)
    p.add_argument("--height", type=int, default=2160)
    p.add_argument("--bounces", type=int, default=2)
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--force-build", action="store_true")
    args = p.parse_args()
    import numpy as np
    from realtimeraytracing_b200 import build as rbuild, capi, synth
    rbuild.build(force=args.force_build)
    with capi.Context(0) as ctx:
        n, W, H = args.tris, args.width, args.height
        tris, meshes, L = synth.triangle_soup(n)
        cam = synth.soup_camera(L, W, H)
        d_tris = ctx.dev_alloc(tris.nbytes); d_meshes = ctx.dev_alloc(meshes.nbytes)
        d_rgba = ctx.dev_alloc(W * H * 16); d_rays = ctx.dev_alloc(8)
        ctx.upload(d_tris, tris); ctx.upload(d_meshes, meshes)
        bvh = capi.Bvh(ctx).build_dev(d_tris, n, n, d_meshes, 1)
        ctx.zero(d_rays, 8)
        for _ in range(2):
            bvh.render_sharded_dev(cam, W, H, d_rgba, 16, 0, 1, bounces=args.bounces)
        ctx.sync()
        ctx.profile_enable(True)
        ctx.zero(d_rays, 8)
        for _ in range(args.reps):
            bvh.render_sharded_dev(cam, W, H, d_rgba, 16, 0, 1, rays_dev=d_rays, bounces=args.bounces)
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        rays = np.zeros(1, np.uint64); ctx.download(rays, d_rays)
        ms, cnt = prof["render_kernel"]
        print("RTR_NVCC_EXTRA=%r render %.3f ms/frame  %.1f Mrays/s  (%d rays/frame)" % (
            os.environ.get("RTR_NVCC_EXTRA", ""), ms / cnt, rays[0] / cnt / (ms / cnt) / 1e3, rays[0] // cnt))
        bvh.close()


if __name__ == "__main__":
    main()
