#!/usr/bin/env python
"""Minimal driver for ncu captures: `frames` full rebuilds + frames of the bench workload, nothing else.

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> \
        -o gpurun_out/<name> python profiles/prof_driver.py --tris 10000000 --frames 2

Numbers printed under a profiler are never bench values; bench.py is the only source of those.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--tris", type=int, default=10_000_000)
    p.add_argument("--width", type=int, default=3840)
    p.add_argument("--height", type=int, default=2160)
    p.add_argument("--bounces", type=int, default=2)
    p.add_argument("--frames", type=int, default=2)
    p.add_argument("--no-render", action="store_true")
    p.add_argument("--sort-only", type=int, default=0, help="sort this many random (key,index) pairs instead")
    args = p.parse_args()

    import numpy as np
    from realtimeraytracing_b200 import build as rbuild, capi, synth
    rbuild.build()
    with capi.Context(0) as ctx:
        if args.sort_only:
            keys = synth.random_keys_u32(args.sort_only, seed=1)
            d_keys = ctx.dev_alloc(keys.nbytes)
            d_vals = ctx.dev_alloc(keys.nbytes)
            for _ in range(args.frames):
                ctx.upload(d_keys, keys)
                ctx.upload(d_vals, np.arange(keys.size, dtype=np.uint32))
                ctx.sort_pairs_u32_dev(d_keys, d_vals, keys.size)
                ctx.sync()
            ctx.dev_free(d_keys); ctx.dev_free(d_vals)
            return
        n, W, H = args.tris, args.width, args.height
        tris, meshes, L = synth.triangle_soup(n)
        cam = synth.soup_camera(L, W, H)
        d_tris = ctx.dev_alloc(tris.nbytes)
        d_meshes = ctx.dev_alloc(meshes.nbytes)
        d_rgba = ctx.dev_alloc(W * H * 16)
        ctx.upload(d_tris, tris)
        ctx.upload(d_meshes, meshes)
        bvh = capi.Bvh(ctx)
        for _ in range(args.frames):
            bvh.build_dev(d_tris, n, n, d_meshes, 1)
            if not args.no_render:
                bvh.render_sharded_dev(cam, W, H, d_rgba, 16, 0, 1, bounces=args.bounces)
            ctx.sync()
        bvh.close()
        for p_ in (d_tris, d_meshes, d_rgba):
            ctx.dev_free(p_)


if __name__ == "__main__":
    main()
