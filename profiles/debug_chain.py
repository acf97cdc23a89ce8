import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from realtimeraytracing_b200 import capi
from realtimeraytracing_b200.layouts import RAY
import test_trace_gpu as T
ctx = capi.Context(0)
for depth in (100, 140, 300, 600):
    tris, meshes, flat = T._chain_scene(depth)
    rays = np.zeros(64, dtype=RAY)
    rays["o"][:, :3] = (0.05, -0.1, -10.0); rays["o"][:, 3] = 1.0
    rays["d"][:, :3] = (0.001, 0.002, 1.0)
    rays["d"][:, :3] /= np.linalg.norm(rays["d"][0, :3])
    d_nodes, d_tris, d_meshes = ctx.dev_alloc(flat.nbytes), ctx.dev_alloc(tris.nbytes), ctx.dev_alloc(meshes.nbytes)
    ctx.upload(d_nodes, flat); ctx.upload(d_tris, tris); ctx.upload(d_meshes, meshes)
    bvh = capi.Bvh(ctx).adopt_dev(d_nodes, tris.size, d_tris, d_meshes, meshes.size)
    d_rays, d_hits = ctx.dev_alloc(rays.nbytes), ctx.dev_alloc(rays.size * 24)
    ctx.upload(d_rays, rays)
    for flags in (0, capi.TRACE_REFERENCE_ORDER):
        for any_hit in (False, True):
            bvh.trace_rays_dev(d_rays, rays.size, d_hits, any_hit=any_hit, flags=flags)
            print(depth, flags, any_hit, "overflows", bvh.stack_overflows())
    for flags in (0, capi.TRACE_REFERENCE_ORDER):
        for any_hit in (False, True):
            tmax = np.full(rays.size, 1e-3, np.float32) if any_hit else None
            try:
                got = bvh.trace_rays(rays, any_hit=any_hit, t_max=tmax, flags=flags)
                print(depth, flags, any_hit, "host: no error; hits", int(got["did_hit"].sum()))
            except capi.RtrError as e:
                print(depth, flags, any_hit, "host: error", e.code, str(e)[:80])
