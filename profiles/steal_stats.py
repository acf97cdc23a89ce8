"""Lane occupancy of the persistent traversal kernel before and after its job pool runs empty (diagnostics build):
    RTR_BUILD_ONLY=trace.cu RTR_NVCC_EXTRA=-DRTR_STEAL_STATS python -m realtimeraytracing_b200.build --force
    python profiles/steal_stats.py [bounces]"""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np
import torch
from realtimeraytracing_b200 import capi, synth
bounces = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, W, H = 10_000_000, 3840, 2160
tris, meshes, L = synth.triangle_soup(n)
cam = synth.soup_camera(L, W, H)
ctx = capi.Context(0)
lib = capi.load_library()
bvh = capi.Bvh(ctx).build(tris, meshes)
d_rgba = ctx.dev_alloc(W * H * 16)
st = torch.cuda.ExternalStream(ctx.stream)
out = (C.c_ulonglong * 8)()


def run(label, fn):
    fn(); ctx.sync()
    lib.rtr_debug_steal_stats(out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); fn(); e1.record(st); ctx.sync()
    lib.rtr_debug_steal_stats(out, 0)
    v = [int(x) for x in out]
    print("%s: %.3f ms" % (label, e0.elapsed_time(e1)))
    print("   pool not empty: %d warp bookkeeping rounds, %.1f lanes walking on average" % (v[6], v[7] / max(1, v[6])))
    print("   pool empty:     %d rounds (%.1f %% of all), %.1f lanes walking on average" % (v[4], 100.0 * v[4] / max(1, v[4] + v[6]), v[5] / max(1, v[4])))
    print("   steal passes %d: idle lanes %.1f, donors %.1f, pairs %.2f per pass" % (v[0], v[1] / max(1, v[0]), v[2] / max(1, v[0]), v[3] / max(1, v[0])))


run("full frame, %d bounces" % bounces, lambda: bvh.render_dev(cam, W, H, d_rgba, bounces=bounces))
run("8/57 of the frame", lambda: bvh.render_stripes_dev(cam, W, H, d_rgba, 16, [1] + [8] * 7, 1, bounces=bounces))
run("1/57 of the frame", lambda: bvh.render_stripes_dev(cam, W, H, d_rgba, 16, [1] + [8] * 7, 0, bounces=bounces))
