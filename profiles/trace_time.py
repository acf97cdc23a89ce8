"""Times the rays of BASELINE config 4 alone (10 M-triangle soup already built, 3840x2160, primary + N bounces) and
prints the frame digest: experiments with compile-time / environment knobs of the traversal kernel.
    python profiles/trace_time.py [bounces] [reps]"""
import hashlib, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
from realtimeraytracing_b200 import capi, synth
bounces = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n, W, H = 10_000_000, 3840, 2160
tris, meshes, L = synth.triangle_soup(n)
cam = synth.soup_camera(L, W, H)
ctx = capi.Context(0)
bvh = capi.Bvh(ctx).build(tris, meshes)
d_rgba = ctx.dev_alloc(W * H * 16)
d_rays = ctx.dev_alloc(8)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
st = torch.cuda.ExternalStream(ctx.stream)
for _ in range(2):
    bvh.render_dev(cam, W, H, d_rgba, rays_dev=d_rays, bounces=bounces)
ctx.sync()
ev[0].record(st)
for i in range(reps):
    bvh.render_dev(cam, W, H, d_rgba, rays_dev=d_rays, bounces=bounces)
    ev[i + 1].record(st)
ctx.sync()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
img = np.zeros((H, W, 4), np.float32)
ctx.download(img, d_rgba)
print("rays ms: min %.3f median %.3f  overflow %d  digest %s" % (min(ms), float(np.median(ms)), bvh.stack_overflows(),
      hashlib.blake2b(img.view(np.uint8).tobytes(), digest_size=8).hexdigest()))
# what a launch costs whatever it traces: 1/57 and 8/57 of the frame (a builder's / a worker's part at 8 GPUs), 4 bounces
for lay_rank, label in ((0, "1/57"), (1, "8/57")):
    t = []
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        bvh.render_stripes_dev(cam, W, H, d_rgba, 16, [1] + [8] * 7, lay_rank, bounces=4)
        e1.record(st)
        ctx.sync()
        t.append(e0.elapsed_time(e1))
    print("%s of the frame, 4 bounces: min %.3f median %.3f ms" % (label, min(t[1:]), float(np.median(t[1:]))))
