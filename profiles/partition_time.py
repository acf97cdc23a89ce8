"""Green-context SM partition: rays of consecutive frames overlapped on two streams of the partition.
    python profiles/partition_time.py [sms] [bounces] [frames] [row_fraction_denominator]"""
import hashlib, sys
sys.path.insert(0, '.')
import numpy as np
import torch
from realtimeraytracing_b200 import capi, synth
sms = int(sys.argv[1]) if len(sys.argv) > 1 else 136
bounces = int(sys.argv[2]) if len(sys.argv) > 2 else 4
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 8
share = int(sys.argv[4]) if len(sys.argv) > 4 else 7     # each "frame" is 1/share of the image (a worker's part at N=8)
n, W, H = 10_000_000, 3840, 2160
tris, meshes, L = synth.triangle_soup(n)
cam = synth.soup_camera(L, W, H)
ctx = capi.Context(0)
bvhs = [capi.Bvh(ctx).build(tris, meshes) for _ in range(2)]   # one per stream: a BVH owns the job counter of its launch
bvh = bvhs[0]
d_rgba = [ctx.dev_alloc(W * H * 16) for _ in range(2)]
main = torch.cuda.ExternalStream(ctx.stream)
layout = [1] + [8] * share if share > 1 else [1]
rank = 1 if share > 1 else 0


def one(j):
    if share > 1:
        bvhs[j].render_stripes_dev(cam, W, H, d_rgba[j], 16, layout, rank, bounces=bounces)
    else:
        bvhs[j].render_dev(cam, W, H, d_rgba[j], bounces=bounces)


def run(streams, label):
    for st in streams:
        torch.cuda.ExternalStream(st).synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    first = torch.cuda.ExternalStream(streams[0])
    e0.record(first)
    for st in streams[1:]:
        torch.cuda.ExternalStream(st).wait_event(e0)
    for f in range(frames):
        ctx.switch_stream(streams[f % len(streams)])
        one(f % 2)
    for st in streams[1:]:
        first.wait_stream(torch.cuda.ExternalStream(st))
    e1.record(first)
    first.synchronize()
    print("%-44s %.3f ms per frame" % (label, e0.elapsed_time(e1) / frames))


base_stream = torch.cuda.Stream()       # (switching to the ctx's own stream would retire it)
base = base_stream.cuda_stream
ctx.switch_stream(base)
for _ in range(2):
    one(0)
ctx.sync()
run([base], "whole device, one stream")
gs, got = ctx.partition_sms(sms, 2)
print("partition: %d SMs" % got)
for s in gs:
    ctx.switch_stream(s); one(0)
torch.cuda.synchronize()
run([gs[0]], "partition, one stream")
run(gs, "partition, two streams (frames overlap)")
ctx.switch_stream(base)
img = np.zeros((H, W, 4), np.float32)
for j in range(2):
    ctx.download(img, d_rgba[j])
    print("frame digest", j, hashlib.blake2b(img.view(np.uint8).tobytes(), digest_size=8).hexdigest(), "overflows", bvhs[j].stack_overflows())
