"""Does a primary-context build still get the whole device once a green-context partition exists?"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
from realtimeraytracing_b200 import capi, synth
n = 2_000_000
tris, meshes, L = synth.triangle_soup(n)
ctx = capi.Context(0)
base = torch.cuda.Stream(); ctx.switch_stream(base.cuda_stream)
d_tris, d_meshes = ctx.dev_alloc(tris.nbytes), ctx.dev_alloc(meshes.nbytes)
ctx.upload(d_tris, tris); ctx.upload(d_meshes, meshes)
bvh = capi.Bvh(ctx)
def timed_build(label):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bvh.build_dev(d_tris, n, n, d_meshes, 1); ctx.sync()
    e0.record(base); bvh.build_dev(d_tris, n, n, d_meshes, 1); e1.record(base); ctx.sync()
    print(label, "%.3f ms" % e0.elapsed_time(e1), flush=True)
timed_build("before the partition exists:")
gs, got = ctx.partition_sms(136, 2)
print("partition", got, flush=True)
timed_build("after, on a primary-context stream:")
