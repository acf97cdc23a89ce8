"""What a rebuild and a BVH broadcast cost each other when they run on the same GPU at the same time, and what a
concurrent receive costs the rays (2 ranks):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 profiles/overlap_n2.py [reserve] [bounces]"""
import os, sys
sys.path.insert(0, '.')
import numpy as np
import torch
import torch.distributed as dist
from realtimeraytracing_b200 import capi, parallel, synth
from realtimeraytracing_b200.layouts import TRIANGLE

reserve = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bounces = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_MAX_NCHANNELS", str(reserve))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, W, H = 10_000_000, 3840, 2160
ctx = capi.Context(lr)
main = torch.cuda.Stream(device=dev)
sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
ctx.set_stream(main.cuda_stream)
comm = parallel.RankComm(ctx)
L = synth.soup_extent(n)
cam = synth.soup_camera(L, W, H)
tris, meshes, _ = synth.triangle_soup(n)
d_tris = torch.from_numpy(tris.view(np.uint8).reshape(-1)).to(dev)
d_meshes = torch.from_numpy(meshes.view(np.uint8).reshape(-1).copy()).to(dev)
d_img = torch.zeros(H * W * 4, dtype=torch.float32, device=dev)
A, B = capi.Bvh(ctx), capi.Bvh(ctx)
torch.cuda.synchronize()


def say(msg):
    sys.stdout.write("[rank %d] %s\n" % (rank, msg)); sys.stdout.flush()


def build(b):
    b.build_dev(d_tris.data_ptr(), n, n, d_meshes.data_ptr(), 1)


def bcast(b):
    b.broadcast(0, traversal_only=True, expected_triangles=n)


def rays(b, layout, r):
    b.render_stripes_dev(cam, W, H, d_img.data_ptr(), 16, layout, r, bounces=bounces)


def sync_all():
    for s in (main, sa, sb):
        s.synchronize()
    dist.barrier()
    torch.cuda.synchronize()


def span(jobs, reps=3):
    """jobs: list of (stream, fn); all start together, per-job ms and the overall span, mean of reps"""
    out = np.zeros((reps, len(jobs) + 1))
    for rep in range(reps):
        sync_all()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        ends = []
        for st, fn in jobs:
            st.wait_event(e0)
            ctx.switch_stream(st.cuda_stream)
            if fn is not None:
                fn()
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            ends.append(e)
        ctx.switch_stream(main.cuda_stream)
        sync_all()
        t = [e0.elapsed_time(e) for e in ends]
        out[rep, :-1] = t
        out[rep, -1] = max(t)
    return out.mean(axis=0)


# both ranks hold a traversable BVH in A and in B
for b in (A, B):
    if rank == 0:
        build(b)
    bcast(b)
sync_all()
ctx.reserve_sms(reserve)

t = span([(sa, (lambda: build(A)) if rank == 0 else None)])
say("rebuild alone                     %.2f ms" % t[0])
t = span([(sb, lambda: bcast(B))])
say("broadcast alone                   %.2f ms" % t[0])
if rank == 0:
    ctx.profile_enable(True)
t = span([(sb, lambda: bcast(B)), (sa, (lambda: build(A)) if rank == 0 else None)])
say("broadcast || rebuild (rank 0)     bcast %.2f  rebuild %.2f  span %.2f ms" % (t[0], t[1], t[2]))
if rank == 0:
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    say("  rebuild kernels beside the broadcast (3 runs, ms per run): " +
        "  ".join("%s %.2f" % (k.split("(")[0][-28:], v[0] / 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]))
    ctx.profile_enable(True)
    span([(sa, lambda: build(A))])
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    say("  rebuild kernels alone: " +
        "  ".join("%s %.2f" % (k.split("(")[0][-28:], v[0] / 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]))
else:
    span([(sa, None)])


def delayed(fn, st, us):
    def go():
        with torch.cuda.stream(st):
            torch.cuda._sleep(int(us * 1.9e3))   # spin, ~1.9 GHz
        fn()
    return go


# who goes first matters: an NCCL kernel that becomes due while a rebuild is running
for us in (0, 100, 1000, 3000):
    t = span([(sa, (lambda: build(A)) if rank == 0 else None), (sb, delayed(lambda: bcast(B), sb, us))])
    say("rebuild first, broadcast %4d us later      rebuild %.2f  bcast %.2f (from the start of the span)" % (us, t[0], t[1]))
for us in (100, 500):
    t = span([(sb, lambda: bcast(B)), (sa, delayed((lambda: build(A)) if rank == 0 else (lambda: None), sa, us))])
    say("broadcast first, rebuild %4d us later      bcast %.2f  rebuild %.2f" % (us, t[0], t[1]))
t = span([(sa, (lambda: (build(A), build(A))) if rank == 0 else None), (sb, delayed(lambda: bcast(B), sb, 1000))])
say("two rebuilds back to back, broadcast 1000 us after the start: rebuilds %.2f  bcast %.2f" % (t[0], t[1]))

for layout, r, label in (([1] + [8] * 7, 1, "8/57 of the frame"),):
    t = span([(main, lambda: rays(A, layout, r))])
    say("rays alone, %s        %.2f ms" % (label, t[0]))
    t = span([(sb, lambda: bcast(B)), (main, lambda: rays(A, layout, r))])
    say("rays || broadcast, %s  bcast %.2f  rays %.2f  span %.2f ms" % (label, t[0], t[1], t[2]))
ctx.reserve_sms(0)
t = span([(main, lambda: rays(A, [1] + [8] * 7, 1))])
say("rays alone, 8/57, no SMs reserved %.2f ms" % t[0])
sync_all()
say("done")
dist.destroy_process_group()
