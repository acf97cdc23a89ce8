#!/usr/bin/env python
"""bench.py -- headline benchmark of the acceleration-structure + ray-cast path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched with torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): a 10M-triangle synthetic
soup, FULL rebuild per frame (scene box + Morton -> Onesweep sort -> PLOC -> flatten), then one 3840x2160
frame of primary + 2 bounce rays.  One "step" is one such frame.  With N>1 the rebuild stays on rank 0
(PLOC does not shard), its result is broadcast over NVLink, image row blocks are dealt to the ranks and
all-gathered (configs[4] layout, same bounce count so the per-N values are comparable): strong scaling.

value  = rays traced by all ranks / device time of the step, inputs resident in HBM
e2e    = the same through the host-pointer C ABI: triangles uploaded from pinned host memory every step
         (rtr_bvh_build), image read back to pinned host memory every step
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mrays/s"
WORKLOAD = "10M-triangle soup, full BVH rebuild per frame + 3840x2160 primary + 2-bounce rays"            # BASELINE config 4
WORKLOAD5 = "10M-triangle soup, full BVH rebuild per frame + 3840x2160 4-bounce paths, BVH broadcast, tiles sharded"  # config 5


def workload_name(tris, w, h, bounces):
    if (tris, w, h, bounces) == (10_000_000, 3840, 2160, 2):
        return WORKLOAD
    if (tris, w, h, bounces) == (10_000_000, 3840, 2160, 4):
        return WORKLOAD5
    return "%d-triangle soup, full BVH rebuild per frame + %dx%d primary + %d-bounce rays" % (tris, w, h, bounces)


def image_hash(rgba: np.ndarray) -> str:
    """64-bit digest of a frame's rgba32f bytes: equal across GPU counts iff the frames are bit-identical."""
    import hashlib
    return hashlib.blake2b(np.ascontiguousarray(rgba).view(np.uint8).tobytes(), digest_size=8).hexdigest()


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--tris", type=int, default=10_000_000)
    p.add_argument("--width", type=int, default=3840)
    p.add_argument("--height", type=int, default=2160)
    p.add_argument("--bounces", type=int, default=None,
                   help="default: 2 on one GPU (BASELINE config 4), 4 on several (config 5: 4-bounce paths, tiles sharded)")
    p.add_argument("--rows-per-block", type=int, default=16)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the per-stage / per-kernel side measurements")
    p.add_argument("--reference-order", action="store_true", help="trace in the shader's exact visiting order (no pruning)")
    p.add_argument("--reserve-sms", type=int, default=0,
                   help="N>1: SMs the traversal leaves free for the NCCL kernels of the broadcast in flight")
    p.add_argument("--partition", action="store_true",
                   help="N>1, experimental: rays in a green-context SM partition, two frames in flight (slower in the full pipeline "
                        "on this box: the exchange chain, not the rays, sets the frame time; DESIGN.md section 7)")
    p.add_argument("--no-pipeline", action="store_true", help="N>1: rebuild, broadcast, render and gather strictly in sequence")
    p.add_argument("--record-hash", action="store_true", help="N=1: store this frame's digest in profiles/image_hashes.json")
    return p.parse_args()


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own bvh.cpp (oracle/_ref) for the rebuild + the oracle's restatement of
# raytracer.glsl for the rays, on the host cores.  Used for cpu_baseline and --impl reference.
# This is the ONLY place bench.py touches oracle/.
# ----------------------------------------------------------------------------------------------
def cpu_sample_sizes(args, budget_s):
    """Bounded sample of the workload: triangles and pixels scaled by the same factor f, chosen from a
    quick calibration (65 536-triangle rebuild + a 256x144 frame) so one step costs about budget_s."""
    cal = CpuArm(args, 65536, 256, 144)
    rays, secs, build_s = cal.step(split=True)
    per_tri = build_s / cal.n * 1.4            # the reference's build cost grows slightly faster than n
    per_ray = (secs - build_s) / max(1, rays) * 2.5   # unpruned walks visit ~N^(1/3) more nodes at scale
    full = args.tris * per_tri + args.width * args.height * 2.4 * per_ray
    f = min(0.1, max(0.001, budget_s / full))
    n = min(int(args.tris * f), 1_000_000)
    w = int(round(args.width * f ** 0.5 / 16)) * 16
    h = int(round(args.height * f ** 0.5 / 16)) * 16
    return max(n, 1000), max(w, 64), max(h, 32)


class CpuArm:
    """Reference build: OMP_NUM_THREADS=1, the reference's fastest configuration -- its OpenMP path is
    slower than serial -- rays: all host cores."""

    def __init__(self, args, n, w, h):
        os.environ["OMP_NUM_THREADS"] = "1"
        from oracle import Oracle, Reference, reference_available
        from realtimeraytracing_b200 import synth
        self.o = Oracle()
        self.n, self.w, self.h, self.bounces = n, w, h, args.bounces
        self.tris, self.meshes, L = synth.triangle_soup(n)
        self.cam = synth.soup_camera(L, w, h)
        cap = 65536 if n <= 65536 else 1048576
        self.ref = Reference(cap) if reference_available(cap) else None
        self.kind = "reference" if self.ref is not None else "port"
        self.cores = os.cpu_count() or 1

    def step(self, split=False):
        """One frame: rebuild + flatten + all rays.  Returns (rays, seconds[, build seconds])."""
        t0 = time.perf_counter()
        if self.ref is not None:
            rb = self.ref.bvh_build(self.tris, self.meshes, want_morton=False)  # the reference's own bvh.cpp
            clusters, left, right = rb.clusters, rb.left, rb.right
        else:
            ob = self.o.bvh_build(self.tris, self.meshes)
            clusters, left, right = ob.clusters, ob.left, ob.right
        flat = self.o.flatten(clusters, left, right)              # scene.cpp:189-208 restated (GL file, not compilable)
        t1 = time.perf_counter()
        _, _, rays = self.o.render(flat, self.tris, self.meshes, self.cam, self.w, self.h, self.w, self.h,
                                   bounces=self.bounces, threads=self.cores)  # raytracer.glsl restated, OpenMP over rows
        t2 = time.perf_counter()
        return (rays, t2 - t0, t1 - t0) if split else (rays, t2 - t0)

    def describe(self):
        build = "reference bvh.cpp via oracle/_ref, 1 thread (its OpenMP path is slower)" \
            if self.kind == "reference" else "oracle C port, 1 thread"
        return "%d-triangle soup rebuild [%s] + %dx%d primary+%d-bounce rays [oracle traversal, %d threads]" % (
            self.n, build, self.w, self.h, self.bounces, self.cores)


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_stage_baselines(args):
    """every stage of the path on the host cores of THIS box, next to the GPU numbers --
    the reference's own code where it compiles (bvh.cpp, its std::sort of pairs), a host radix sort of the same design
    as the Onesweep kernel, the oracle's traversal.  Thread counts are stated per entry.  (oracle/, CPU arm only.)"""
    from oracle import Oracle, Reference, reference_available
    from realtimeraytracing_b200 import synth
    o = Oracle()
    cores = os.cpu_count() or 1
    out = {"cores_available": cores, "cpu": cpu_model()}
    ref = Reference(65536) if reference_available(65536) else None
    for label, n in (("10M", 10_000_000), ("1M_config1", 1_000_000)):
        keys = synth.random_keys_u32(n, seed=1)
        idx = np.arange(n, dtype=np.uint32)
        if ref is not None:
            _, _, ms = ref.sort_pairs_ms(keys)
            out["sort_pairs_%s_std_sort" % label] = {"ms": ms, "gkeys_s": n / ms / 1e6, "cores": 1,
                                                     "what": "std::sort of pair<code,index>, the reference's sort (bvh.cpp:223-231)"}
        for th in sorted({1, cores}):
            k2, v2 = keys.copy(), idx.copy()   # sorted in place
            t0 = time.perf_counter()
            o.lib.orc_radix_sort_pairs_mt(k2.ctypes.data, v2.ctypes.data, n, th)
            ms = (time.perf_counter() - t0) * 1e3
            assert k2[0] <= k2[n // 2] <= k2[-1]
            out["sort_pairs_%s_host_lsd_%dt" % (label, th)] = {"ms": ms, "gkeys_s": n / ms / 1e6, "cores": th,
                                                              "what": "stable LSD radix-256 sort of (key, index) pairs, OpenMP"}
    # the reference's bvh.cpp (capacity-patched): 1 thread at 1 M triangles, and 1 vs all threads at 262 144
    if reference_available(1048576):
        big = Reference(1048576)
        for n, threads in ((1_000_000, 1), (262_144, 1), (262_144, cores)):
            tris, meshes, _ = synth.triangle_soup(n)
            big.set_threads(threads)
            ms = big.bvh_build(tris, meshes, want_morton=False, timing_only=True).build_ms
            out["bvh_cpp_build_%d_tris_%dt" % (n, threads)] = {"ms": ms, "mtris_s": n / ms / 1e3, "cores": threads,
                                                            "what": "cr::BVH constructor of the reference (bvh.cpp:11-24), OpenMP threads as stated"}
        big.set_threads(1)
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    n, w, h = cpu_sample_sizes(args, 150.0 / max(1, total))
    arm = CpuArm(args, n, w, h)
    for _ in range(args.warmup):
        arm.step()
    rays, secs = 0, 0.0
    for _ in range(args.steps):
        r, s = arm.step()
        rays += r; secs += s
    val = rays / secs / 1e6
    line = {
        "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args.tris, args.width, args.height, args.bounces), "sample": arm.describe()},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": arm.cores, "kind": arm.kind, "sample": arm.describe()},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
ALGO_BYTES = {  # algorithmic HBM bytes per launch as a function of (n triangles / keys); DESIGN.md section 4
    "onesweep_kernel<u32,pairs>": lambda n: 16.0 * n,   # read 8 + write 8 per (code,index) pair and pass
    "onesweep_kernel<u32,keys>": lambda n: 8.0 * n,
    "radix_histogram_kernel": lambda n: 4.0 * n,
    "scene_aabb_kernel": lambda n: 64.0 * n,
    "morton_kernel": lambda n: 72.0 * n,                 # 64 read + code 4 + index 4
    "leaf_init_kernel": lambda n: 116.0 * n,             # index 4 + triangle 64 + node 48 (32-byte record + active id + the traversal copy is extra)
}


def main():
    args = parse_args()
    if os.environ.get("RTR_BENCH_WATCHDOG"):   # debugging aid: dump every thread's stack and quit after that many seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["RTR_BENCH_WATCHDOG"]), exit=True)
    if args.bounces is None:
        args.bounces = 2 if max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) == 1 else 4
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from realtimeraytracing_b200 import build as rbuild, capi, parallel, synth
    from realtimeraytracing_b200.layouts import MESH, TRIANGLE

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs `python -m torch.distributed.run --nproc-per-node %d bench.py ...`" % (args.gpus, args.gpus))
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: librtr_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.reserve_sms <= 0:
        # SMs left to NCCL: the rays run in a green-context partition of the other SMs (a multiple of 8 on sm_100, so
        # 136 + 12), or -- without the partition -- vacate whole SMs (rtr_ctx_reserve_sms)
        args.reserve_sms = 12 if args.partition else (8 if world <= 2 else 16)
    if world > 1:
        if not args.no_pipeline:
            # the broadcast of the next frame's BVH runs beside the rays of the current one on --reserve-sms SMs:
            # NCCL gets as many channels (one CTA each) as there are SMs left to it
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(max(1, args.reserve_sms)))
        dist.init_process_group("nccl", device_id=dev)
    rbuild.build()

    ctx = capi.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    comm = parallel.RankComm(ctx) if world > 1 else None

    n, W, H, bounces, rpb = args.tris, args.width, args.height, args.bounces, args.rows_per_block
    flags = capi.TRACE_REFERENCE_ORDER if args.reference_order else capi.TRACE_DEFAULT
    dw, dh = synth.reference_denominators(W, H)
    L = synth.soup_extent(n)
    cam = synth.soup_camera(L, W, H)

    # ---- inputs: generated on the host once, resident in HBM for `value`, pinned for `e2e` ----
    tris_pinned = meshes_np = slice_pinned = None
    d_tris = d_meshes = None
    slice_lo, slice_hi = parallel.slice_range(n, rank, world)
    if world > 1 and rank != 0:
        # e2e with sharded host traffic: this rank uploads triangles [slice_lo, slice_hi) over its own PCIe link
        tris_np, _, _ = synth.triangle_soup(n)
        slice_pinned = torch.empty((slice_hi - slice_lo) * TRIANGLE.itemsize, dtype=torch.uint8, pin_memory=True)
        slice_pinned.numpy().view(TRIANGLE)[:] = tris_np[slice_lo:slice_hi]
        del tris_np
    if rank == 0:
        tris_np, meshes_np, _ = synth.triangle_soup(n)
        tris_pinned = torch.empty(n * TRIANGLE.itemsize, dtype=torch.uint8, pin_memory=True)
        tris_pinned.numpy().view(TRIANGLE)[:] = tris_np
        del tris_np
        slice_pinned = tris_pinned[slice_lo * TRIANGLE.itemsize: slice_hi * TRIANGLE.itemsize]
        with torch.cuda.stream(stream):
            d_tris = torch.empty(n * TRIANGLE.itemsize, dtype=torch.uint8, device=dev)
            d_tris.copy_(tris_pinned, non_blocking=True)
            d_meshes = torch.from_numpy(meshes_np.view(np.uint8).copy()).to(dev)
    with torch.cuda.stream(stream):
        d_rgba = torch.zeros(H * W * 4, dtype=torch.float32, device=dev)
        d_rays = torch.zeros(1, dtype=torch.int64, device=dev)
    rgba_pinned = torch.empty(H * W * 4, dtype=torch.float32, pin_memory=True) if rank == 0 else None
    stream.synchronize()

    bvh = capi.Bvh(ctx)

    def frame_device():
        """value path: inputs resident in HBM."""
        if rank == 0:
            bvh.build_dev(d_tris.data_ptr(), n, n, d_meshes.data_ptr(), 1)
        if world > 1:
            bvh.broadcast(0)
        bvh.render_sharded_dev(cam, W, H, d_rgba.data_ptr(), rpb, rank, world, rays_dev=d_rays.data_ptr(),
                               bounces=bounces, flags=flags)
        if world > 1:
            ctx.allgather_rows(d_rgba.data_ptr(), W, H, 16, rpb)

    def frame_e2e():
        """e2e path: host triangles in, host image out, through the host-pointer C ABI."""
        if rank == 0:
            bvh.build(tris_pinned.numpy().view(TRIANGLE), meshes_np)      # H2D of the triangles inside rtr_bvh_build
        if world > 1:
            bvh.broadcast(0)
        bvh.render_sharded_dev(cam, W, H, d_rgba.data_ptr(), rpb, rank, world, rays_dev=d_rays.data_ptr(),
                               bounces=bounces, flags=flags)
        if world > 1:
            ctx.allgather_rows(d_rgba.data_ptr(), W, H, 16, rpb)
        if rank == 0:
            ctx.download(rgba_pinned.numpy(), d_rgba.data_ptr())          # D2H of the frame (synchronises)
        else:
            ctx.sync()

    def barrier():
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """device time of `steps` calls: events on the launching stream, barrier + sync both sides, max over ranks"""
        barrier()
        with torch.cuda.stream(stream):
            d_rays.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = ctx.launch_count
        stream.synchronize()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return reduce_timing(e0.elapsed_time(e1), launches0)

    def reduce_timing(ms, launches0):
        launches = ctx.launch_count - launches0
        rays = int(d_rays.item())
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            r = torch.tensor([rays, launches], dtype=torch.int64, device=dev)
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
            rays, launches = int(r[0].item()), int(r[1].item())
        return ms, rays, launches

    # ---- N > 1: pipelined frames, three stages on three streams: rank 0 rebuilds frame f+2 (stream A), every rank
    #      exchanges frame f+1 (stream B), every rank traces its rows of frame f (stream R); three BVHs rotate.  The
    #      row blocks are dealt with weights so that the rebuilding rank, which has less time left for rays, gets
    #      fewer of them (parallel.stripe_layout), and the frame is gathered on rank 0 (the rank that hands it on). ----
    pipelined = world > 1 and not args.no_pipeline
    layout = [1] * world
    rank_zero_defers = False
    if pipelined:
        NB = 3
        stream_a = torch.cuda.Stream(device=dev)
        stream_b = torch.cuda.Stream(device=dev)
        stream_r = stream
        bvhs = [bvh] + [capi.Bvh(ctx) for _ in range(NB - 1)]
        built = [torch.cuda.Event() for _ in range(NB)]       # rank 0: BVH k rebuilt (not yet sent)
        ready = [torch.cuda.Event() for _ in range(NB)]       # BVH k rebuilt and received
        released = [torch.cuda.Event() for _ in range(NB)]    # the rays of the frame that used BVH k are done
        sent = [torch.cuda.Event() for _ in range(NB)]        # rank 0: BVH k has left (it may be rebuilt)
        last_built = [None]

        # e2e arm, host traffic SHARDED over the ranks' own PCIe links: every rank uploads its slice of the frame's
        # triangles (rtr_dev_upload_async), the slices are assembled on rank 0 over NVLink (rtr_gather_slices), and every
        # rank downloads the rows it traced straight into one pinned host image all ranks share
        # (rtr_download_stripes_async) -- the frame is never gathered on a device.  All double-buffered.
        x_in, x_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        slice_bytes = (slice_hi - slice_lo) * TRIANGLE.itemsize
        with torch.cuda.stream(stream):
            x_rgba = [d_rgba, torch.empty_like(d_rgba)]
            x_tris = [d_tris, torch.empty_like(d_tris)] if rank == 0 else [None, None]
            x_slice = [torch.empty(max(slice_bytes, 16), dtype=torch.uint8, device=dev) for _ in range(2)]
        shm_names = ["/dev/shm/rtr_bench_%s_img%d" % (os.environ.get("MASTER_PORT", "0"), j) for j in range(2)]
        if rank == 0:
            for nm in shm_names:
                np.memmap(nm, dtype=np.float32, mode="w+", shape=(H, W, 4)).flush()
        dist.barrier()
        x_host_img = [np.memmap(nm, dtype=np.float32, mode="r+", shape=(H, W, 4)) for nm in shm_names]
        for img in x_host_img:
            capi.host_register(img)
        x_meshes_pinned = torch.from_numpy(meshes_np.view(np.uint8).copy()).pin_memory() if rank == 0 else None
        x_up_done, x_slice_free, x_tris_ready, x_tris_free, x_frame_done, x_img_free, x_img_done = (
            [torch.cuda.Event(), torch.cuda.Event()] for _ in range(7))
        # The rays of consecutive frames go to two streams of a green-context partition of all SMs but the ones left to
        # NCCL: frame f+1's persistent launch fills in as the last, longest paths of frame f finish (about 0.7 ms of
        # every launch), and NCCL's kernels always find free SMs.  (Without the partition: one stream -- two launches
        # in flight would feed the later one's CTAs into the SMs the earlier one vacated for NCCL, where they leave.)
        rays_sms = 0
        if not args.partition:
            streams_r = [stream_r, stream_r]
            ctx.reserve_sms(args.reserve_sms)
        else:
            handles, rays_sms = ctx.partition_sms(ctx.sm_count - args.reserve_sms, 2)
            streams_r = [torch.cuda.ExternalStream(h, device=dev) for h in handles]
            if os.environ.get("RTR_BENCH_ONE_RAY_STREAM"):
                streams_r[1] = streams_r[0]
            # (launches on the partition's streams ignore this; the cooperative rebuild kernels leave these SMs alone)
            ctx.reserve_sms(args.reserve_sms)
        stream.synchronize()
        marks = []

        def mark(what, f, st):
            if os.environ.get("RTR_BENCH_TRACE"):
                e = torch.cuda.Event(enable_timing=True)
                e.record(st)
                marks.append((what, f, e))

        def submit_upload(f):
            """e2e, every rank: its slice of frame f's host triangles goes up its own PCIe link, then the slices are
            assembled on rank 0 over NVLink (an NCCL point: issued by all ranks at the same place of the schedule)."""
            j = f % 2
            ctx.switch_stream(x_in.cuda_stream)
            x_in.wait_event(x_slice_free[j])
            if slice_bytes:
                ctx.upload_async(x_slice[j].data_ptr(), slice_pinned.numpy())
            if rank == 0:
                ctx.upload_async(d_meshes.data_ptr(), x_meshes_pinned.numpy())
            x_up_done[j].record(x_in)
            ctx.switch_stream(stream_b.cuda_stream)
            stream_b.wait_event(x_up_done[j])
            if rank == 0:
                stream_b.wait_event(x_tris_free[j])
            ctx.gather_slices(x_slice[j].data_ptr(), x_tris[j].data_ptr() if rank == 0 else None, n, TRIANGLE.itemsize, 0)
            x_slice_free[j].record(stream_b)
            if rank == 0:
                x_tris_ready[j].record(stream_b)
            mark("tris", f, stream_b)

        def submit_build(f, e2e):
            """stage A, rank 0 only: rebuild BVH f % NB (asynchronous: the host does not wait for the device)."""
            if rank != 0:
                return
            k = f % NB
            ctx.switch_stream(stream_a.cuda_stream)
            stream_a.wait_event(released[k])   # own rays of the frame that used this BVH
            stream_a.wait_event(sent[k])       # and its broadcast
            if e2e:
                j = f % 2
                stream_a.wait_event(x_tris_ready[j])
                bvhs[k].build_dev(x_tris[j].data_ptr(), n, n, d_meshes.data_ptr(), 1)
                x_tris_free[j].record(stream_a)
            else:
                bvhs[k].build_dev(d_tris.data_ptr(), n, n, d_meshes.data_ptr(), 1)
            built[k].record(stream_a)
            last_built[0] = built[k]
            mark("built", f, stream_a)

        def submit_exchange(f, after=None):
            """stage B, every rank: BVH f % NB travels from rank 0 (64 B*(2n-1), enqueue only)."""
            k = f % NB
            ctx.switch_stream(stream_b.cuda_stream)
            if rank == 0:
                stream_b.wait_event(built[k])
                if after is not None:
                    stream_b.wait_event(after)
            else:
                stream_b.wait_event(released[k])   # the rays of the frame that used this receive buffer
            bvhs[k].broadcast(0, traversal_only=not args.reference_order, expected_triangles=n)
            ready[k].record(stream_b)
            if rank == 0:
                sent[k].record(stream_b)
            mark("bcast", f, stream_b)

        def submit_rays(f, e2e):
            """stage R, every rank: the rays of its stripes of frame f, then the frame is gathered on rank 0."""
            k = f % NB
            j = f % 2
            sr = streams_r[j]
            ctx.switch_stream(sr.cuda_stream)
            sr.wait_event(ready[k])
            sr.wait_event(x_img_done[j])          # frame f-2 has left this image buffer
            if rank == 0 and layout[0] > 0 and last_built[0] is not None:
                # the persistent traversal kernel would hold the SMs a rebuild needs: on the building rank the rays
                # start once the rebuild submitted last is through
                sr.wait_event(last_built[0])
            mark("rays0", f, sr)
            img = x_rgba[j]
            if e2e:
                sr.wait_event(x_img_free[j])   # the download of the frame that used this image buffer is through
            bvhs[k].render_stripes_dev(cam, W, H, img.data_ptr(), rpb, layout, rank, rays_dev=d_rays.data_ptr(),
                                       bounces=bounces, flags=flags)
            released[k].record(sr)
            mark("rays1", f, sr)
            if e2e:
                # every rank sends the rows it traced to the shared host image over its own PCIe link
                x_frame_done[j].record(sr)
                x_img_done[j].record(sr)
                ctx.switch_stream(x_out.cuda_stream)
                x_out.wait_event(x_frame_done[j])
                ctx.download_stripes_async(x_host_img[j].ctypes.data, img.data_ptr(), W, H, 16, rpb, layout)
                x_img_free[j].record(x_out)
                ctx.switch_stream(sr.cuda_stream)
            elif args.partition or (rank_zero_defers and not os.environ.get("RTR_BENCH_NO_DEFER")):
                # the frame is gathered on rank 0 by the communicator's stream, issued by every rank at the same place of
                # the schedule (after exchange(f+1)): the next frame's rays follow at once instead of waiting for rank 0 to
                # receive (RTR_BENCH_TRACE at 2 GPUs: 1.1 ms of every 23.5 ms frame of rank 1).  Not at 8 GPUs: measured
                # there, the broadcasts queued behind these gathers take 8 ms instead of 3.6 (2359 against 2694 Mrays/s).
                x_frame_done[j].record(sr)
                ctx.switch_stream(stream_b.cuda_stream)
                stream_b.wait_event(x_frame_done[j])
                ctx.gather_stripes(img.data_ptr(), W, H, 16, rpb, layout, 0)
                x_img_done[j].record(stream_b)
                mark("gather", f, stream_b)
                ctx.switch_stream(sr.cuda_stream)
            else:
                ctx.gather_stripes(img.data_ptr(), W, H, 16, rpb, layout, 0)
                x_img_done[j].record(sr)
                mark("gather", f, sr)

        def run_pipelined(steps, e2e):
            # every rank issues its NCCL calls in the same order:
            #   value: exchange(f+1), gather(f), exchange(f+2), ...     e2e: slices(f+2), exchange(f+1), slices(f+3), ...
            # (the slices of f+2 go BEFORE the exchange of f+1 on the communicator's stream: they only wait for host
            #  uploads, whereas the exchange waits for the rebuild of f+1 -- so rank 0 can rebuild f+2 beside it)
            last_built[0] = None
            # Where rank 0 traces a good part of the frame itself (2 and 4 GPUs), its broadcast of f+1 is held back until
            # its rebuild of f+2 is through: it then travels beside rank 0's rays of f, which make room for NCCL, instead of
            # beside the rebuild -- RTR_BENCH_TRACE at 2 GPUs: rebuild 9 ms and broadcast 6.7 ms side by side, 4.9 and 4.3
            # apart.  At 8 GPUs rank 0 hardly traces and the workers would wait for the later broadcast: not done there.
            # (Holding it back in the e2e arm of 8 GPUs as well -- there the broadcast starts with rank 0's next rebuild and
            # takes 8.8 ms beside it -- was measured: the broadcast drops to 3.5 ms, but the slices of the frame after next
            # then queue behind it on the communicator and the rebuild behind them: 2 046 against 2 252 Mrays/s.  It needs the
            # slices issued one frame earlier; not done, no GPU time left to validate it.)
            hold = bool(rank_zero_defers and not os.environ.get("RTR_BENCH_NO_DEFER"))
            for op, f in parallel.frame_schedule(steps, e2e, hold):   # the issue order, with its invariants: parallel.py
                if op == "upload":
                    submit_upload(f)
                elif op == "build":
                    submit_build(f, e2e)
                elif op == "exchange":
                    submit_exchange(f)
                elif op == "exchange_after_build":
                    submit_exchange(f, after=last_built[0])
                else:
                    submit_rays(f, e2e)

        def timed_pipelined(steps, e2e):
            barrier(); stream_a.synchronize(); stream_b.synchronize(); streams_r[0].synchronize(); streams_r[1].synchronize()
            del marks[:]
            with torch.cuda.stream(stream_r):
                d_rays.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            launches0 = ctx.launch_count
            stream_r.synchronize()
            e0.record(stream_r)
            for st in (stream_a, stream_b, x_in, streams_r[0], streams_r[1]):
                st.wait_event(e0)            # nothing of the region starts before it
            run_pipelined(steps, e2e)
            for st in (stream_a, stream_b, x_out, streams_r[0], streams_r[1]):
                stream_r.wait_stream(st)
            e1.record(stream_r)
            barrier()
            for st in (stream_a, stream_b, x_in, x_out, streams_r[0], streams_r[1]):
                st.synchronize()
            ctx.switch_stream(stream_r.cuda_stream)
            if marks:  # RTR_BENCH_TRACE=1: when each phase of each frame ended, ms after the start of the timed region
                sys.stderr.write("rank %d: " % rank + "  ".join("%s%d@%.1f" % (w, f, e0.elapsed_time(e)) for w, f, e in marks) + "\n")
                del marks[:]
            return reduce_timing(e0.elapsed_time(e1), launches0)

        # weights from this box's own timings: rebuild + broadcast on rank 0, a full frame of rays on one GPU
        for _ in range(2):
            frame_device()
        ctx.gather_stripes(d_rgba.data_ptr(), W, H, 16, rpb, [1] * world, 0)   # first use connects the send/recv channels
        barrier()
        t0, t1, t2, t3, t4 = (torch.cuda.Event(enable_timing=True) for _ in range(5))
        t0.record(stream)
        if rank == 0:
            bvh.build_dev(d_tris.data_ptr(), n, n, d_meshes.data_ptr(), 1)
        t1.record(stream)   # the broadcast runs beside the rays: only the rebuild is serial on rank 0
        bvh.broadcast(0, traversal_only=not args.reference_order, expected_triangles=n)
        t2.record(stream)
        bvh.render_sharded_dev(cam, W, H, d_rgba.data_ptr(), rpb, 0, 1, bounces=bounces, flags=flags)
        t3.record(stream)
        ctx.gather_stripes(d_rgba.data_ptr(), W, H, 16, rpb, [1] * world, 0)
        t4.record(stream)
        barrier()
        tt = torch.tensor([t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3), t3.elapsed_time(t4)],
                          dtype=torch.float64, device=dev)
        dist.broadcast(tt, src=0)
        serial_ms, bcast_ms, render_ms, gather_ms = (float(x) for x in tt.tolist())
        layout = parallel.stripe_layout(world, parallel.builder_share_for(serial_ms, render_ms, world))
        # Rank 0 always keeps one stripe (at 8 GPUs the balance says 0.66 of one, a coin toss between 0 and 1): its rays are
        # what makes room for NCCL on the rebuilding rank -- the persistent kernel vacates --reserve-sms SMs, the rebuild's
        # kernels do not -- and with back-to-back rebuilds only the broadcasts starve (measured with the re-ordered
        # schedule: 2 169...2 305 Mrays/s without rows on rank 0, 2 694 with one stripe).
        layout[0] = max(1, layout[0])
        rank_zero_defers = 3 * layout[0] >= max(layout)   # 6 of 8 stripes at 2 GPUs, 4 (or 3) at 4: yes; 1 at 8: no
        phases = {"rebuild_ms": serial_ms, "broadcast_ms": bcast_ms, "full_frame_rays_ms_one_gpu": render_ms,
                  "gather_ms": gather_ms, "stripes_of_rank": layout,
                  "broadcast_held_until_next_rebuild_done": bool(rank_zero_defers and not os.environ.get("RTR_BENCH_NO_DEFER"))}
        barrier()
        if args.partition:
            # First launch inside the green context, with nothing else in flight: loading the kernel into a new context
            # synchronises the device, which must not happen later beside an NCCL kernel that waits for a peer.
            for sr in set(streams_r):
                ctx.switch_stream(sr.cuda_stream)
                bvh.render_stripes_dev(cam, W, H, d_rgba.data_ptr(), rpb, [1] * world, rank, bounces=0, flags=flags)
                sr.synchronize()
            ctx.switch_stream(stream.cuda_stream)
            barrier()

    def progress(what):
        if os.environ.get("RTR_BENCH_WATCHDOG"):
            sys.stderr.write("[rank %d] %s\n" % (rank, what)); sys.stderr.flush()

    if os.environ.get("RTR_BENCH_WATCHDOG") and pipelined:
        def report():   # which streams still hold work, which events have fired, shortly before the watchdog quits
            time.sleep(float(os.environ["RTR_BENCH_WATCHDOG"]) - 5.0)
            names = {"main": stream, "A(build)": stream_a, "B(comm)": stream_b, "x_in": x_in, "x_out": x_out,
                     "rays0": streams_r[0], "rays1": streams_r[1]}
            busy = [k for k, st in names.items() if not st.query()]
            evs = {"built": built, "ready": ready, "released": released, "sent": sent, "frame_done": x_frame_done, "img_done": x_img_done}
            def q(e):
                try:
                    return "1" if e.query() else "0"
                except Exception:
                    return "?"
            sys.stderr.write("[rank %d] busy streams: %s | events %s\n" % (
                rank, busy, {k: "".join(q(e) for e in v) for k, v in evs.items()}))
            sys.stderr.flush()
        threading.Thread(target=report, daemon=True).start()

    # ---- warm-up, then the timed region ----
    progress("setup done")
    sampler = ClockSampler(local_rank)
    if pipelined:
        run_pipelined(max(3, args.warmup), False)
        progress("warm-up submitted")
        if rank == 0:
            sampler.start()
        ms, rays, launches = timed_pipelined(args.steps, False)
        progress("timed region done")
    else:
        for _ in range(max(3, args.warmup)):
            frame_device()
        if rank == 0:
            sampler.start()
        ms, rays, launches = timed(frame_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = rays / (ms * 1e-3) / 1e6
    # the frame the timed region produced last (device-gathered on rank 0 when N > 1)
    hashes = {}
    if rank == 0:
        last_img = x_rgba[(args.steps - 1) % 2] if pipelined else d_rgba
        ctx.download(rgba_pinned.numpy(), last_img.data_ptr())
        hashes["value_frame"] = image_hash(rgba_pinned.numpy())

    # ---- end to end through the host-pointer ABI ----
    e2e_steps = args.steps
    e2e_how = "rtr_bvh_build (host triangles, synchronous upload) + frame + rtr_dev_download of the image, per step"
    if pipelined:
        e2e_how = ("pipelined frames, host traffic sharded: every rank uploads 1/%d of frame f+2's triangles over its own PCIe "
                   "link (rtr_dev_upload_async), rtr_gather_slices assembles them on rank 0 over NVLink, and every rank downloads "
                   "the rows it traced into one pinned host image shared by all ranks (rtr_download_stripes_async); the byte "
                   "counts are the totals over all ranks" % world)
        run_pipelined(2, True)
        progress("e2e warm-up submitted")
        e2e_ms, e2e_rays, _ = timed_pipelined(e2e_steps, True)
        progress("e2e region done")
    elif world == 1 and not args.no_pipeline:
        # Single GPU, double-buffered host traffic: the triangles of frame f+1 are uploaded (rtr_dev_upload_async,
        # pinned host memory, its own stream) while frame f is rebuilt and traced, and the image of frame f is
        # downloaded (rtr_dev_download_async) while frame f+1 is rebuilt.  Every step still moves its 640 MB in and
        # its 133 MB out inside the timed region.
        e2e_how = ("double-buffered: rtr_dev_upload_async of frame f+1's triangles and rtr_dev_download_async of frame "
                   "f-1's image run beside rtr_bvh_build_dev + rays of frame f (three streams, rtr_ctx_switch_stream)")
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            d_tris2 = [d_tris, torch.empty_like(d_tris)]
            d_rgba2 = [d_rgba, torch.empty_like(d_rgba)]
        rgba_pinned2 = [rgba_pinned, torch.empty_like(rgba_pinned).pin_memory()]
        meshes_pinned = torch.from_numpy(meshes_np.view(np.uint8).copy()).pin_memory()
        up_done, tris_free, frame_done, img_free = ([torch.cuda.Event(), torch.cuda.Event()] for _ in range(4))
        stream.synchronize()

        def e2e_upload(f):
            k = f % 2
            ctx.switch_stream(s_in.cuda_stream)
            s_in.wait_event(tris_free[k])
            ctx.upload_async(d_tris2[k].data_ptr(), tris_pinned.numpy())
            ctx.upload_async(d_meshes.data_ptr(), meshes_pinned.numpy())
            up_done[k].record(s_in)

        def e2e_frame(f):
            # (starting the upload of f+1 with the rays of f instead of beside its rebuild was measured: 30.42 against
            #  30.39 ms per step, profiles/exp_r02_19.sh -- the copy does not disturb the rebuild)
            k = f % 2
            ctx.switch_stream(stream.cuda_stream)
            stream.wait_event(up_done[k])
            bvh.build_dev(d_tris2[k].data_ptr(), n, n, d_meshes.data_ptr(), 1)
            stream.wait_event(img_free[k])      # only the rays need the image buffer back
            bvh.render_sharded_dev(cam, W, H, d_rgba2[k].data_ptr(), rpb, 0, 1, rays_dev=d_rays.data_ptr(),
                                   bounces=bounces, flags=flags)
            frame_done[k].record(stream)
            tris_free[k].record(stream)
            ctx.switch_stream(s_out.cuda_stream)
            s_out.wait_event(frame_done[k])
            ctx.download_async(rgba_pinned2[k].numpy(), d_rgba2[k].data_ptr())
            img_free[k].record(s_out)
            ctx.switch_stream(stream.cuda_stream)

        def run_e2e(steps):
            e2e_upload(0)
            for f in range(steps):
                if f + 1 < steps:
                    e2e_upload(f + 1)
                e2e_frame(f)
            stream.wait_stream(s_out)
            stream.wait_stream(s_in)

        run_e2e(2)
        barrier()
        with torch.cuda.stream(stream):
            d_rays.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = ctx.launch_count
        stream.synchronize()
        e0.record(stream)
        s_in.wait_event(e0)
        run_e2e(e2e_steps)
        e1.record(stream)
        barrier(); s_in.synchronize(); s_out.synchronize()
        e2e_ms, e2e_rays, _ = reduce_timing(e0.elapsed_time(e1), launches0)
        # the last image really is in host memory
        assert float(rgba_pinned2[(e2e_steps - 1) % 2][3]) == 1.0
    else:
        for _ in range(2):
            frame_e2e()
        e2e_ms, e2e_rays, _ = timed(frame_e2e, e2e_steps)
    e2e_value = e2e_rays / (e2e_ms * 1e-3) / 1e6
    # ---- correctness gate: the frame that came out of the (pipelined, sharded) loop is the frame one GPU renders ----
    image_check = None
    if rank == 0:
        if pipelined:
            hashes["e2e_frame"] = image_hash(np.asarray(x_host_img[(e2e_steps - 1) % 2]))
            ctx.switch_stream(stream.cuda_stream)
            bvh.build_dev(d_tris.data_ptr(), n, n, d_meshes.data_ptr(), 1)
            bvh.render_sharded_dev(cam, W, H, d_rgba.data_ptr(), rpb, 0, 1, bounces=bounces, flags=flags)
            ctx.download(rgba_pinned.numpy(), d_rgba.data_ptr())
            hashes["one_gpu_frame"] = image_hash(rgba_pinned.numpy())
        elif world == 1 and not args.no_pipeline:
            hashes["e2e_frame"] = image_hash(rgba_pinned2[(e2e_steps - 1) % 2].numpy())
        else:
            hashes["e2e_frame"] = image_hash(rgba_pinned.numpy())
        key = "%d_%dx%d_b%d" % (n, W, H, bounces)
        hash_file = os.path.join(ROOT, "profiles", "image_hashes.json")
        try:
            known = json.load(open(hash_file))
        except Exception:
            known = {}
        if args.record_hash and world == 1:
            known[key] = hashes["value_frame"]
            json.dump(known, open(hash_file, "w"), indent=1, sort_keys=True)
        distinct = sorted(set(hashes.values()) | ({known[key]} if key in known else set()))
        image_check = {"hash": hashes["value_frame"], "frames_compared": sorted(hashes), "expected_n1": known.get(key),
                       "ok": len(distinct) == 1,
                       "how": "blake2b-64 of the rgba32f frame: the timed loop's last frame (device-gathered), the e2e loop's last "
                              "frame (host image)" + (", rank 0 rendering the frame alone" if pipelined else "") +
                              " and the digest recorded at N=1 (profiles/image_hashes.json) must all be equal"}
        if not image_check["ok"]:
            raise RuntimeError("frames differ: %s (expected %s)" % (hashes, known.get(key)))
    # a ray that ran out of traversal stack would have dropped subtrees: such a run is not a measurement
    for k, used in enumerate(bvhs if pipelined else [bvh]):
        try:
            dropped = used.stack_overflows()
        except capi.RtrError:  # a BVH object this rank never filled
            continue
        if dropped:
            raise RuntimeError("rank %d, BVH %d: %d rays ran out of traversal stack" % (rank, k, dropped))
    h2d = n * TRIANGLE.itemsize + MESH.itemsize + 284
    d2h = H * W * 16

    extras = {}
    roofline = None
    if rank == 0 and not args.no_extras:
        # per-stage device times of the rebuild (events inside the library)
        bvh.enable_stage_timing(True)
        stage = np.zeros(6)
        reps = 5
        for _ in range(reps):
            bvh.build_dev(d_tris.data_ptr(), n, n, d_meshes.data_ptr(), 1)
            stage += bvh.stage_ms()
        stage /= reps
        bvh.enable_stage_timing(False)
        active, merges = bvh.iteration_trace()
        extras["build_ms"] = {"morton": stage[0], "sort": stage[1], "leaf_init": stage[2], "ploc_loop": stage[3],
                              "flatten": stage[4], "total": stage[5], "ploc_iterations": int(active.size),
                              "sum_active_over_n": float(active.sum()) / n}
        extras["sort_gkeys_s"] = n / (stage[1] * 1e-3) / 1e9
        # algorithmic bytes of the whole rebuild from this run's own iteration trace (SURVEY.md 8d formula)
        ploc_bytes = float((active.astype(np.float64) * 40).sum() + (merges.astype(np.float64) * 64).sum())
        build_bytes = n * (64 + 64 + 8) + n * 68.0 + n * (4 + 64 + 48) + ploc_bytes + (2 * n - 1) * 96.0
        # per-kernel times of one frame (CUDA events around every launch of the dominant kernels)
        ctx.profile_enable(True)
        if world == 1:
            frame_device()
        else:
            bvh.build_dev(d_tris.data_ptr(), n, n, d_meshes.data_ptr(), 1)
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        step_ms = ms / args.steps
        kern = {}
        for name, (tot, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            kern[name] = {"ms_total": tot, "launches": cnt, "share_of_step": tot / step_ms}
            if name in ALGO_BYTES:
                gbs = ALGO_BYTES[name](n) * cnt / (tot * 1e-3) / 1e9
                kern[name]["algorithmic_GBps"] = gbs
        if "ploc_loop_kernel" in kern:
            big = active[active > 1024].astype(np.float64)
            kern["ploc_loop_kernel"]["algorithmic_GBps"] = float(
                (big * 40).sum() + (merges[: big.size].astype(np.float64) * 64).sum()) / (kern["ploc_loop_kernel"]["ms_total"] * 1e-3) / 1e9
        extras["kernels"] = kern
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # dominant HBM-bound kernel of the rebuild
        hbm_kernels = [(k, v) for k, v in kern.items() if "algorithmic_GBps" in v]
        if hbm_kernels:
            name, v = max(hbm_kernels, key=lambda kv: kv[1]["ms_total"])
            traffic = traffic_note = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic, traffic_note = tj.get(name), tj.get(name + "_note")
            except Exception:
                pass
            roofline = {"bound": "hbm", "kernel": name, "achieved": v["algorithmic_GBps"], "peak": peak, "unit": "GB/s",
                        "frac": v["algorithmic_GBps"] / peak, "traffic": traffic, "traffic_note": traffic_note,
                        "peak_source": peak_src, "avg_launch_ms": v["ms_total"] / v["launches"]}
        extras["build_roofline"] = {"algorithmic_bytes": build_bytes, "achieved_GBps": build_bytes / (stage[5] * 1e-3) / 1e9,
                                    "frac_of_peak": build_bytes / (stage[5] * 1e-3) / 1e9 / peak}
        extras["sort_roofline"] = {"algorithmic_bytes": 68.0 * n, "achieved_GBps": 68.0 * n / (stage[1] * 1e-3) / 1e9,
                                   "frac_of_peak": 68.0 * n / (stage[1] * 1e-3) / 1e9 / peak}
        # traversal alone (BVH already built): the Mrays/s the rays see
        tr_ms, tr_rays, _ = timed(lambda: bvh.render_sharded_dev(cam, W, H, d_rgba.data_ptr(), rpb, 0, 1, rays_dev=d_rays.data_ptr(),
                                                                 bounces=bounces, flags=flags), 3) if world == 1 else (None, None, None)
        if tr_ms:
            extras["trace_only"] = {"ms_per_frame": tr_ms / 3, "mrays_s": tr_rays / (tr_ms * 1e-3) / 1e6,
                                    "rays_per_frame": tr_rays // 3}
            # warp execution efficiency and L2 hit rate of the traversal kernel (north_star): not measurable inside a
            # timed run, so they come from the committed ncu --set full capture of the same kernel on the same workload
            try:
                ev = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("trace_persistent_kernel_evidence")
                if ev and (n, W, H, bounces) == (10_000_000, 3840, 2160, 2):
                    extras["trace_only"].update({"warp_exec_eff": ev["warp_exec_eff"], "threads_per_inst": ev["threads_per_inst"],
                                                 "l2_hit_rate": ev["l2_hit_rate"], "evidence": ev["source"]})
            except Exception:
                pass

    cpu_baseline = cpu_stages = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cn, cw, ch = cpu_sample_sizes(args, 12.0)
        arm = CpuArm(args, cn, cw, ch)
        crays, csecs, cbuild = arm.step(split=True)
        cpu_baseline = {"value": crays / csecs / 1e6, "unit": "Mrays/s", "cores": arm.cores, "kind": arm.kind,
                        "sample": arm.describe(), "seconds": csecs,
                        "rebuild_seconds": cbuild, "traversal_mrays_s": crays / max(1e-9, csecs - cbuild) / 1e6,
                        "note": "a bounded sample (%.1f %% of the triangles, %.1f %% of the pixels of the GPU workload): unpruned CPU "
                                "Mrays/s falls as the scene grows, so value / this understates the like-for-like ratio"
                                % (100.0 * cn / n, 100.0 * cw * ch / (W * H))}
        cpu_stages = cpu_stage_baselines(args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n, W, H, bounces),
                       "triangles": n, "image": [W, H], "traced_pixels": [dw, dh], "bounces": bounces,
                       "rays_per_step": rays // args.steps, "search_radius": 16,
                       "trace_order": "reference" if args.reference_order else "pruned (identical records)",
                       "parallelism": ("build on rank 0 + NCCL broadcast + %d-row blocks dealt to %d rank(s)" % (rpb, world)) +
                                      ((", frames pipelined (rebuild of f+2 | broadcast of f+1 | rays of f, gathered on rank 0), stripes per rank %s, %d SMs left to NCCL%s"
                                        % (layout, args.reserve_sms, (", rays in a green-context partition of %d SMs, two frames in flight" % rays_sms) if rays_sms else "")) if pipelined else ""),
                       "l2": "inputs larger than L2 (640 MB of triangles + 960 MB of nodes per step), no flush needed"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "how": e2e_how},
            "gpu_launches": launches,
            "image_check": image_check,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "cpu_stage_baselines": cpu_stages,
        }
        if pipelined:
            line["multi_gpu"] = phases
        line.update(extras)
        print(json.dumps(line, default=float))

    if pipelined:
        for img in x_host_img:
            capi.host_unregister(img)
        del x_host_img
        dist.barrier()
        if rank == 0:
            for nm in shm_names:
                try:
                    os.unlink(nm)
                except OSError:
                    pass
        for other in bvhs[1:]:
            other.close()
    bvh.close()
    if comm:
        comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
