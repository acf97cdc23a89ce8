"""Two real GPUs (skipped on a single-GPU box): the NCCL replica of a BVH -- full and traversal-only -- traces
bit-identically to the rank that built it, and weighted stripes gathered over NCCL give the plain frame."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes
    from realtimeraytracing_b200 import capi, parallel, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = capi.Context(rank)
        comm = parallel.RankComm(ctx)
        tris, meshes, L = scenes.soup(30000)
        W, H, rpb = 256, 208, 16
        cam = synth.soup_camera(L, W, H)
        layout = parallel.stripe_layout(world, 0.5, stripes_per_rank=4)
        expected = None
        for traversal_only in (False, True):
            bvh = capi.Bvh(ctx)
            if rank == 0:
                bvh.build(tris, meshes)
                if expected is None:
                    expected, _, _ = bvh.render(cam, W, H, W, H, bounces=2)
            bvh.broadcast(0, traversal_only=traversal_only, expected_triangles=tris.size if traversal_only else 0)
            d_rgba = ctx.dev_alloc(W * H * 16)
            ctx.zero(d_rgba, W * H * 16)
            bvh.render_stripes_dev(cam, W, H, d_rgba, rpb, layout, rank, denom_w=W, denom_h=H, bounces=2)
            ctx.allgather_stripes(d_rgba, W, H, 16, rpb, layout)
            got = np.zeros((H, W, 4), np.float32)
            ctx.download(got, d_rgba)
            ctx.dev_free(d_rgba)
            exp_t = torch.from_numpy(np.ascontiguousarray(expected.reshape(H, W, 4)) if rank == 0 else np.zeros((H, W, 4), np.float32)).cuda()
            dist.broadcast(exp_t, src=0)
            assert np.array_equal(got, exp_t.cpu().numpy()), "rank %d traversal_only=%s" % (rank, traversal_only)
            # the same frame gathered on rank 0 only (every block travels once)
            d_img = ctx.dev_alloc(W * H * 16)
            ctx.zero(d_img, W * H * 16)
            bvh.render_stripes_dev(cam, W, H, d_img, rpb, layout, rank, denom_w=W, denom_h=H, bounces=2)
            ctx.gather_stripes(d_img, W, H, 16, rpb, layout, 0)
            got0 = np.zeros((H, W, 4), np.float32)
            ctx.download(got0, d_img)
            ctx.dev_free(d_img)
            if rank == 0:
                assert np.array_equal(got0, exp_t.cpu().numpy()), "gather_stripes traversal_only=%s" % traversal_only
            if traversal_only and rank != 0:
                with pytest.raises(capi.RtrError):
                    bvh.trace_primary(cam, W, H, W, H, flags=capi.TRACE_REFERENCE_ORDER)
                with pytest.raises(capi.RtrError):
                    bvh.flat_nodes()
            bvh.close()
        comm.close()
        ctx.close()
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_replicas_trace_identically_on_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(2))
