"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) gloo process groups exercise exactly the partition
/ broadcast / gather code of realtimeraytracing_b200.parallel that bench.py runs over NCCL:
the build result is replicated by broadcast, image row blocks are dealt round-robin and all-gathered."""
import os
import socket
import sys

import numpy as np
import pytest

from realtimeraytracing_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_blocks_tile_the_image():
    for height, rpb in ((2160, 16), (1072, 16), (67, 16), (5, 8), (16, 16)):
        blocks = parallel.row_blocks(height, rpb)
        assert blocks[0][0] == 0 and blocks[-1][1] == height
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_every_row_has_exactly_one_owner(world):
    height, rpb = 2160, 16
    seen = np.zeros(height, dtype=np.int32)
    for rank in range(world):
        rows = parallel.rows_of_rank(height, rank, world, rpb)
        seen[rows] += 1
        assert parallel.pixels_of_rank(3840, height, rank, world, rpb) == rows.size * 3840
    assert (seen == 1).all()
    # same dealing as RowMap in csrc/trace.cu: block b -> rank b % world
    for b, (r0, _) in enumerate(parallel.row_blocks(height, rpb)):
        assert r0 in parallel.rows_of_rank(height, b % world, world, rpb)


def test_balance_is_within_one_block():
    for world in (2, 4, 8):
        sizes = [parallel.rows_of_rank(2160, r, world, 16).size for r in range(world)]
        assert max(sizes) - min(sizes) <= 16


def test_bad_arguments():
    with pytest.raises(ValueError):
        parallel.row_blocks(0, 16)
    with pytest.raises(ValueError):
        parallel.blocks_of_rank(64, 2, 2, 16)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, height, width, rpb, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from realtimeraytracing_b200 import parallel as par
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the build result exists on rank 0 only and is replicated by broadcast
        n_nodes = 2 * 1000 - 1
        flat = torch.zeros(n_nodes * 12, dtype=torch.int32)
        tris = torch.zeros(1000 * 16, dtype=torch.int32)
        if rank == 0:
            flat = torch.arange(n_nodes * 12, dtype=torch.int32)
            tris = torch.arange(1000 * 16, dtype=torch.int32) * 3
        par.broadcast_arrays([flat, tris], src=0)
        assert int(flat[-1]) == n_nodes * 12 - 1 and int(tris[5]) == 15
        # 2. every rank "renders" its own row blocks into a full-size image (value = f(row, col, owner))
        image = torch.full((height, width), -1.0)
        for r0, r1 in par.blocks_of_rank(height, rank, world, rpb):
            rows = torch.arange(r0, r1, dtype=torch.float32).unsqueeze(1)
            cols = torch.arange(width, dtype=torch.float32).unsqueeze(0)
            image[r0:r1] = rows * 1000.0 + cols + 0.25 * rank
        # 3. all-gather of the row blocks: afterwards every rank holds the whole frame
        par.gather_rows(image, height, rpb)
        rows = torch.arange(height, dtype=torch.float32).unsqueeze(1)
        cols = torch.arange(width, dtype=torch.float32).unsqueeze(0)
        owner = torch.tensor([par.owner_of_block(r // rpb, world) for r in range(height)], dtype=torch.float32).unsqueeze(1)
        assert torch.equal(image, rows * 1000.0 + cols + 0.25 * owner)
        # 4. whole-job ray count = sum over ranks (what bench.py all-reduces)
        mine = torch.tensor([par.pixels_of_rank(width, height, rank, world, rpb)], dtype=torch.int64)
        dist.all_reduce(mine)
        assert int(mine) == width * height
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 80), (2, 67), (3, 100)])
def test_broadcast_and_row_gather_over_gloo(tmp_path, world, height):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, height, 24, 16, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


# ---- weighted dealing (the rank that rebuilds the BVH renders fewer rows) --------------------------------
def _rowmap_rows(height, rpb, offsets, total):
    """The rows RowMap (csrc/trace.cu) enumerates for a rank owning the stripes `offsets` of `total`."""
    if not offsets:
        return []
    blocks = (height + rpb - 1) // rpb
    first, span = min(offsets), len(offsets)
    cycles = (blocks - first + total - 1) // total if blocks > first else 0
    rows = []
    for y_local in range(cycles * span * rpb):
        blk = y_local // rpb
        y = ((blk // span) * total + offsets[blk % span]) * rpb + (y_local % rpb)
        if y < height:
            rows.append(y)
    return rows


@pytest.mark.parametrize("world,share", [(1, 1.0), (2, 0.75), (2, 0.0), (4, 0.5), (8, 0.2), (8, 1.0)])
def test_stripe_layout_partitions_the_image(world, share):
    height, rpb = 2160, 16
    layout = parallel.stripe_layout(world, share)
    assert len(layout) == world and sum(layout) >= 1
    seen = np.zeros(height, dtype=np.int32)
    for rank in range(world):
        rows = [r for a, b in parallel.blocks_of_rank_striped(height, rank, layout, rpb) for r in range(a, b)]
        seen[rows] += 1
        # the device-side enumeration visits exactly these rows
        offsets = [v for v, r in enumerate(parallel.stripe_owners(layout)) if r == rank]
        assert sorted(_rowmap_rows(height, rpb, offsets, sum(layout))) == rows
    assert (seen == 1).all()
    if world > 1:
        sizes = [len(parallel.blocks_of_rank_striped(height, r, layout, rpb)) for r in range(world)]
        assert sizes[0] <= min(sizes[1:]) + 1  # the builder never renders more than the others
        assert max(sizes[1:]) - min(sizes[1:]) <= 1


def test_stripe_layout_with_full_share_is_plain_dealing():
    assert parallel.stripe_layout(4, 1.0, stripes_per_rank=1) == [1, 1, 1, 1]
    for b in range(40):
        assert parallel.owner_of_block_striped(b, [1, 1, 1, 1]) == parallel.owner_of_block(b, 4)


def test_builder_share():
    assert parallel.builder_share_for(5.0, 40.0, 1) == 1.0
    assert parallel.builder_share_for(5.0, 40.0, 8) == pytest.approx((45.0 / 8 - 5.0) / ((40.0 - (45.0 / 8 - 5.0)) / 7))
    assert parallel.builder_share_for(10.0, 20.0, 4) == 0.0      # the rebuild alone fills the builder's frame
    assert 0.0 < parallel.builder_share_for(4.4, 40.0, 2) < 1.0


def _striped_worker(rank, world, port, height, width, rpb, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from realtimeraytracing_b200 import parallel as par
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layout = par.stripe_layout(world, 0.5, stripes_per_rank=4)
        image = torch.full((height, width), -1.0)
        for r0, r1 in par.blocks_of_rank_striped(height, rank, layout, rpb):
            image[r0:r1] = float(rank)
        par.gather_rows_striped(image, height, layout, rpb)
        owner = torch.tensor([par.owner_of_block_striped(r // rpb, layout) for r in range(height)], dtype=torch.float32)
        assert torch.equal(image, owner.unsqueeze(1).expand(height, width))
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 200), (3, 333)])
def test_striped_gather_over_gloo(tmp_path, world, height):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_striped_worker, args=(world, port, height, 8, 16, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


# ---- sharded host traffic: every rank uploads a slice of the triangles and downloads its own rows ------------
def test_slices_partition_the_array():
    for n in (0, 1, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            cuts = [parallel.slice_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    with pytest.raises(ValueError):
        parallel.slice_range(10, 2, 2)


def _sharded_worker(rank, world, port, n, height, width, rpb, layout, out_dir):
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from realtimeraytracing_b200 import parallel as par
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # input: rank r holds only its slice of the frame's triangles; the build rank assembles them
        lo, hi = par.slice_range(n, rank, world)
        mine = torch.arange(lo, hi, dtype=torch.int64) * 7
        full = torch.full((n,), -1, dtype=torch.int64) if rank == 0 else None
        par.gather_slices(mine, full, n, root=0)
        if rank == 0:
            assert torch.equal(full, torch.arange(n, dtype=torch.int64) * 7)
        # output: every rank writes its own row blocks into one image all ranks share (a file mapping here, a pinned
        # host buffer in bench.py); nobody gathers
        shared = np.memmap(os.path.join(out_dir, "image"), dtype=np.float32, mode="r+", shape=(height, width))
        local = np.full((height, width), -1.0, np.float32)
        for r0, r1 in par.blocks_of_rank_striped(height, rank, layout, rpb):
            local[r0:r1] = np.arange(r0, r1, dtype=np.float32)[:, None] * 100 + rank
        par.write_stripes(shared, local, height, layout, rank, rpb)
        shared.flush()
        dist.barrier()
        if rank == 0:
            owner = np.array([par.owner_of_block_striped(r // rpb, layout) for r in range(height)], np.float32)
            exp = (np.arange(height, dtype=np.float32) * 100 + owner)[:, None] * np.ones((1, width), np.float32)
            assert np.array_equal(np.asarray(shared), exp)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,layout", [(2, [3, 8]), (3, [0, 4, 4])])
def test_sharded_upload_and_download_over_gloo(tmp_path, world, layout):
    import numpy as np
    import torch.multiprocessing as mp
    height, width = 100, 12
    np.memmap(str(tmp_path / "image"), dtype=np.float32, mode="w+", shape=(height, width)).flush()
    port = _free_port()
    mp.spawn(_sharded_worker, args=(world, port, 1001, height, width, 8, layout, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_stripe_layout_for_knows_what_a_launch_costs():
    """The builder's stripes minimise the modelled frame time; with a launch floor a single stripe is never worth a
    launch of its own at 8 ranks, and the 2 / 4 rank layouts keep the shares that balance rebuild and rays."""
    from realtimeraytracing_b200 import parallel
    assert parallel.stripe_layout_for(4.9, 37.9, 2.0, 1) == [1]
    assert parallel.stripe_layout_for(4.9, 35.8, 1.0, 2) == [6, 8]
    assert parallel.stripe_layout_for(4.9, 37.8, 1.0, 4)[0] in (3, 4) and parallel.stripe_layout_for(4.9, 37.8, 1.0, 4)[1:] == [8, 8, 8]
    assert parallel.stripe_layout_for(4.9, 37.9, 2.0, 8) == [0] + [8] * 7
    assert parallel.stripe_layout_for(0.0, 40.0, 0.0, 4) == [8, 8, 8, 8]           # nothing to rebuild: plain dealing
    assert parallel.stripe_layout_for(100.0, 40.0, 1.0, 2) == [0, 8]              # the rebuild dominates: the builder only builds
    for world in (2, 3, 4, 8):
        lay = parallel.stripe_layout_for(4.9, 37.9, 2.0, world)
        assert len(lay) == world and all(0 <= k <= 8 for k in lay) and sum(lay) > 0
        owners = parallel.stripe_owners(lay)
        assert sorted(set(owners)) == [r for r in range(world) if lay[r] > 0]


def _old_inline_loop(steps, e2e, defer):
    """the loop bench.py ran before the schedule moved into parallel.frame_schedule (the one the multi-GPU numbers of
    DESIGN.md §7 were measured with), transcribed: the two must stay identical"""
    ops = []
    if e2e:
        ops.append(("upload", 0))
        if steps > 1:
            ops.append(("upload", 1))
    ops.append(("build", 0))
    ops.append(("exchange", 0))
    if steps > 1:
        ops.append(("build", 1))
    for f in range(steps):
        if e2e and f + 2 < steps:
            ops.append(("upload", f + 2))
        if f + 1 < steps and not (defer and f + 2 < steps):
            ops.append(("exchange", f + 1))
        if f + 2 < steps:
            ops.append(("build", f + 2))
            if defer:
                ops.append(("exchange_after_build", f + 1))
        ops.append(("rays", f))
    return ops


@pytest.mark.parametrize("e2e", [False, True])
@pytest.mark.parametrize("hold", [False, True])
def test_frame_schedule_invariants(e2e, hold):
    """What the pipelined frames rely on, for every length: each operation once per frame; upload < build < exchange <
    rays within a frame; a receiver posts the exchange of f+1 before it launches the rays of f; and no buffer slot --
    three BVHs, two upload / triangle buffers -- is handed to a new frame before the operation that frees it has been
    issued (an event waited for before it is recorded again names the OLD record: the hazard the first re-ordering hit)."""
    from realtimeraytracing_b200 import parallel
    assert parallel.frame_schedule(0, e2e, hold) == []
    for steps in range(1, 14):
        ops = parallel.frame_schedule(steps, e2e, hold)
        assert ops == _old_inline_loop(steps, e2e, hold)
        at = {}
        for i, (op, f) in enumerate(ops):
            key = ("exchange" if op.startswith("exchange") else op, f)
            assert key not in at, "issued twice: %r" % (key,)
            at[key] = i
            assert 0 <= f < steps
        for f in range(steps):
            assert ("build", f) in at and ("exchange", f) in at and ("rays", f) in at
            assert (("upload", f) in at) == e2e
            if e2e:
                assert at[("upload", f)] < at[("build", f)]
            assert at[("build", f)] < at[("exchange", f)] < at[("rays", f)]
            if f + 1 < steps:
                assert at[("exchange", f + 1)] < at[("rays", f)]          # posted before the host turns to the rays
            if f + 3 < steps:                                             # BVH slot f % 3: rays and broadcast of f issued
                assert at[("build", f + 3)] > at[("rays", f)] and at[("build", f + 3)] > at[("exchange", f)]
                assert at[("exchange", f + 3)] > at[("rays", f)]          # receivers: the buffer is released by then
            if e2e and f + 2 < steps:                                     # triangle / slice slot f % 2: rebuilt from first
                assert at[("upload", f + 2)] > at[("build", f)]
            if f + 2 < steps:                                             # image slot f % 2
                assert at[("rays", f + 2)] > at[("rays", f)]
        for i, (op, f) in enumerate(ops):
            if op == "exchange_after_build":
                assert hold and ops[i - 1] == ("build", f + 1)
        if not hold:
            assert all(op != "exchange_after_build" for op, _ in ops)
