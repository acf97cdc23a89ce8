"""One process per GPU: what is replicated and what is sharded.

* The build (Morton -> sort -> PLOC -> flatten) does NOT shard: PLOC iterations are globally
  dependent.  It runs on one rank; its RESULT (flat nodes + triangles + meshes) is broadcast.
* Rays are independent: image rows are dealt to ranks in blocks of ``rows_per_block``
  (block b -> rank b % world), every rank traces all bounces of its rows, the rows are gathered.

The reference has no multi-GPU path; this module is the host-side logic of the new
one.  Collectives go through ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests, which
exercise exactly this partition/gather logic) or through the C ABI's own NCCL calls
(rtr_comm_init / rtr_bvh_broadcast / rtr_allgather_rows) that ``RankComm`` bootstraps.
"""
from typing import List, Tuple

import numpy as np

DEFAULT_ROWS_PER_BLOCK = 16  # the reference's work-group height (raytracer.glsl:58)


def row_blocks(height: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK) -> List[Tuple[int, int]]:
    """[(row0, row1)] of every block of the image, in block order."""
    if height <= 0 or rows_per_block <= 0:
        raise ValueError("height and rows_per_block must be positive")
    return [(r, min(height, r + rows_per_block)) for r in range(0, height, rows_per_block)]


def owner_of_block(block: int, world: int) -> int:
    return block % world


def blocks_of_rank(height: int, rank: int, world: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK):
    """Row ranges rendered by `rank` (the same mapping as RowMap in csrc/trace.cu)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    return [rb for b, rb in enumerate(row_blocks(height, rows_per_block)) if owner_of_block(b, world) == rank]


def rows_of_rank(height: int, rank: int, world: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK) -> np.ndarray:
    blocks = blocks_of_rank(height, rank, world, rows_per_block)
    if not blocks:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate([np.arange(a, b) for a, b in blocks])


def pixels_of_rank(width: int, height: int, rank: int, world: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK) -> int:
    return int(rows_of_rank(height, rank, world, rows_per_block).size) * width


def slice_range(n_elems: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous slice of an n_elems array that `rank` uploads itself (rtr_gather_slices)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    return n_elems * rank // world, n_elems * (rank + 1) // world


def gather_slices(my_slice, full, n_elems: int, root: int = 0, group=None):
    """Host-side mirror of rtr_gather_slices: `full` (root only) receives every rank's slice at its place.
    Works on CPU tensors with gloo and CUDA tensors with NCCL; every slice travels once."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = slice_range(n_elems, rank, world)
    if rank == root:
        full[lo:hi] = my_slice
        works = []
        for r in range(world):
            a, b = slice_range(n_elems, r, world)
            if r != root and b > a:
                works.append(dist.irecv(full[a:b], src=r, group=group))
        for w in works:
            w.wait()
    elif hi > lo:
        dist.send(my_slice, dst=root, group=group)
    return full


def write_stripes(shared_image, local_image, height: int, layout: List[int], rank: int,
                  rows_per_block: int = DEFAULT_ROWS_PER_BLOCK):
    """Host-side mirror of rtr_download_stripes_async: this rank's row blocks go to the same rows of an image all
    ranks share (no collective)."""
    for r0, r1 in blocks_of_rank_striped(height, rank, layout, rows_per_block):
        shared_image[r0:r1] = local_image[r0:r1]
    return shared_image


# ---- weighted dealing: the rank that rebuilds the BVH renders fewer rows ------------------------------------
def stripe_layout(world: int, builder_share: float = 1.0, stripes_per_rank: int = 8) -> List[int]:
    """Stripes owned by each rank.  Block b of the image belongs to stripe b % sum(layout); rank r owns
    layout[r] of them (see stripe_owners).  Every rank but the builder (rank 0) gets
    `stripes_per_rank`; the builder gets round(builder_share * stripes_per_rank), possibly 0 --
    builder_share is the fraction of a full share of rendering it can take next to the rebuild
    (1.0: plain dealing, the layout of rtr_render_sharded_dev with stripes_per_rank == 1)."""
    if world <= 0 or stripes_per_rank <= 0:
        raise ValueError("world and stripes_per_rank must be positive")
    if world == 1:
        return [1]
    share = min(1.0, max(0.0, float(builder_share)))
    return [int(round(share * stripes_per_rank))] + [stripes_per_rank] * (world - 1)


def stripe_owners(layout: List[int]) -> List[int]:
    """Owner of every stripe of a cycle (rtr_stripe_owners in csrc/common.cuh): round k hands one stripe to every
    rank with more than k stripes, in rank order, so the ranks' blocks interleave."""
    owners = []
    for k in range(max(layout)):
        owners += [r for r, n in enumerate(layout) if n > k]
    return owners


def owner_of_block_striped(block: int, layout: List[int]) -> int:
    owners = stripe_owners(layout)
    return owners[block % len(owners)]


def blocks_of_rank_striped(height: int, rank: int, layout: List[int], rows_per_block: int = DEFAULT_ROWS_PER_BLOCK):
    """Row ranges rendered by `rank` under a stripe layout (the mapping of RowMap with span > 1 in csrc/trace.cu)."""
    return [rb for b, rb in enumerate(row_blocks(height, rows_per_block)) if owner_of_block_striped(b, layout) == rank]


def builder_share_for(build_ms: float, render_ms: float, world: int) -> float:
    """Share of a full rendering slice the building rank can take so that all ranks finish together:
    with per-frame work B (rebuild, rank 0 only) and R (rendering, divisible), every rank should be busy
    (B + R) / world; rank 0 renders (B + R) / world - B of it, a full share is what the others render."""
    if world <= 1:
        return 1.0
    t = (build_ms + render_ms) / world
    mine = t - build_ms
    if mine <= 0.0:
        return 0.0
    others = (render_ms - mine) / (world - 1)
    return max(0.0, min(1.0, mine / others))


def stripe_layout_for(build_ms: float, render_ms: float, launch_ms: float, world: int, stripes_per_rank: int = 8) -> List[int]:
    """stripe_layout with the builder's stripes chosen by a model that knows what a launch costs whatever it traces
    (`launch_ms`: the frame's longest paths, walked by a few lanes): a rank with k of V stripes takes
    launch_ms + k / V * (render_ms - launch_ms), the builder `build_ms` on top -- and nothing at all with k == 0.  The
    k with the shortest frame wins (ties: the larger k).  At 8 GPUs that is 0: one stripe would cost the builder a
    whole launch."""
    if world <= 1:
        return [1]
    bulk = max(0.0, render_ms - launch_ms)
    best_k, best_t = 0, float("inf")
    for k in range(stripes_per_rank + 1):
        total = k + stripes_per_rank * (world - 1)
        builder = build_ms + ((launch_ms + k / total * bulk) if k else 0.0)
        worker = launch_ms + stripes_per_rank / total * bulk
        t = max(builder, worker)
        if t <= best_t + 1e-9:
            best_k, best_t = k, t
    return [best_k] + [stripes_per_rank] * (world - 1)


def frame_schedule(steps: int, e2e: bool, hold_broadcast: bool) -> List[Tuple[str, int]]:
    """Issue order of the pipelined frames of `bench.py` (N > 1), the same on every rank -- which is what keeps the NCCL
    calls of the ranks in step.  Operations: ("upload", f) this rank's slice of frame f's triangles goes up and the slices
    are assembled on rank 0 (e2e arm only), ("build", f) rank 0 rebuilds BVH f, ("exchange", f) BVH f travels from rank 0,
    ("exchange_after_build", f) the same, held back on rank 0 until the rebuild issued just before it (f + 1) is through,
    ("rays", f) every rank traces its stripes of frame f (and the frame is gathered / downloaded).

    Steady state, per frame f:  upload(f+2) | exchange(f+1) | build(f+2) | rays(f)   -- or, with hold_broadcast (ranks 0
    that trace a good part of the frame themselves), upload(f+2) | build(f+2) | exchange_after_build(f+1) | rays(f): the
    broadcast then travels beside rank 0's rays of f instead of beside its rebuild of f+2.  The buffers behind it:
    three BVHs (f % 3), two upload / triangle / image buffers (f % 2); tests/test_parallel_cpu.py checks that no slot is
    re-used before the operation that frees it has been issued."""
    ops: List[Tuple[str, int]] = []
    if steps <= 0:
        return ops
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
        held = hold_broadcast and f + 2 < steps
        if f + 1 < steps and not held:
            ops.append(("exchange", f + 1))
        if f + 2 < steps:
            ops.append(("build", f + 2))
            if held:
                ops.append(("exchange_after_build", f + 1))
        ops.append(("rays", f))
    return ops


def gather_rows_striped(image, height: int, layout: List[int], rows_per_block: int = DEFAULT_ROWS_PER_BLOCK, group=None):
    """gather_rows for a stripe layout (same schedule as rtr_allgather_stripes)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return image
    works = []
    for b, (r0, r1) in enumerate(row_blocks(height, rows_per_block)):
        src = owner_of_block_striped(b, layout)
        works.append(dist.broadcast(image[r0:r1], src=dist.get_global_rank(group, src) if group else src, group=group, async_op=True))
    for w in works:
        w.wait()
    return image


def gather_rows(image, height: int, rows_per_block: int = DEFAULT_ROWS_PER_BLOCK, group=None):
    """All-gather the row blocks of a full-size image tensor [height, ...] in place: after the call
    every rank holds every block.  Each block is broadcast by its owner (same schedule as
    rtr_allgather_rows); works on CPU tensors with gloo and CUDA tensors with NCCL."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return image
    works = []
    for b, (r0, r1) in enumerate(row_blocks(height, rows_per_block)):
        works.append(dist.broadcast(image[r0:r1], src=dist.get_global_rank(group, owner_of_block(b, world)) if group else owner_of_block(b, world),
                                    group=group, async_op=True))
    for w in works:
        w.wait()
    return image


def broadcast_arrays(arrays, src: int = 0, group=None):
    """Broadcast a list of equally-shaped-per-rank tensors (the flat BVH, triangles, meshes)."""
    import torch.distributed as dist
    for a in arrays:
        dist.broadcast(a, src=src, group=group)
    return arrays


class RankComm:
    """Bootstraps the C ABI's NCCL communicator from an initialised torch.distributed group: rank 0
    creates the NCCL unique id, torch.distributed carries it to the other ranks."""

    def __init__(self, ctx, group=None):
        import torch.distributed as dist
        self.ctx = ctx
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 1:
            box = [ctx.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            ctx.comm_init(box[0], self.rank, self.world)

    def close(self):
        if self.world > 1:
            self.ctx.comm_destroy()
