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
