"""Host-side partitioning for the two multi-GPU modes.  Pure Python, no GPU needed.

* frame-parallel: a camera path of n poses is cut into contiguous blocks, one per rank; frames are
  independent (the only cross-sample dependence of fut/voxel_renderer.fut:229-250 is along z inside one
  column), maps are replicated, there is no data-path collective.
* column-split: one frame is cut into column slabs aligned to the 32-column expand tile; every rank
  needs the whole map (rays fan out), and the one exchange is "slabs -> rank 0".
"""


def pose_shard(n_total, rank, world):
    """-> (first, count): contiguous block of the camera path rendered by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def pose_interleave(n_total, rank, world):
    """-> list of pose indices for `rank` under round-robin sharding (pose i -> rank i mod world).  Every rank's poses
    span the whole camera path, so the per-GPU work of a weak-scaling run does not depend on the rank."""
    return list(range(rank, n_total, world))


def column_bounds(w, world, align=32):
    """-> world+1 boundaries: slab r is [b[r], b[r+1]); interior boundaries are multiples of `align`."""
    units = (w + align - 1) // align
    b = [0]
    for r in range(1, world):
        b.append(min(w, ((units * r) // world) * align))
    b.append(w)
    return b


def gather_columns(dist, slab, bounds, h, w, rank, world, dst=0):
    """Collective 'slabs -> rank dst' through torch.distributed (NCCL on GPUs, gloo on CPU).

    slab: this rank's [h][wmax] tensor (int32 view of the u32 pixels), wmax = widest slab, columns beyond
    the rank's own width are padding.  Returns the assembled [h][w] frame on rank dst, None elsewhere."""
    import torch
    wmax = max(bounds[i + 1] - bounds[i] for i in range(world))
    assert tuple(slab.shape) == (h, wmax)
    parts = [torch.empty_like(slab) for _ in range(world)] if rank == dst else None
    dist.gather(slab, parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat([parts[r][:, : bounds[r + 1] - bounds[r]] for r in range(world)], dim=1)
