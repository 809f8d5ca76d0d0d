"""world_size-2 gloo tests of the multi-GPU host logic (no GPU): the frame-parallel pose partition and
the column-split + gather-to-rank-0 path, with the oracle standing in for the per-rank renderer."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import SKY, ROOT


def test_pose_shard_partitions_everything():
    from futspace_b200.shard import pose_shard
    for n in (1, 7, 512, 513):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, count = pose_shard(n, r, world)
                seen += list(range(first, first + count))
            assert seen == list(range(n))
            sizes = [pose_shard(n, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_pose_interleave_partitions_everything():
    from futspace_b200.shard import pose_interleave
    for n in (1, 7, 512, 2048):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in pose_interleave(n, r, world))
            assert seen == list(range(n))
    assert pose_interleave(8, 1, 4) == [1, 5]


def test_column_bounds_aligned_and_complete():
    from futspace_b200.shard import column_bounds
    for w in (1, 31, 32, 1000, 1920, 3840, 7680):
        for world in (1, 2, 4, 8):
            b = column_bounds(w, world)
            assert b[0] == 0 and b[-1] == w and len(b) == world + 1
            assert all(b[i] <= b[i + 1] for i in range(world))
            assert all(x % 32 == 0 for x in b[1:-1])
    assert column_bounds(7680, 8) == [960 * i for i in range(9)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import futspace_b200 as F
    import oracle_lib as O
    from futspace_b200.shard import pose_shard, column_bounds, gather_columns
    col, hgt = F.terrain_fbm(256)
    prm = O.default_params()
    h, w = 96, 200
    # --- column split: every rank renders its slab (oracle renders the frame, we keep the slab), gather to rank 0
    cam = O.Camera(100.3, 77.7, 200, 2.2, 30, 150, 1.2, SKY)
    full = O.render(cam, prm, col, hgt, h, w, nthreads=1)
    b = column_bounds(w, world)
    wmax = max(b[i + 1] - b[i] for i in range(world))
    slab = np.zeros((h, wmax), np.uint32)
    slab[:, : b[rank + 1] - b[rank]] = full[:, b[rank]: b[rank + 1]]
    frame = gather_columns(dist, torch.from_numpy(slab.view(np.int32)), b, h, w, rank, world)
    if rank == 0:
        assert np.array_equal(frame.numpy().view(np.uint32), full)
    else:
        assert frame is None
    # --- frame-parallel: each rank renders its block of the path; a checksum of checksums must match a serial run
    n = 5
    first, count = pose_shard(n, rank, world)
    cams = [O.Camera(100.3 + 3 * i, 77.7, 200, 2.2 + 0.1 * i, 30, 100, 1.2, SKY) for i in range(n)]
    local = sum(int(O.render(cams[i], prm, col, hgt, 48, 64, nthreads=1).astype(np.uint64).sum()) for i in range(first, first + count))
    t = torch.tensor([local], dtype=torch.int64)
    dist.all_reduce(t)
    serial = sum(int(O.render(c, prm, col, hgt, 48, 64, nthreads=1).astype(np.uint64).sum()) for c in cams)
    assert int(t.item()) == serial
    # timing is reported as the max over ranks
    tm = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    assert tm.item() == float(world)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")


def test_two_rank_gloo(tmp_path):
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
