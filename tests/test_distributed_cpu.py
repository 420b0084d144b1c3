"""Host-side logic of the frame-sharded multi-GPU path on CPU: world_size 2, gloo backend.

The data path has exactly two exchange points (multi-GPU): the all-gather of the packed 2-D
points that feeds the replicated bundle adjustment (frame axis = dim 1 of (7, T, 38, 2)) and ONE
all-gather of the 3-D joints (frame axis = dim 0).  Here the sharding arithmetic and both gathers
(frame order preserved rank-major) run with gloo; the same code runs with NCCL on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepfly3d_b200.pipeline import gather_frames, reorder_calib, shard_frames


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_frames(T, rank, world)
    full = torch.arange(T * 38 * 3, dtype=torch.float64).reshape(T, 38, 3)
    local = full[lo:hi].clone()
    gathered = gather_frames(local)
    # the 2-D points of every rank's frames, gathered along the frame axis of (7, T, 38, 2)
    full2d = torch.arange(7 * T * 38 * 2, dtype=torch.float64).reshape(7, T, 38, 2)
    gathered2d = gather_frames(full2d[:, lo:hi].contiguous(), dim=1)
    torch.save({"gathered": gathered, "gathered2d": gathered2d, "range": (lo, hi)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_shard_and_gather_two_ranks(tmp_path):
    T, world = 64, 2
    mp.spawn(_worker, args=(world, _free_port(), T, str(tmp_path)), nprocs=world, join=True)
    full = torch.arange(T * 38 * 3, dtype=torch.float64).reshape(T, 38, 3)
    ranges = []
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(d["gathered"], full)          # frame order preserved on every rank
        assert torch.equal(d["gathered2d"], torch.arange(7 * T * 38 * 2, dtype=torch.float64).reshape(7, T, 38, 2))
        ranges.append(d["range"])
    assert ranges == [(0, 32), (32, 64)]


@pytest.mark.parametrize("T,world", [(100000, 8), (15, 2), (7, 8), (0, 4)])
def test_shard_frames_covers_everything_once(T, world):
    seen = np.zeros(T, dtype=int)
    for r in range(world):
        lo, hi = shard_frames(T, r, world)
        assert 0 <= lo <= hi <= T
        seen[lo:hi] += 1
    assert np.all(seen == 1)
    if T == 100000:
        assert shard_frames(T, 0, world) == (0, 12500)     # 12 500 frames per rank


def test_reorder_calib_matches_reference_rule(golden):
    """calib_reordered[cidx] = calib[idx] for (idx, cidx) in enumerate(camera_ordering) (core.py:240-242)."""
    calib = {k: golden["calib"][k] for k in ("R", "tvec", "intr", "distort")}
    order = [6, 5, 4, 3, 2, 1, 0]
    out = reorder_calib(calib, order)
    for idx, cidx in enumerate(order):
        assert np.array_equal(out["R"][cidx], calib["R"][idx])
    assert np.array_equal(reorder_calib(calib, range(7))["tvec"], calib["tvec"])
