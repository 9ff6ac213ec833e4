"""Multi-GPU host logic on CPU: contiguous env shards and the rollout-statistics all-gather
(the only collective on this path), exercised with gloo at world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from robot_gym.controllers.mpc.batched_mpc_controller import gather_rollout_stats, reduce_rollout_stats, shard_bounds
from robot_gym.model.robots.descriptions import GHOST
from robot_gym.util import synthetic


def test_shard_bounds_partition_exactly():
    for n, w in ((1 << 20, 8), (4096, 3), (10, 4), (3, 8)):
        bounds = [shard_bounds(n, r, w) for r in range(w)]
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        assert all(bounds[r][1] == bounds[r + 1][0] for r in range(w - 1))
        sizes = [hi - lo for lo, hi in bounds]
        assert max(sizes) - min(sizes) <= 1


def test_sharded_inputs_are_slices_of_the_global_batch():
    full = synthetic.make_states(1000, GHOST)
    lo, hi = shard_bounds(1000, 1, 4)
    part = full.slice(lo, hi)
    again = synthetic.make_states(1000, GHOST).slice(lo, hi)
    for name in ("base_rpy", "foot_positions_base", "planned_contacts", "command", "time_since_reset"):
        np.testing.assert_array_equal(getattr(part, name), getattr(again, name))
        np.testing.assert_array_equal(getattr(part, name), getattr(full, name)[lo:hi])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(1001, rank, world)
    local = torch.tensor([hi - lo, 5.0 * (hi - lo), 7.0 + rank, 2.0 * (hi - lo), hi - lo, 0.0, 0.0, 100.0 * (rank + 1)],
                         dtype=torch.float64)
    gathered = gather_rollout_stats(local)
    total = reduce_rollout_stats(gathered)
    torch.save((gathered, total), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_rollout_stats_all_gather_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0, t0 = torch.load(tmp_path / "r0.pt")
    g1, t1 = torch.load(tmp_path / "r1.pt")
    assert torch.equal(g0, g1) and g0.shape == (2, 8)
    assert t0[0].item() == 1001 and t0[2].item() == 8.0 and t0[7].item() == 300.0
    assert torch.equal(t0, t1)


def test_gather_without_process_group_is_identity():
    v = torch.arange(8, dtype=torch.float64)
    assert torch.equal(gather_rollout_stats(v), v.unsqueeze(0))
