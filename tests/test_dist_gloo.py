"""world_size-2 gloo test of the multi-GPU host logic (sharding + the one all-gather per step) on CPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_lowcostrobot_b200.dist import ShardedEnv, pack_record, shard_range, unpack_record


class FakeEnv:
    """Deterministic stand-in for a local env: outputs are functions of the GLOBAL env index and the action."""

    def __init__(self, lo, hi, obs_dim=15):
        self.idx = torch.arange(lo, hi, dtype=torch.float32)
        self.obs_dim = obs_dim

    def step_flat(self, a):
        obs = self.idx[:, None] + torch.arange(self.obs_dim, dtype=torch.float32)[None] * 0.01 + a.sum(1, keepdim=True)
        return obs, -self.idx, (self.idx % 2 == 0).to(torch.uint8), (self.idx % 3 == 0).to(torch.uint8), (self.idx % 5 == 0).to(torch.uint8)


def _worker(rank, world, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, world, rank)
    sh = ShardedEnv(FakeEnv(lo, hi), n_total, world, rank)
    torch.manual_seed(0)
    actions = torch.rand(n_total, 5)
    out = sh.step(actions)
    ref = FakeEnv(0, n_total).step_flat(actions)
    ok = all(torch.equal(a.float(), b.float()) for a, b in zip(out, ref))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_range_and_record_roundtrip():
    assert shard_range(16, 4, 2) == (8, 12)
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)
    e = FakeEnv(0, 6)
    out = e.step_flat(torch.zeros(6, 5))
    rec = pack_record(*out)
    assert rec.shape == (6, 19)
    back = unpack_record(rec)
    assert torch.equal(back[0], out[0]) and torch.equal(back[2], out[2].bool()) and torch.equal(back[4], out[4].bool())


def test_two_rank_all_gather_reassembles_the_full_batch():
    world, n_total = 2, 16
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
