"""Multi-GPU sharding: envs are independent, so rank r owns the contiguous block
[r*n_local, (r+1)*n_local) and the only exchange is ONE all-gather per step of the packed output
record (obs | reward | terminated | truncated | success, float32) so that every rank (the learner)
sees the whole batch.  Actions need no collective: each rank slices its rows.

The reference has no parallelism at all (SB3 DummyVecEnv in examples/gym_manipulation_sb3.py:34-35
steps envs sequentially); this is the data-parallel path the north star adds.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total, world_size, rank):
    """Contiguous block of env indices owned by ``rank`` (n_total must divide evenly)."""
    if n_total % world_size:
        raise ValueError("num_envs must be divisible by world_size")
    n = n_total // world_size
    return rank * n, (rank + 1) * n


def pack_record(obs, reward, terminated, truncated, success, out=None):
    """[n, O+4] float32 record: obs, reward, terminated, truncated, success."""
    n, o = obs.shape
    if out is None:
        out = torch.empty(n, o + 4, dtype=torch.float32, device=obs.device)
    out[:, :o] = obs
    out[:, o] = reward
    out[:, o + 1] = terminated
    out[:, o + 2] = truncated
    out[:, o + 3] = success
    return out


def unpack_record(rec):
    o = rec.shape[1] - 4
    return rec[:, :o], rec[:, o], rec[:, o + 1] > 0.5, rec[:, o + 2] > 0.5, rec[:, o + 3] > 0.5


class ShardedEnv:
    """Wraps a local env (anything with ``step_flat(actions_local)``) into a world-wide batch."""

    def __init__(self, local_env, n_total, world_size=None, rank=None, group=None):
        self.env = local_env
        self.group = group
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.rank = dist.get_rank(group) if rank is None else rank
        self.lo, self.hi = shard_range(n_total, self.world_size, self.rank)
        self.n_total, self.n_local = n_total, self.hi - self.lo
        self._local = None
        self._full = None

    def step(self, actions_full):
        """actions_full: [n_total, A] (every rank holds the full action batch, e.g. from a replicated
        policy).  Returns the gathered (obs, reward, terminated, truncated, success) of all envs."""
        a = actions_full[self.lo:self.hi]
        if hasattr(self.env, "step_packed"):  # CUDA env: one fused pack kernel writes the send buffer
            self._local = self.env.step_packed(a)
        else:
            self._local = pack_record(*self.env.step_flat(a), out=self._local)
        if self.world_size == 1:
            return unpack_record(self._local)
        if self._full is None:
            self._full = torch.empty(self.n_total, self._local.shape[1], dtype=torch.float32, device=self._local.device)
        nvtx = getattr(self.env, "nvtx", False)
        if nvtx:
            torch.cuda.nvtx.range_push(f"lcr_all_gather[{self.n_total} x {self._local.shape[1]} f32]")
        dist.all_gather_into_tensor(self._full, self._local, group=self.group)
        if nvtx:
            torch.cuda.nvtx.range_pop()
        return unpack_record(self._full)
