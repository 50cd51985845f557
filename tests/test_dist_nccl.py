"""dist.ShardedEnv over NCCL on two real GPUs (skipped on a one-GPU box): every rank steps its shard of the simulator and the
all-gathered batch must equal the single-device batch of the same global seeds, bit for bit -- the path `bench.py --gpus N`
times.  Reference: the reference has no parallelism (examples/gym_manipulation_sb3.py:34-35 steps envs one after another)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_total, steps, ret):
    import gym_lowcostrobot_b200 as glr
    from gym_lowcostrobot_b200.dist import ShardedEnv

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n = n_total // world
    env = glr.make("PushCube-v0", num_envs=n, device=f"cuda:{rank}", autoreset=True, env_offset=rank * n, max_episode_steps=4)
    env.nvtx = True
    env.reset(seed=11)
    sh = ShardedEnv(env, n_total, world, rank)
    full = glr.make("PushCube-v0", num_envs=n_total, device=f"cuda:{rank}", autoreset=True, max_episode_steps=4) if rank == 0 else None
    if full is not None:
        full.reset(seed=11)
    g = torch.Generator(device=dev).manual_seed(3)  # the same action batch on every rank (a replicated policy)
    ok = True
    for t in range(steps):
        a = torch.rand(n_total, env.action_dim, generator=g, device=dev) * 2 - 1
        obs, reward, te, tr, su = sh.step(a)
        if full is not None:
            o2, r2, te2, tr2, su2 = full.step_flat(a)
            ok &= torch.equal(obs, o2) and torch.equal(reward, r2) and torch.equal(te, te2.bool()) and torch.equal(tr, tr2.bool()) and torch.equal(su, su2.bool())
    ret[rank] = bool(ok) and obs.shape == (n_total, env.obs_dim)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_env_over_nccl_reproduces_the_single_device_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world, n_total, steps = 2, 2048, 9
    port = 29500 + (os.getpid() % 2000)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_total, steps, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
