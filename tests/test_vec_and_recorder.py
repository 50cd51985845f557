"""FlatVecEnv (SB3-style facade) and TrajectoryRecorder driven by a deterministic CPU stand-in for the batched env, plus a
GPU smoke test on the real simulator.  Reference callers: examples/gym_manipulation_sb3.py:26-46 (make_vec_env + flattened
observations) and envs/wrappers/record_hdf5.py:20-151 (dataset names observations/qpos, observations/qvel, action)."""
import numpy as np
import pytest
import torch

from gym_lowcostrobot_b200.vec import FlatVecEnv, TrajectoryRecorder


class FakeEnv:
    """Env i terminates after 3 + i steps; obs[:, 0] = steps since reset, obs[:, 1] = env index, obs[:, 6] = episode number."""

    def __init__(self, n=4, obs_dim=15, action_dim=5):
        self.num_envs, self.obs_dim, self.action_dim, self.device = n, obs_dim, action_dim, torch.device("cpu")
        self.single_action_space, self.cfg = None, None
        self._obs = torch.zeros(n, obs_dim)
        self.t = torch.zeros(n, dtype=torch.long)
        self.ep = torch.zeros(n, dtype=torch.long)

    def _fill(self):
        self._obs[:, 0] = self.t.float()
        self._obs[:, 1] = torch.arange(self.num_envs).float()
        self._obs[:, 6] = self.ep.float()

    def reset(self, seed=None, options=None, mask=None):
        m = torch.ones(self.num_envs, dtype=torch.bool) if mask is None else mask.bool()
        self.t[m] = 0
        self.ep[m] += 1
        self._fill()
        return {"all": self._obs.clone()}, {}

    def step_flat(self, a):
        self.t += 1
        self._fill()
        done = self.t >= 3 + torch.arange(self.num_envs)
        return self._obs, -self.t.float(), done.to(torch.uint8), torch.zeros(self.num_envs, dtype=torch.uint8), done.to(torch.uint8)

    def close(self):
        pass


def test_flat_vec_env_resets_finished_envs_in_the_same_step():
    v = FlatVecEnv(FakeEnv())
    obs = v.reset()
    assert obs.shape == (4, 15) and torch.all(obs[:, 0] == 0)
    for t in range(1, 8):
        obs, rew, done, info = v.step(torch.zeros(4, 5))
        for i in range(4):
            ends = t % (3 + i) == 0 if t <= 3 + i else None
            if t == 3 + i:  # first episode of env i ends exactly here
                assert done[i] and obs[i, 0] == 0 and obs[i, 6] == 2  # already the first obs of episode 2
                k = info["done_index"].tolist().index(i)
                assert info["terminal_observation"][k, 0] == 3 + i and info["terminal_observation"][k, 6] == 1
            elif t < 3 + i:
                assert not done[i] and obs[i, 0] == t
        assert rew.shape == (4,) and info["is_success"].dtype == torch.bool


def test_flat_vec_env_numpy_mode_and_autoreset_guard():
    v = FlatVecEnv(FakeEnv(), to_numpy=True)
    assert isinstance(v.reset(), np.ndarray)
    obs, rew, done, info = v.step(np.zeros((4, 5), np.float32))
    assert isinstance(obs, np.ndarray) and done.dtype == np.bool_

    class Cfg:
        autoreset = 1
    e = FakeEnv()
    e.cfg = Cfg()
    with pytest.raises(ValueError):
        FlatVecEnv(e)


def test_trajectory_recorder_cuts_episodes_per_env(tmp_path):
    env = FakeEnv()
    v = FlatVecEnv(env)
    rec = TrajectoryRecorder(4, 5, horizon=10, device="cpu")
    v.reset()
    for t in range(1, 9):
        a = torch.full((4, 5), float(t))
        # record what the policy saw / did this step: the pre-reset observation of finished envs is the terminal one
        obs, rew, done, info = v.step(a)
        step_obs = obs.clone()
        if done.any():
            step_obs[info["done_index"]] = info["terminal_observation"]
        rec.record(step_obs, a, done)
    lens = sorted((ep["env"], len(ep["action"])) for ep in rec.episodes)
    assert lens == [(0, 3), (0, 3), (1, 4), (1, 4), (2, 5), (3, 6)]
    ep = next(e for e in rec.episodes if e["env"] == 2)
    assert ep["observations/qpos"].shape == (5, 6) and ep["observations/qvel"].shape == (5, 6)
    np.testing.assert_array_equal(ep["observations/qpos"][:, 0], np.arange(1, 6))  # steps since reset, as recorded
    np.testing.assert_array_equal(ep["action"][:, 0], np.arange(1, 6))
    assert rec.save(tmp_path) == 6
    z = np.load(tmp_path / "episode_0.npz")
    assert set(z.files) == {"observations__qpos", "observations__qvel", "action", "env"}


@pytest.mark.gpu
def test_facade_and_recorder_on_the_simulator():
    import gym_lowcostrobot_b200 as glr

    env = glr.make("PushCube-v0", num_envs=64, max_episode_steps=6)
    v = FlatVecEnv(env)
    rec = TrajectoryRecorder(64, env.action_dim, horizon=6, device="cuda:0")
    obs = v.reset(seed=0)
    assert obs.shape == (64, 18) and obs.is_cuda
    g = torch.Generator(device="cuda").manual_seed(0)
    age = torch.zeros(64, dtype=torch.long, device="cuda")
    for t in range(13):
        a = torch.rand(64, env.action_dim, generator=g, device="cuda") * 2 - 1
        obs, rew, done, info = v.step(a)
        age += 1
        assert done[age >= 6].all()  # TimeLimit(6); success may end an episode earlier
        assert torch.equal(info["TimeLimit.truncated"] | info["is_success"], done)
        step_obs = obs.clone()
        if done.any():
            step_obs[info["done_index"]] = info["terminal_observation"]
            assert torch.all(obs[done][:, 0:6] == 0)  # freshly reset arms (reference reset: qpos[:6] = 0)
        rec.record(step_obs, a, done)
        age[done] = 0
    assert rec.n_finished >= 128 and all(len(ep["action"]) <= 6 for ep in rec.episodes)
    v.close()
