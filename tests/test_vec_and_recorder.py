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


def test_flat_vec_env_filters_keys_and_needs_no_sync_for_the_step():
    e = FakeEnv()
    e.obs_layout = (("arm_qpos", 6), ("arm_qvel", 6), ("cube_pos", 3))
    v = FlatVecEnv(e, keys=["cube_pos", "arm_qpos"])  # FilterObservation + FlattenObservation of examples/gym_manipulation_sb3.py:26-30
    assert v.obs_dim == 9 and v.observation_space.shape == (9,) and v.observation_space.contains(np.zeros(9, np.float32))
    obs = v.reset()
    assert obs.shape == (4, 9)
    for t in range(1, 4):
        obs, rew, done, info = v.step(torch.zeros(4, 5))
    assert done.tolist() == [True, False, False, False]
    assert obs[1, 3] == 3 and obs[1, 4] == 1  # columns 3.. are arm_qpos: steps since reset, env index
    assert info["last_observation"].shape == (4, 9) and info["last_observation"][0, 3] == 3  # before the reset of env 0
    assert "done_index" in info and info["done_index"].tolist() == [0] and info["terminal_observation"].shape == (1, 9)
    with pytest.raises(KeyError):
        FlatVecEnv(e, keys=["nope"])


def test_sb3_vec_env_adapter_with_stand_in_packages(monkeypatch):
    """make_sb3_vec_env builds a real VecEnv subclass when stable-baselines3 / gymnasium are importable; here minimal
    stand-ins for the two packages (neither is in this image) check the wiring: spaces, numpy batches, per-env info dicts."""
    import sys
    import types

    class VecEnv:
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

        def _get_indices(self, indices):
            return range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape if shape is not None else np.shape(low), dtype

    sb3, common, vec_env = types.ModuleType("stable_baselines3"), types.ModuleType("stable_baselines3.common"), types.ModuleType("stable_baselines3.common.vec_env")
    vec_env.VecEnv = VecEnv
    gymn, gspaces = types.ModuleType("gymnasium"), types.ModuleType("gymnasium.spaces")
    gspaces.Box = Box
    gymn.spaces = gspaces
    for name, mod in (("stable_baselines3", sb3), ("stable_baselines3.common", common), ("stable_baselines3.common.vec_env", vec_env),
                      ("gymnasium", gymn), ("gymnasium.spaces", gspaces)):
        monkeypatch.setitem(sys.modules, name, mod)
    from gym_lowcostrobot_b200 import spaces
    from gym_lowcostrobot_b200.vec import make_sb3_vec_env

    e = FakeEnv()
    e.single_action_space = spaces.Box(-1.0, 1.0, (5,), np.float32)
    v = make_sb3_vec_env(e)
    assert isinstance(v, VecEnv) and v.num_envs == 4 and v.observation_space.shape == (15,) and v.action_space.shape == (5,)
    assert v.seed(7) == [7, 8, 9, 10]
    obs = v.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (4, 15)
    for t in range(3):
        v.step_async(np.zeros((4, 5)))
        obs, rew, done, infos = v.step_wait()
    assert done.tolist() == [True, False, False, False] and len(infos) == 4
    assert infos[0]["terminal_observation"][0] == 3 and "terminal_observation" not in infos[1]
    assert infos[0]["is_success"] is True and infos[1]["TimeLimit.truncated"] is False
    assert v.get_attr("num_envs") == [4, 4, 4, 4] and v.env_is_wrapped(object) == [False] * 4


def _run_recorder(rec, steps=9):
    env = FakeEnv()
    v = FlatVecEnv(env)
    v.reset()
    for t in range(1, steps):
        a = torch.full((4, 5), float(t))
        # record what the policy saw / did this step: the pre-reset observation of finished envs is the terminal one
        obs, rew, done, info = v.step(a)
        rec.record(info["last_observation"], a, done)
    return rec


def test_trajectory_recorder_cuts_episodes_per_env(tmp_path):
    rec = _run_recorder(TrajectoryRecorder(4, 5, horizon=10, device="cpu", pool_episodes=8))  # pool of 8: drained every 2 steps
    assert len(rec.episodes) < 6  # the last finished episodes are still in the pool ...
    rec.flush()                   # ... until the host asks
    lens = sorted((ep["env"], len(ep["action"])) for ep in rec.episodes)
    assert lens == [(0, 3), (0, 3), (1, 4), (1, 4), (2, 5), (3, 6)] and rec.n_finished == 6 and rec.n_dropped == 0
    ep = next(e for e in rec.episodes if e["env"] == 2)
    assert ep["observations/qpos"].shape == (5, 6) and ep["observations/qvel"].shape == (5, 6)
    np.testing.assert_array_equal(ep["observations/qpos"][:, 0], np.arange(1, 6))  # steps since reset, as recorded
    np.testing.assert_array_equal(ep["action"][:, 0], np.arange(1, 6))
    assert rec.save(tmp_path) == 6
    z = np.load(tmp_path / "hdf5_record-episode-0.npz")  # (h5py is not in this image: same dataset names in an .npz)
    assert set(z.files) == {"observations/qpos", "observations/qvel", "action", "env"}


def test_trajectory_recorder_writes_hdf5_with_the_wrappers_dataset_names(tmp_path, monkeypatch):
    """With h5py importable the recorder writes <prefix>-episode-<k>.hdf5 holding observations/qpos, observations/qvel and
    action (record_hdf5.py:52-61,110); a stand-in h5py records the calls (h5py is not in this image)."""
    import sys
    import types

    written = {}

    class File:
        def __init__(self, path, mode):
            assert mode == "w"
            self.sets = written.setdefault(path, {})
            self.attrs = {}

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def create_dataset(self, name, data):
            self.sets[name] = np.asarray(data)

    h5 = types.ModuleType("h5py")
    h5.File = File
    monkeypatch.setitem(sys.modules, "h5py", h5)
    rec = _run_recorder(TrajectoryRecorder(4, 5, horizon=10, device="cpu"))
    assert rec.save(tmp_path, name_prefix="run") == 6
    assert sorted(written) == [str(tmp_path / f"run-episode-{k}.hdf5") for k in range(6)]
    for sets in written.values():
        assert set(sets) == {"observations/qpos", "observations/qvel", "action"}
        assert sets["observations/qpos"].shape[1] == 6 and sets["action"].shape == (sets["observations/qvel"].shape[0], 5)


def test_gymnasium_registration_with_a_stand_in_package(monkeypatch):
    """register() gives the reference's six IDs (gym_lowcostrobot/__init__.py:9-43) a vector_entry_point (the batched class) and a
    single-env entry_point, max_episode_steps=50; a stand-in gymnasium records the calls (gymnasium is not in this image)."""
    import sys
    import types

    calls = {}
    gymn, envs, reg = types.ModuleType("gymnasium"), types.ModuleType("gymnasium.envs"), types.ModuleType("gymnasium.envs.registration")
    reg.registry = {"PushCube-v0": "taken by the reference package"}
    reg.register = lambda **kw: calls.__setitem__(kw["id"], kw)
    envs.registration, gymn.envs = reg, envs
    for name, mod in (("gymnasium", gymn), ("gymnasium.envs", envs), ("gymnasium.envs.registration", reg)):
        monkeypatch.setitem(sys.modules, name, mod)
    import gym_lowcostrobot_b200 as glr
    from gym_lowcostrobot_b200 import envs as our_envs, gymnasium_compat

    ids = glr.register()
    assert sorted(ids) == sorted(["ReachCube-v0", "LiftCube-v0", "PickPlaceCube-v0", "StackTwoCubes-v0", "PushCubeLoop-v0"])  # PushCube-v0 left alone
    assert "PushCube-v0" in glr.register(force=True)
    for kw in calls.values():
        assert kw["max_episode_steps"] == 50
        mod, cls = kw["vector_entry_point"].split(":")
        assert mod == "gym_lowcostrobot_b200.envs" and issubclass(getattr(our_envs, cls), our_envs.BatchedLowCostRobotEnv)
        assert callable(getattr(gymnasium_compat, kw["entry_point"].split(":")[1]))
    assert glr.register(namespace="b200", force=True)[0].startswith("b200/")


@pytest.mark.gpu
def test_facade_and_recorder_on_the_simulator():
    import gym_lowcostrobot_b200 as glr

    env = glr.make("PushCube-v0", num_envs=64, max_episode_steps=6)
    v = FlatVecEnv(env)
    rec = TrajectoryRecorder(64, env.action_dim, horizon=6, device="cuda:0", pool_episodes=128)  # CUDA kernel (lcr_record_append)
    ref = TrajectoryRecorder(64, env.action_dim, horizon=6, device="cpu", pool_episodes=128)     # the same step in torch
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
        if done.any():
            assert torch.all(obs[done][:, 0:6] == 0)  # freshly reset arms (reference reset: qpos[:6] = 0)
            assert torch.equal(info["terminal_observation"], info["last_observation"][info["done_index"]])
        rec.record(info["last_observation"], a, done)
        ref.record(info["last_observation"].cpu(), a.cpu(), done.cpu())
        age[done] = 0
    rec.flush(), ref.flush()
    assert rec.n_finished >= 128 and rec.n_finished == ref.n_finished and rec.n_dropped == 0
    key = lambda ep: (ep["env"], len(ep["action"]), float(ep["action"].sum()))
    for x, y in zip(sorted(rec.episodes, key=key), sorted(ref.episodes, key=key)):
        assert x["env"] == y["env"] and all(np.array_equal(x[k], y[k]) for k in ("observations/qpos", "observations/qvel", "action"))
    assert all(len(ep["action"]) <= 6 for ep in rec.episodes)
    # filtered keys on the simulator: cube_pos | arm_qpos columns of the 18-wide observation
    v2 = FlatVecEnv(env, keys=["cube_pos", "arm_qpos"])
    o2 = v2.reset(seed=0)
    full = torch.cat([t for t in env.reset(seed=0)[0].values()], 1)
    assert o2.shape == (64, 9) and torch.equal(o2, torch.cat([full[:, 15:18], full[:, 0:6]], 1))
    v.close()
