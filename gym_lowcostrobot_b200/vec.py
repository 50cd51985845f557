"""Callers on either side of the hot path, kept on the device:

* ``FlatVecEnv`` -- the vectorised-env facade the reference's training recipes build around the envs
  (``examples/gym_manipulation_sb3.py:26-46``: ``make_vec_env`` + ``FilterObservation`` / ``FlattenObservation``;
  ``examples/rl_zoo3_conf.yaml``): SB3 ``VecEnv`` call pattern (``reset`` / ``step_async`` / ``step_wait``, same-step
  auto-reset with ``terminal_observation``) over ONE batched simulator instead of a Python loop over envs, an optional
  observation-key filter, observations flattened to one ``[num_envs, obs_dim]`` float32 tensor that never leaves the GPU
  unless asked, and no host synchronisation per step (finished envs are reset by mask; the list of finished envs is only
  materialised when the caller reads it).
* ``make_sb3_vec_env`` -- the same thing as a real ``stable_baselines3.common.vec_env.VecEnv`` subclass (numpy in / out, list
  of info dicts), built when stable-baselines3 and gymnasium are importable.
* ``TrajectoryRecorder`` -- the batched replacement of ``RecordHDF5Wrapper`` (``envs/wrappers/record_hdf5.py:20-151``):
  per-env trajectories of ``observations/qpos``, ``observations/qvel`` and ``action`` (the wrapper's dataset names,
  ``record_hdf5.py:52-61``, minus images) are appended by one kernel of liblcrsim.so per step (``lcr_record_append``);
  finished episodes collect in a device pool that the host drains every few steps in one transfer, and are written as
  ``<prefix>-episode-<k>.hdf5`` through h5py when it is importable (``.npz`` with the same dataset names otherwise).

All of them only use the public env API (``reset(mask=...)``, ``step_flat``), so they work with any object that has it (the
CPU tests drive them with a deterministic stand-in; the GPU tests with the real simulator).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import spaces


def _flat(obs):
    return torch.cat([obs[k] for k in obs], 1) if isinstance(obs, dict) else obs


class _Infos(dict):
    """Per-step infos of ``FlatVecEnv``: device tensors; ``done_index`` / ``terminal_observation`` (rows of the finished
    envs only, like SB3's per-env ``terminal_observation``) are derived -- with the one host synchronisation that the
    variable row count needs -- only when they are read."""

    def __init__(self, base, done, last_obs, out):
        super().__init__(base)
        self._done, self._last, self._out = done, last_obs, out

    def __missing__(self, key):
        if key not in ("done_index", "terminal_observation"):
            raise KeyError(key)
        idx = self._done.nonzero(as_tuple=False).squeeze(1)
        self["done_index"] = self._out(idx)
        self["terminal_observation"] = self._out(self._last[idx])
        return self[key]

    def __contains__(self, key):
        return key in ("done_index", "terminal_observation") or super().__contains__(key)


class FlatVecEnv:
    """SB3-style vector env over a batched simulator.

    ``step_wait`` returns ``(obs [n, obs_dim], reward [n], done [n] bool, infos)``; envs whose episode ended in this
    step are reset immediately (reference envs behind SB3's ``DummyVecEnv`` behave the same way) and ``obs`` holds their
    first observation of the new episode.  ``infos["last_observation"]`` is the observation batch before those resets
    (``[n, obs_dim]``, meaningful where ``done``); ``infos["terminal_observation"]`` / ``infos["done_index"]`` are its rows
    for the finished envs and their indices, computed on first access.  ``infos["TimeLimit.truncated"]`` and
    ``infos["is_success"]`` are per-env bool tensors.  ``keys`` selects and orders observation keys like the
    ``FilterObservation`` + ``FlattenObservation`` pair of the reference's SB3 example (default: all keys in the env's order).
    With ``to_numpy=True`` everything is returned as numpy arrays (one device->host copy per step) for learners that need
    host data.
    """

    def __init__(self, env, to_numpy=False, keys=None):
        if getattr(env, "cfg", None) is not None and getattr(env.cfg, "autoreset", 0):
            raise ValueError("FlatVecEnv resets finished envs itself: create the env with autoreset=False")
        self.env, self.to_numpy = env, to_numpy
        self.num_envs = env.num_envs
        self.action_dim = env.action_dim
        self._cols = None
        if keys is not None:
            layout = getattr(env, "obs_layout", None)
            if layout is None:
                raise ValueError("the env has no obs_layout: observation keys cannot be filtered")
            off, k = {}, 0
            for name, width in layout:
                off[name] = (k, width)
                k += width
            missing = [name for name in keys if name not in off]
            if missing:
                raise KeyError(f"unknown observation keys {missing}; available: {list(off)}")
            cols = [c for name in keys for c in range(off[name][0], off[name][0] + off[name][1])]
            self._cols = torch.as_tensor(cols, dtype=torch.long, device=env.device)
        self.obs_dim = env.obs_dim if self._cols is None else int(self._cols.numel())
        self.observation_space = spaces.Box(-np.inf, np.inf, (self.obs_dim,), np.float32)
        self.action_space = env.single_action_space
        self._actions = None
        self._obs = None

    def _out(self, t):
        return t.detach().cpu().numpy() if self.to_numpy else t

    def _select(self, flat):
        return flat.clone() if self._cols is None else flat.index_select(1, self._cols)

    def reset(self, seed=None):
        obs, _ = self.env.reset(seed=seed)
        self._obs = self._select(_flat(obs))
        return self._out(self._obs)

    def step_async(self, actions):
        self._actions = torch.as_tensor(actions, device=self.env.device, dtype=torch.float32)

    def step_wait(self):
        obs, reward, te, tr, su = self.env.step_flat(self._actions)
        done = (te | tr).bool()
        last = self._select(obs)
        base = {"TimeLimit.truncated": tr.bool() & ~te.bool(), "is_success": su.bool().clone(), "last_observation": last}
        new_obs, _ = self.env.reset(mask=done)  # masked reset: nothing happens when no env is done, and nobody has to ask
        self._obs = self._select(_flat(new_obs))
        infos = _Infos({k: self._out(v) for k, v in base.items()}, done, last, self._out)
        return self._out(self._obs), self._out(reward.clone()), self._out(done), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()


def make_sb3_vec_env(env, keys=None):
    """``env`` behind a real ``stable_baselines3.common.vec_env.VecEnv`` (what ``make_vec_env`` returns in
    ``examples/gym_manipulation_sb3.py:34-35``): numpy observations ``[n, obs_dim]``, numpy actions, a list of per-env info
    dicts with ``terminal_observation`` / ``TimeLimit.truncated`` / ``is_success``.  Raises ``ImportError`` when
    stable-baselines3 or gymnasium are not installed -- ``FlatVecEnv`` is the dependency-free form of the same facade."""
    from gymnasium import spaces as gspaces
    from stable_baselines3.common.vec_env import VecEnv

    flat = FlatVecEnv(env, to_numpy=True, keys=keys)

    class LowCostRobotSB3VecEnv(VecEnv):
        def __init__(self):
            a = flat.action_space
            super().__init__(flat.num_envs, gspaces.Box(-np.inf, np.inf, (flat.obs_dim,), np.float32),
                             gspaces.Box(np.asarray(a.low, np.float32), np.asarray(a.high, np.float32), dtype=np.float32))
            self.flat, self._seed = flat, None

        def reset(self):
            seed, self._seed = self._seed, None
            return flat.reset(seed=seed)

        def step_async(self, actions):
            flat.step_async(np.asarray(actions, np.float32))

        def step_wait(self):
            obs, rew, done, info = flat.step_wait()
            infos = [{"TimeLimit.truncated": bool(info["TimeLimit.truncated"][i]), "is_success": bool(info["is_success"][i])}
                     for i in range(flat.num_envs)]
            for k, i in enumerate(np.asarray(info["done_index"]).tolist()):
                infos[i]["terminal_observation"] = info["terminal_observation"][k]
            return obs, rew, done, infos

        def close(self):
            flat.close()

        def seed(self, seed=None):
            self._seed = seed
            return [None if seed is None else seed + i for i in range(flat.num_envs)]

        def get_attr(self, attr_name, indices=None):
            return [getattr(flat.env, attr_name)] * len(self._get_indices(indices))

        def set_attr(self, attr_name, value, indices=None):
            setattr(flat.env, attr_name, value)

        def env_method(self, method_name, *args, indices=None, **kwargs):
            return [getattr(flat.env, method_name)(*args, **kwargs)] * len(self._get_indices(indices))

        def env_is_wrapped(self, wrapper_class, indices=None):
            return [False] * len(self._get_indices(indices))

    return LowCostRobotSB3VecEnv()


class TrajectoryRecorder:
    """Records ``observations/qpos``, ``observations/qvel`` (arm joints, as in the reference's observation dict) and
    ``action`` of every env on the device and hands finished episodes to the host.

    Call ``record(obs_flat, actions, done)`` once per step with the step's flat observation ``[n, obs_dim]`` (columns
    0:6 = arm_qpos, 6:12 = arm_qvel), the actions ``[n, A]`` and the done mask (or ``record(obs, actions, terminated,
    truncated)``).  On CUDA tensors the step is ONE launch of ``lcr_record_append`` (liblcrsim.so): rows go to the env's open
    trajectory, finished trajectories move to a device pool of ``pool_episodes`` slots.  The host drains the pool with one
    transfer every ``pool_episodes // num_envs`` steps -- the number of steps after which it could be full in the worst case,
    so nothing is ever dropped (``n_dropped`` stays 0) -- and on ``flush()`` / ``save()``; there is no per-step
    synchronisation.  ``episodes`` collects dicts with the wrapper's dataset names; ``save(dir)`` writes one file per
    episode.  (CPU tensors -- the host-logic tests -- take the same path in plain torch.)
    """

    def __init__(self, num_envs, action_dim, horizon=50, device="cuda:0", keep=True, pool_episodes=None):
        self.n, self.h, self.a, self.dev, self.keep = num_envs, horizon, action_dim, torch.device(device), keep
        self.w = 12 + action_dim
        self.cap = int(pool_episodes or 4 * num_envs)
        if self.cap < num_envs:
            raise ValueError("pool_episodes must be at least num_envs (every env can finish in the same step)")
        z = lambda *shape, dt=torch.float32: torch.zeros(*shape, dtype=dt, device=self.dev)
        self.traj, self.len = z(num_envs, horizon, self.w), z(num_envs, dt=torch.int32)
        self.pool, self.meta = z(self.cap, horizon, self.w), z(self.cap, 2, dt=torch.int32)
        self.count = z(2, dt=torch.int32)  # episodes in the pool, episodes dropped
        self._steps_since_flush = 0
        self._flush_every = max(1, self.cap // num_envs)
        self.episodes = []
        self.n_finished = self.n_dropped = 0
        self._L = None
        if self.dev.type == "cuda":
            from . import capi

            self._L, self._check = capi.lib(), capi.check

    def record(self, obs_flat, actions, done, truncated=None):
        if self._steps_since_flush >= self._flush_every:
            self.flush()
        self._steps_since_flush += 1
        te = done if done.dtype == torch.uint8 else done.to(torch.uint8)
        tr = torch.zeros_like(te) if truncated is None else (truncated if truncated.dtype == torch.uint8 else truncated.to(torch.uint8))
        if self._L is not None:
            obs_flat, actions = obs_flat.contiguous(), actions.to(torch.float32).contiguous()
            if obs_flat.dtype != torch.float32 or tuple(actions.shape) != (self.n, self.a) or obs_flat.shape[0] != self.n:
                raise ValueError("record: obs [n, obs_dim] float32 and actions [n, action_dim] expected")
            p = lambda t: C.c_void_p(t.data_ptr())
            with torch.cuda.device(self.dev):
                self._check(self._L.lcr_record_append(p(obs_flat), obs_flat.shape[1], p(actions), self.a, p(te.contiguous()), p(tr.contiguous()),
                                                      self.n, self.h, p(self.traj), p(self.len), p(self.pool), p(self.meta), p(self.count),
                                                      self.cap, C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)))
            return
        # the same step in torch (CPU tensors)
        rows = torch.arange(self.n, device=self.dev)
        t = self.len.clamp(max=self.h - 1).long()
        self.traj[rows, t] = torch.cat([obs_flat[:, :12].float(), actions.float()], 1)
        self.len += 1
        d = (te | tr).bool()
        slot = self.count[0] + torch.cumsum(d.to(torch.int32), 0) - 1
        ok = d & (slot < self.cap)
        idx = slot[ok].long()
        self.pool[idx] = self.traj[ok]
        self.meta[idx, 0] = rows[ok].to(torch.int32)
        self.meta[idx, 1] = self.len[ok].clamp(max=self.h)
        self.count[0] += d.sum().to(torch.int32)
        self.count[1] += (d & ~ok).sum().to(torch.int32)
        self.len[d] = 0

    def flush(self):
        """Drain the device pool: one synchronisation, one device->host copy of the filled slots."""
        self._steps_since_flush = 0
        c, dropped = (int(x) for x in self.count.cpu())
        k = min(c, self.cap)
        self.n_finished += c
        self.n_dropped += dropped
        if k and self.keep:
            pool, meta = self.pool[:k].cpu().numpy(), self.meta[:k].cpu().numpy()
            for s in range(k):
                L = int(meta[s, 1])
                self.episodes.append({"env": int(meta[s, 0]), "observations/qpos": pool[s, :L, 0:6].copy(),
                                      "observations/qvel": pool[s, :L, 6:12].copy(), "action": pool[s, :L, 12:].copy()})
        self.count.zero_()
        return k

    def save(self, directory, name_prefix="hdf5_record"):
        """One file per finished episode, named like the wrapper's (``record_hdf5.py:110``): HDF5 with the datasets
        ``observations/qpos``, ``observations/qvel``, ``action`` (``record_hdf5.py:52-61``) when h5py is importable,
        otherwise ``.npz`` holding the same dataset names.  Returns the number of episodes written."""
        self.flush()
        os.makedirs(directory, exist_ok=True)
        try:
            import h5py
        except ImportError:
            h5py = None
        for k, ep in enumerate(self.episodes):
            stem = os.path.join(directory, f"{name_prefix}-episode-{k}")
            data = {key: v for key, v in ep.items() if key != "env"}
            if h5py is not None:
                with h5py.File(stem + ".hdf5", "w") as f:
                    for key, v in data.items():
                        f.create_dataset(key, data=v)
                    f.attrs["env"] = ep["env"]
            else:
                np.savez_compressed(stem + ".npz", env=np.int64(ep["env"]), **data)
        return len(self.episodes)
