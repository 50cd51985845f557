"""Callers on either side of the hot path, kept on the device:

* ``FlatVecEnv`` -- the vectorised-env facade the reference's training recipes build around the envs
  (``examples/gym_manipulation_sb3.py:26-46``: ``make_vec_env`` + ``FilterObservation`` / ``FlattenObservation``;
  ``examples/rl_zoo3_conf.yaml``): SB3 ``VecEnv`` call pattern (``reset`` / ``step_async`` / ``step_wait``, same-step
  auto-reset with ``terminal_observation``) over ONE batched simulator instead of a Python loop over envs, observations
  flattened to one ``[num_envs, obs_dim]`` float32 tensor that never leaves the GPU unless asked.
* ``TrajectoryRecorder`` -- the batched replacement of ``RecordHDF5Wrapper`` (``envs/wrappers/record_hdf5.py:20-151``):
  per-env episode buffers of ``observations/qpos``, ``observations/qvel`` and ``action`` (the wrapper's dataset names,
  ``record_hdf5.py:52-61``, minus images) live in a device ring; finished episodes are copied to the host in one
  transfer per step and written as ``.npz`` files (h5py is not available in this image).

Both only use the public env API (``reset(mask=...)``, ``step_flat``), so they work with any object that has it (the CPU
tests drive them with a deterministic stand-in; the GPU test with the real simulator).
"""
from __future__ import annotations

import os

import numpy as np
import torch


class FlatVecEnv:
    """SB3-style vector env over a batched simulator.

    ``step_wait`` returns ``(obs [n, obs_dim], reward [n], done [n] bool, infos)``; envs whose episode ended in this
    step are reset immediately (reference envs behind SB3's ``DummyVecEnv`` behave the same way) and ``obs`` holds their
    first observation of the new episode, the last observation of the finished one is in
    ``infos["terminal_observation"]`` (rows of the envs in ``infos["done_index"]``).  ``infos["TimeLimit.truncated"]``
    and ``infos["is_success"]`` are per-env bool tensors.  With ``to_numpy=True`` everything is returned as numpy arrays
    (one device->host copy per step) for learners that need host data.
    """

    def __init__(self, env, to_numpy=False):
        if getattr(env, "cfg", None) is not None and getattr(env.cfg, "autoreset", 0):
            raise ValueError("FlatVecEnv resets finished envs itself: create the env with autoreset=False")
        self.env, self.to_numpy = env, to_numpy
        self.num_envs = env.num_envs
        self.obs_dim, self.action_dim = env.obs_dim, env.action_dim
        self.observation_space = ("Box", (self.obs_dim,), "float32")
        self.action_space = env.single_action_space
        self._actions = None
        self._obs = None

    def _out(self, t):
        return t.detach().cpu().numpy() if self.to_numpy else t

    def reset(self, seed=None):
        self.env.reset(seed=seed)
        self._obs = self.env._obs.clone() if hasattr(self.env, "_obs") else self.env.flat_obs().clone()
        return self._out(self._obs)

    def step_async(self, actions):
        self._actions = torch.as_tensor(actions, device=self.env.device, dtype=torch.float32)

    def step_wait(self):
        obs, reward, te, tr, su = self.env.step_flat(self._actions)
        done = (te | tr).bool()
        obs = obs.clone()
        infos = {"TimeLimit.truncated": tr.bool() & ~te.bool(), "is_success": su.bool().clone()}
        if bool(done.any()):
            idx = done.nonzero(as_tuple=False).squeeze(1)
            infos["done_index"] = self._out(idx)
            infos["terminal_observation"] = self._out(obs[idx].clone())
            new_obs, _ = self.env.reset(mask=done)
            flat = torch.cat([new_obs[k] for k in new_obs], 1) if isinstance(new_obs, dict) else new_obs
            obs[idx] = flat[idx]
        self._obs = obs
        return self._out(obs), self._out(reward.clone()), self._out(done), {k: (self._out(v) if torch.is_tensor(v) else v) for k, v in infos.items()}

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()


class TrajectoryRecorder:
    """Records ``observations/qpos``, ``observations/qvel`` (arm joints, as in the reference's observation dict) and
    ``action`` of every env into a device buffer ``[num_envs, horizon, ...]`` and hands finished episodes to the host.

    Call ``record(obs_flat, actions, done)`` once per step with the step's flat observation ``[n, obs_dim]`` (columns
    0:6 = arm_qpos, 6:12 = arm_qvel), actions ``[n, A]`` and the done mask.  Episodes that ended are gathered on the
    device and copied out with ONE device->host transfer; ``episodes`` collects them as dicts, and ``save(dir)`` writes
    one ``episode_<k>.npz`` per episode with the wrapper's dataset names.
    """

    def __init__(self, num_envs, action_dim, horizon=50, device="cuda:0", keep=True):
        self.n, self.h, self.dev, self.keep = num_envs, horizon, torch.device(device), keep
        self.qpos = torch.zeros(num_envs, horizon, 6, dtype=torch.float32, device=self.dev)
        self.qvel = torch.zeros(num_envs, horizon, 6, dtype=torch.float32, device=self.dev)
        self.act = torch.zeros(num_envs, horizon, action_dim, dtype=torch.float32, device=self.dev)
        self.len = torch.zeros(num_envs, dtype=torch.long, device=self.dev)
        self._rows = torch.arange(num_envs, device=self.dev)
        self.episodes = []
        self.n_finished = 0

    def record(self, obs_flat, actions, done):
        t = self.len.clamp(max=self.h - 1)
        self.qpos[self._rows, t] = obs_flat[:, 0:6]
        self.qvel[self._rows, t] = obs_flat[:, 6:12]
        self.act[self._rows, t] = actions.to(torch.float32)
        self.len += 1
        done = done.bool()
        if bool(done.any()):
            idx = done.nonzero(as_tuple=False).squeeze(1)
            L = self.len[idx].clamp(max=self.h)
            pack = torch.cat([self.qpos[idx].flatten(1), self.qvel[idx].flatten(1), self.act[idx].flatten(1), L[:, None].float(),
                              idx[:, None].float()], 1).cpu().numpy()  # one D2H copy for all finished episodes
            self.n_finished += len(idx)
            if self.keep:
                a = self.act.shape[2]
                for row in pack:
                    n = int(row[-2])
                    qp = row[:6 * self.h].reshape(self.h, 6)[:n]
                    qv = row[6 * self.h:12 * self.h].reshape(self.h, 6)[:n]
                    ac = row[12 * self.h:12 * self.h + a * self.h].reshape(self.h, a)[:n]
                    self.episodes.append({"env": int(row[-1]), "observations/qpos": qp.copy(), "observations/qvel": qv.copy(), "action": ac.copy()})
            self.len[idx] = 0

    def save(self, directory):
        os.makedirs(directory, exist_ok=True)
        for k, ep in enumerate(self.episodes):
            np.savez_compressed(os.path.join(directory, f"episode_{k}.npz"), **{key.replace("/", "__"): v for key, v in ep.items() if key != "env"},
                                env=np.int64(ep["env"]))
        return len(self.episodes)
