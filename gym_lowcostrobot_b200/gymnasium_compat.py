"""gymnasium registration of the env IDs the reference registers (``gym_lowcostrobot/__init__.py:9-43``), available when
gymnasium is importable (it is not a dependency of the simulator).

``register()`` gives every ID

* a ``vector_entry_point`` -- the batched env class itself, so that ``gymnasium.make_vec("ReachCube-v0", num_envs=4096,
  vectorization_mode="vector_entry_point", observation_mode="state", action_mode="joint")`` returns the B200 simulator, and
* an ``entry_point`` -- ``SingleEnv``: one env (``num_envs=1``) behind the ``gymnasium.Env`` interface with numpy
  observations, so that ``gymnasium.make`` on the same ID (what ``examples/*.py`` and ``tests/test_env.py:9-12`` do) keeps
  working; gymnasium's own ``TimeLimit(max_episode_steps=50)`` then applies as in the reference.
"""
from __future__ import annotations

import numpy as np

from .config import ENV_IDS, MAX_EPISODE_STEPS

_CLASS = {"reach": "ReachCubeEnv", "push": "PushCubeEnv", "lift": "LiftCubeEnv", "pick_place": "PickPlaceCubeEnv",
          "stack": "StackTwoCubesEnv", "push_loop": "PushCubeLoopEnv"}


def register(namespace=None, force=False):
    """Register the six env IDs with gymnasium; returns the list of IDs registered.  IDs that are already in the registry
    (the reference package imported first) are left alone unless ``force``; ``namespace="b200"`` registers
    ``b200/ReachCube-v0`` ... beside them instead."""
    import gymnasium
    from gymnasium.envs.registration import register as gym_register

    done = []
    for env_id, task in ENV_IDS.items():
        full = f"{namespace}/{env_id}" if namespace else env_id
        if full in getattr(gymnasium.envs.registration, "registry", {}) and not force:
            continue
        gym_register(id=full, entry_point=f"gym_lowcostrobot_b200.gymnasium_compat:make_single_{task}",
                     vector_entry_point=f"gym_lowcostrobot_b200.envs:{_CLASS[task]}", max_episode_steps=MAX_EPISODE_STEPS)
        done.append(full)
    return done


def _single_env_class():
    import gymnasium
    from gymnasium import spaces as gspaces

    class SingleEnv(gymnasium.Env):
        """One simulated env behind ``gymnasium.Env``: numpy in / out, the reference's observation dict and action space."""

        metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 25}

        def __init__(self, task, **kwargs):
            from .envs import ENV_CLASSES

            kwargs.setdefault("max_episode_steps", 0)  # gymnasium's TimeLimit wrapper truncates, like in the reference
            self.batched = ENV_CLASSES[task](num_envs=1, **kwargs)
            a, o = self.batched.single_action_space, self.batched.single_observation_space
            self.action_space = gspaces.Box(np.asarray(a.low, np.float32), np.asarray(a.high, np.float32), dtype=np.float32)
            self.observation_space = gspaces.Dict({k: gspaces.Box(np.asarray(s.low, np.float32), np.asarray(s.high, np.float32), dtype=np.float32)
                                                   for k, s in o.spaces.items()})

        @staticmethod
        def _np(obs):
            return {k: v[0].cpu().numpy() for k, v in obs.items()}

        def reset(self, seed=None, options=None):
            super().reset(seed=seed)
            obs, info = self.batched.reset(seed=seed, options=options)
            return self._np(obs), info

        def step(self, action):
            a = np.asarray(action, np.float32)
            if a.shape != self.action_space.shape:
                raise ValueError("Action dimension mismatch")  # reach_cube_env.py:231-232
            obs, reward, te, tr, info = self.batched.step(a[None])
            info = {k: (v[0].item() if hasattr(v, "shape") and getattr(v, "ndim", 0) >= 1 else v) for k, v in info.items()}
            return self._np(obs), float(reward[0]), bool(te[0]), bool(tr[0]), info

        def close(self):
            self.batched.close()

    return SingleEnv


def _make_single(task):
    def make(**kwargs):
        return _single_env_class()(task, **kwargs)

    make.__name__ = f"make_single_{task}"
    return make


for _task in _CLASS:
    globals()[f"make_single_{_task}"] = _make_single(_task)
