"""gym_lowcostrobot_b200 -- B200-native batched simulator behind the gym_lowcostrobot env API.

``make("ReachCube-v0", num_envs=4096, observation_mode="state", action_mode="joint")`` mirrors
``gym.make`` on the IDs the reference registers (``gym_lowcostrobot/__init__.py:9-37``).
"""
from .config import ENV_IDS, MAX_EPISODE_STEPS

__version__ = "0.1.0"


def make(env_id, **kwargs):
    if env_id not in ENV_IDS:
        raise KeyError(f"unknown env id {env_id!r}; available: {sorted(ENV_IDS)}")
    from .envs import ENV_CLASSES

    kwargs.setdefault("max_episode_steps", MAX_EPISODE_STEPS)
    return ENV_CLASSES[ENV_IDS[env_id]](**kwargs)
