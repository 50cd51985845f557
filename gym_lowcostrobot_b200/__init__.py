"""gym_lowcostrobot_b200 -- B200-native batched simulator behind the gym_lowcostrobot env API.

``make("ReachCube-v0", num_envs=4096, observation_mode="state", action_mode="joint")`` mirrors
``gym.make`` on the IDs the reference registers (``gym_lowcostrobot/__init__.py:9-43``).  When gymnasium is importable the
same IDs are registered with it at import (``vector_entry_point`` = the batched env class, ``entry_point`` = a one-env
``gymnasium.Env`` adapter; see ``gymnasium_compat.register``).
"""
from .config import ENV_IDS, MAX_EPISODE_STEPS

__version__ = "0.2.0"


def make(env_id, **kwargs):
    if env_id not in ENV_IDS:
        raise KeyError(f"unknown env id {env_id!r}; available: {sorted(ENV_IDS)}")
    from .envs import ENV_CLASSES

    kwargs.setdefault("max_episode_steps", MAX_EPISODE_STEPS)
    return ENV_CLASSES[ENV_IDS[env_id]](**kwargs)


def register(namespace=None, force=False):
    """Register the env IDs with gymnasium (``gymnasium_compat.register``); raises ImportError without gymnasium."""
    from .gymnasium_compat import register as _register

    return _register(namespace=namespace, force=force)


try:  # the reference registers its IDs at import (gym_lowcostrobot/__init__.py:9-43); so does this package when it can
    import gymnasium as _gymnasium  # noqa: F401
except ImportError:
    pass
else:
    register()
