"""lcr_ik (the batched IK helper of the C-ABI, reference inverse_kinematics reach_cube_env.py:148-221 called on its own)
against the oracle; the helper must not modify the simulation.  The in-step IK that shares the device function is covered
by tests/test_gpu_parity.py (ee rollouts) and tests/test_reference_glue.py (reference fixtures).  Green on B200:
profiles/r01l_ik_helper_gpu.log."""
import numpy as np
import pytest
import torch

import gym_lowcostrobot_b200 as glr
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("float64", 2e-6), ("float32", 2e-3)])
def test_ik_helper_matches_oracle_and_leaves_the_state_untouched(precision, tol):
    n = 32
    rng = np.random.default_rng(17)
    lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
    hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])
    env = glr.make("ReachCube-v0", num_envs=n, precision=precision)
    env.reset(seed=0)
    qpos = env.get_state()["qpos"].cpu().numpy()
    qpos[:, :6] = rng.uniform(0.5 * lo, 0.5 * hi, size=(n, 6))
    env.set_state(qpos=qpos)
    env.substeps(0)
    before = {k: v.clone() for k, v in env.get_state().items()}
    site = before["aux"][:, 4:7].cpu().numpy()
    target = (site + rng.uniform(-0.06, 0.06, size=(n, 3))).astype(np.float32)
    q = env.inverse_kinematics(torch.from_numpy(target).cuda()).cpu().numpy()
    after = env.get_state()
    for k in before:
        assert torch.equal(before[k], after[k]), k  # the helper does not modify the simulation
    for i in range(n):
        o = Oracle("reach")
        o.reset(seed=i)
        o.set_state(qpos=qpos[i])
        o.forward()
        np.testing.assert_allclose(q[i], o.ik(target[i]), rtol=0, atol=tol)
    env.close()
