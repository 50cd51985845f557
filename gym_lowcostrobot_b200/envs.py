"""Batched drop-in for ``gym_lowcostrobot.envs`` (state observations, joint / ee actions).

Each class mirrors the constructor kwargs, ``reset`` / ``step`` / ``close`` interface, observation
keys and action shapes of its reference counterpart
(``gym_lowcostrobot/envs/{reach,push,lift,pick_place}_cube_env.py``, ``stack_two_cubes_env.py``,
``push_cube_loop_env.py``) but
holds ``num_envs`` instances on one GPU and returns batched torch tensors.  Everything behind
``step`` runs in liblcrsim.so (hand-written sm_100a CUDA); torch only owns the buffers and streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi, config, model, spaces

_OBS_LAYOUT = {
    "reach": (("arm_qpos", 6), ("arm_qvel", 6), ("cube_pos", 3)),
    "lift": (("arm_qpos", 6), ("arm_qvel", 6), ("cube_pos", 3)),
    "push": (("arm_qpos", 6), ("arm_qvel", 6), ("target_pos", 3), ("cube_pos", 3)),
    "pick_place": (("arm_qpos", 6), ("arm_qvel", 6), ("target_pos", 3), ("cube_pos", 3)),
    "stack": (("arm_qpos", 6), ("arm_qvel", 6), ("cube_red_pos", 3), ("cube_blue_pos", 3)),
    "push_loop": (("arm_qpos", 6), ("arm_qvel", 6), ("cube_pos", 3)),  # push_cube_loop_env.py:286-300
}


_EXEC_MODES = {"fused": 0, "phased": 1, "lockstep": 2, "flow": 3}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def pcg64_states(seeds):
    """[n,4] uint64 (state_hi, state_lo, inc_hi, inc_lo) of numpy ``PCG64(seed)`` for every seed.

    gymnasium's ``Env.reset(seed=s)`` sets ``np_random = Generator(PCG64(SeedSequence(s)))``; the
    device continues exactly that stream (reference reach_cube_env.py:299-302).
    """
    out = np.empty((len(seeds), 4), dtype=np.uint64)
    m64 = (1 << 64) - 1
    for i, s in enumerate(seeds):
        st = np.random.PCG64(int(s)).state["state"]
        out[i] = (st["state"] >> 64, st["state"] & m64, st["inc"] >> 64, st["inc"] & m64)
    return out


class BatchedLowCostRobotEnv:
    """``num_envs`` independent copies of one reference env, stepped in lockstep on one GPU."""

    task = None
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 25}

    def __init__(self, num_envs=1, device="cuda:0", observation_mode="state", action_mode="joint", reward_type="sparse",
                 block_gripper=None, distance_threshold=0.05, height_threshold=0.1, cube_xy_range=0.3,
                 target_xy_range=0.3, goal_z_range=0.1, n_substeps=20, render_mode=None, max_episode_steps=50,
                 autoreset=False, precision="float32", assets_path=None, collision_mask=model.COLLIDE_ALL,
                 env_offset=0, exec_mode="auto", seed=None, report_failures=False):
        if observation_mode not in ("state", "image", "both"):
            raise ValueError("observation_mode must be 'image', 'state' or 'both'")
        if render_mode is not None:
            raise NotImplementedError("render_mode (interactive viewer / 640x640 rgb_array of camera_vizu) is out of scope")
        if not torch.cuda.is_available():
            raise capi.LcrError("CUDA device required: the simulator has no CPU fallback")
        self.num_envs = int(num_envs)
        self.report_failures = bool(report_failures)
        self.device = torch.device(device)
        self.observation_mode, self.action_mode, self.reward_type = observation_mode, action_mode, reward_type
        self.cfg = config.make_cfg(self.task, action_mode=action_mode, reward_type=reward_type, block_gripper=block_gripper,
                                   distance_threshold=distance_threshold, height_threshold=height_threshold,
                                   cube_xy_range=cube_xy_range, target_xy_range=target_xy_range, goal_z_range=goal_z_range,
                                   n_substeps=n_substeps, max_episode_steps=max_episode_steps, autoreset=autoreset,
                                   collision_mask=collision_mask, exec_mode=_EXEC_MODES[self._pick_exec_mode(exec_mode, int(num_envs))])
        self.exec_mode = {v: k for k, v in _EXEC_MODES.items()}[self.cfg.exec_mode]
        self.block_gripper = bool(self.cfg.block_gripper)
        self.compiled = model.load_compiled(self.task, assets_path)
        self.cmodel, self.verts = model.pack_model(self.compiled)
        self.nq, self.nv = self.cmodel.nq, self.cmodel.nv
        self.action_dim, self.obs_dim = config.action_dim(self.cfg), config.obs_dim(self.task)
        self.precision = {"float32": capi.F32, "float64": capi.F64}[precision]
        self.env_offset = int(env_offset)  # global index of local env 0 (multi-GPU sharding)
        self._L = capi.lib()
        h = C.c_void_p()
        if self.device.index is None:  # "cuda" = the current device, which need not be 0
            self.device = torch.device("cuda", torch.cuda.current_device())
        capi.check(self._L.lcr_create(C.byref(self.cmodel), self.verts.ctypes.data_as(C.c_void_p), C.byref(self.cfg),
                                      self.num_envs, self.device.index, self.precision, C.byref(h)))
        self._h = h
        n, dev = self.num_envs, self.device
        self._obs = torch.zeros(n, self.obs_dim, dtype=torch.float32, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flags = torch.zeros(3, n, dtype=torch.uint8, device=dev)
        self._record = None
        self.single_action_space = spaces.Box(-1.0, 1.0, (self.action_dim,), np.float32)
        sub = {"arm_qpos": spaces.Box(-np.pi, np.pi, (6,)), "arm_qvel": spaces.Box(-10.0, 10.0, (6,))}
        for key, width in _OBS_LAYOUT[self.task][2:]:
            sub[key] = spaces.Box(-10.0, 10.0, (width,))
        # observation keys by mode (reach_cube_env.py:281-295): the images replace the state-only keys (cube positions) in "image" mode
        self._state_only = {"cube_pos", "cube_red_pos", "cube_blue_pos"}
        self._renderer = None
        if observation_mode in ("image", "both"):
            from .render import HEIGHT, WIDTH, BatchRenderer

            self._renderer = BatchRenderer(self)
            keep = {k: v for k, v in sub.items() if k not in self._state_only}
            keep["image_front"] = spaces.Box(0, 255, (HEIGHT, WIDTH, 3), np.uint8)
            keep["image_top"] = spaces.Box(0, 255, (HEIGHT, WIDTH, 3), np.uint8)
            if observation_mode == "both":
                keep.update({k: v for k, v in sub.items() if k in self._state_only})
            sub = keep
        self.single_observation_space = spaces.Dict(sub)
        self.action_space = spaces.batch_space(self.single_action_space, n)
        self.observation_space = spaces.batch_space(self.single_observation_space, n)
        # like gymnasium, an env that is never given a seed draws one from the OS (reference: Env.reset(seed=None))
        self.seed(int(np.random.SeedSequence().entropy % (1 << 62)) if seed is None else seed)

    @classmethod
    def _pick_exec_mode(cls, exec_mode, num_envs):
        """"auto" = the fastest mode measured on B200 for the batch size (profiles/README.md, stationary window): the lockstep
        kernel (one launch per step, CTAs of 16 envs aligned at the phase boundaries, CTA-wide narrowphase job pool) for small
        batches; the phased chain (one small kernel per mj_step phase over all envs, the whole step replayed as ONE CUDA graph)
        from 256 envs (ReachCube mid-episode, ms per step phased / lockstep: 5.4 / 5.9 at 256 envs, 6.5 / 8.7 at 512, 6.6 / 7.2 at
        1 024, 8.8 / 13.0 at 4 096; profiles/r02b_sweep_groups.txt).  "flow" (one persistent kernel per step, phases run from
        device-side queues, csrc/lcr_flow.cuh) reaches 0.65 - 1.0 of them and is selectable.  All modes give bit-identical
        results (tests/test_gpu_parity.py)."""
        if exec_mode != "auto":
            return exec_mode
        return "phased" if num_envs >= 256 else "lockstep"

    @property
    def obs_layout(self):
        """((key, width), ...) of the flat observation ``[num_envs, obs_dim]``, in the reference's key order"""
        return _OBS_LAYOUT[self.task]

    # -- helpers ---------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _split(self, flat):
        """observation dict of the mode, in the reference's key order: state keys, images, state-only keys"""
        out, late, k = {}, {}, 0
        for key, width in _OBS_LAYOUT[self.task]:
            (late if key in self._state_only else out)[key] = flat[:, k:k + width]
            k += width
        if self._renderer is not None:
            out.update({key: img.clone() for key, img in self._renderer.render().items()})
        if self.observation_mode != "image":
            out.update(late)
        return out

    def seed(self, seed, mask=None):
        """Env ``i`` gets the stream of ``np.random.default_rng(seed + env_offset + i)``; with ``mask`` (uint8 device
        tensor) only the selected envs are reseeded, the streams of the others go on."""
        base = int(seed) + self.env_offset
        st = pcg64_states(range(base, base + self.num_envs))
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_seed(self._h, st.ctypes.data_as(C.c_void_p), _ptr(mask), self._stream()))

    # -- gymnasium-style API -----------------------------------------------------------------
    def reset(self, seed=None, options=None, mask=None):
        """Reference ``reset`` (reach_cube_env.py:297-311).  ``mask`` (bool[num_envs]) resets a subset."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            if m.shape != (self.num_envs,):
                raise ValueError("mask must have shape (num_envs,)")
        if seed is not None:
            self.seed(seed, m)
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_reset(self._h, _ptr(m), _ptr(self._obs), self._stream()))
        return self._split(self._obs.clone()), {}

    def step_flat(self, actions, record=None):
        """One control step; returns views of the internal output buffers (no copies).  ``record`` (optional,
        float32 ``[num_envs, obs_dim + 4]``) also receives the packed record, written by the step kernels themselves."""
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions), device=self.device)
        if tuple(actions.shape) != (self.num_envs, self.action_dim):
            raise ValueError("Action dimension mismatch")  # reach_cube_env.py:231-232
        a = actions.to(device=self.device, dtype=torch.float32).contiguous()
        f = self._flags
        if self.nvtx:
            torch.cuda.nvtx.range_push(f"lcr_step[{self.task} x{self.num_envs} {self.exec_mode}]")
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_step_rec(self._h, _ptr(a), _ptr(self._obs), _ptr(self._reward), _ptr(f[0]), _ptr(f[1]),
                                            _ptr(f[2]), _ptr(record), self._stream()))
        if self.nvtx:
            torch.cuda.nvtx.range_pop()
        return self._obs, self._reward, f[0], f[1], f[2]

    def step_packed(self, actions, out=None):
        """One control step, outputs as one float32 record per env ``[num_envs, obs_dim + 4]`` = obs | reward |
        terminated | truncated | success, written by the step kernels (no pack launch): the all-gather / device->host
        unit of ``dist.ShardedEnv``."""
        if out is None:
            if self._record is None:
                self._record = torch.empty(self.num_envs, self.obs_dim + 4, dtype=torch.float32, device=self.device)
            out = self._record
        if out.dtype != torch.float32 or tuple(out.shape) != (self.num_envs, self.obs_dim + 4) or not out.is_contiguous():
            raise ValueError("record buffer must be a contiguous float32 [num_envs, obs_dim + 4] tensor")
        self.step_flat(actions, record=out)
        return out

    def step(self, actions):
        obs, reward, te, tr, su = self.step_flat(actions)
        info = {} if self.task == "lift" else {"is_success": su.bool()}  # lift_cube_env.py:337
        if self.report_failures:
            info.update(self.failure_info())
        return self._split(obs.clone()), reward.clone(), te.bool(), tr.bool(), info

    report_failures = False  # constructor kwarg; True: every step() adds failure_info() to info (one more tiny launch per step;
                             # off by default so that info carries exactly the reference's keys)
    nvtx = False             # True: NVTX range around every step's enqueue (and around the all-gather in dist.ShardedEnv)

    def failure_info(self):
        """Per-env failure flags of the last step (device tensors, no host sync): ``nan_reset`` = the env blew up
        (NaN / huge state or acceleration) and was restored by mj_resetData like MuJoCo does, ``nan_resets`` the running
        count, ``contact_overflow`` = contacts dropped past the big workspace caps (0 on every tested workload)."""
        d = self.diagnostics()
        prev = getattr(self, "_nan_seen", None)
        now = d["nan_resets"]
        flag = now > prev if prev is not None else now > 0
        self._nan_seen = now.clone()
        return {"nan_reset": flag, "nan_resets": now, "contact_overflow": d["overflow"]}

    def flow_status(self):
        """(unfinished envs, watchdog code) of the last flow-mode step; (0, 0) when healthy.  Synchronises the device."""
        s = (C.c_int32 * 8)()
        capi.check(self._L.lcr_flow_status(self._h, s))
        self.flow_debug = [int(x) for x in s[2:]]
        return int(s[0]), int(s[1])

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.lcr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self):
        raise NotImplementedError("rendering is out of scope")

    # -- state access (replaces env.data.qpos / qvel / ctrl pokes; checkpoint / resume) ----------
    def get_state(self):
        n, dev = self.num_envs, self.device
        z = lambda w, dt=torch.float64: torch.zeros(n, w, dtype=dt, device=dev)
        st = dict(qpos=z(self.nq), qvel=z(self.nv), ctrl=z(6), warm=z(self.nv), aux=z(model.NAUX),
                  ints=z(model.NINT, torch.int32), rng=z(4, torch.int64))  # rng: the 4 x uint64 PCG64 state, bit pattern in int64
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_get_state(self._h, *[_ptr(st[k]) for k in ("qpos", "qvel", "ctrl", "warm", "aux", "ints", "rng")],
                                             self._stream()))
        return st

    def set_state(self, qpos=None, qvel=None, ctrl=None, warm=None, aux=None, ints=None, rng=None):
        def prep(x, w, dt=torch.float64):
            if x is None:
                return None
            t = torch.as_tensor(x).to(device=self.device, dtype=dt).contiguous()
            if tuple(t.shape) != (self.num_envs, w):
                raise ValueError(f"expected shape {(self.num_envs, w)}, got {tuple(t.shape)}")
            return t
        args = [prep(qpos, self.nq), prep(qvel, self.nv), prep(ctrl, 6), prep(warm, self.nv), prep(aux, model.NAUX),
                prep(ints, model.NINT, torch.int32), prep(rng, 4, torch.int64)]
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_set_state(self._h, *[_ptr(a) for a in args], self._stream()))
            torch.cuda.current_stream(self.device).synchronize()  # args are temporaries

    def substeps(self, n):
        """``n`` raw ``mj_step`` calls with the current ctrl (``n == 0``: ``mj_forward`` only)."""
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_substeps(self._h, int(n), self._stream()))

    def inverse_kinematics(self, ee_target_pos):
        """Batched damped-least-squares IK from the current arm pose (reach_cube_env.py:148-221)."""
        t = torch.as_tensor(ee_target_pos).to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(t.shape) != (self.num_envs, 3):
            raise ValueError("ee_target_pos must have shape (num_envs, 3)")
        q = torch.zeros(self.num_envs, 6, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_ik(self._h, _ptr(t), _ptr(q), self._stream()))
        return q

    def diagnostics(self):
        d = torch.zeros(self.num_envs, model.NDIAG, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_get_diag(self._h, _ptr(d), self._stream()))
        return dict(zip(("ncon", "nefc", "niter", "max_nefc", "overflow", "nan_resets"), d.unbind(1)))

    def debug_contacts(self):
        """(contacts [n, MAXCON, 12] float64, ncon [n]) of ``mj_forward`` on the current state (state untouched)."""
        c = torch.zeros(self.num_envs, model.MAXCON, 12, dtype=torch.float64, device=self.device)
        k = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_debug_contacts(self._h, _ptr(c), _ptr(k), self._stream()))
        return c, k

    @property
    def kernel_launches(self):
        return int(self._L.lcr_kernel_launches(self._h))


class ReachCubeEnv(BatchedLowCostRobotEnv):
    task = "reach"


class PushCubeEnv(BatchedLowCostRobotEnv):
    task = "push"


class LiftCubeEnv(BatchedLowCostRobotEnv):
    task = "lift"


class PickPlaceCubeEnv(BatchedLowCostRobotEnv):
    task = "pick_place"


class StackTwoCubesEnv(BatchedLowCostRobotEnv):
    task = "stack"


class PushCubeLoopEnv(BatchedLowCostRobotEnv):
    """``PushCubeLoop-v0`` (push_cube_loop_env.py): push the cube back and forth between two goal regions inside four
    rails.  The reference constructor takes ``observation_mode, action_mode, block_gripper, n_substeps, render_mode``
    (:77-84); ``step`` never terminates (:331-333, the registered TimeLimit truncates), the reward is the overlap
    reward of ``get_reward`` (:337-365) and ``info`` carries ``timestamp`` and ``success`` (:329).  The goal an env
    currently pushes towards (``current_goal``, :136) lives in ``aux[:, 1]`` of ``get_state`` / ``set_state``."""

    task = "push_loop"

    def __init__(self, num_envs=1, device="cuda:0", observation_mode="state", action_mode="joint", block_gripper=True,
                 n_substeps=20, render_mode=None, **kwargs):
        for k in ("reward_type", "distance_threshold", "height_threshold", "cube_xy_range", "target_xy_range", "goal_z_range"):
            if k in kwargs:
                raise TypeError(f"PushCubeLoopEnv has no argument {k!r}")  # not in the reference signature
        super().__init__(num_envs=num_envs, device=device, observation_mode=observation_mode, action_mode=action_mode,
                         block_gripper=block_gripper, n_substeps=n_substeps, render_mode=render_mode, **kwargs)

    def _timestamp(self):
        aux = torch.zeros(self.num_envs, model.NAUX, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            capi.check(self._L.lcr_get_state(self._h, None, None, None, None, _ptr(aux), None, None, self._stream()))
        return aux[:, 0]

    @property
    def current_goal(self):
        """0 / 1 per env: the goal region the cube is to be pushed into (push_cube_loop_env.py:136,346)."""
        return self.get_state()["aux"][:, 1].to(torch.int64)

    def reset(self, seed=None, options=None, mask=None):
        obs, _ = super().reset(seed=seed, options=options, mask=mask)
        return obs, {"timestamp": 0.0}  # push_cube_loop_env.py:320

    def step(self, actions):
        obs, reward, te, tr, su = self.step_flat(actions)
        info = {"timestamp": self._timestamp(), "success": su.to(torch.int64)}  # push_cube_loop_env.py:329
        if self.report_failures:
            info.update(self.failure_info())
        return self._split(obs.clone()), reward.clone(), te.bool(), tr.bool(), info


ENV_CLASSES = {"reach": ReachCubeEnv, "push": PushCubeEnv, "lift": LiftCubeEnv, "pick_place": PickPlaceCubeEnv,
               "stack": StackTwoCubesEnv, "push_loop": PushCubeLoopEnv}
