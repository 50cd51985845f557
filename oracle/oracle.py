"""ctypes front-end of the CPU float64 oracle (oracle/lcr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gym_lowcostrobot_b200 import config as _config
from gym_lowcostrobot_b200 import model as _model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liblcr_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("lcr_oracle.c", "lcr_oracle_convex.inc")] + [
        os.path.join(_HERE, "..", "include", "lcr_model.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liblcr_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_rng_double.restype = C.c_double
        for f in ("orc_destroy", "orc_forward", "orc_substep"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_substeps.argtypes = [C.c_void_p, C.c_int]
        assert L.orc_sizeof_model() == C.sizeof(_model.LcrModel), "LcrModel layout mismatch"
        assert L.orc_sizeof_cfg() == C.sizeof(_model.LcrEnvCfg), "LcrEnvCfg layout mismatch"
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pcg64_state(seed):
    """numpy PCG64 bit-generator state of ``default_rng(seed)`` as 4 uint64 (hi/lo of state, inc)."""
    st = np.random.PCG64(seed).state["state"]
    m = (1 << 64) - 1
    return np.array([st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)


class Oracle:
    """One float64 environment."""

    def __init__(self, task, compiled=None, **kwargs):
        self.task = task
        self.compiled = compiled if compiled is not None else _model.load_compiled(task)
        self.cmodel, self.verts = _model.pack_model(self.compiled)
        self.cfg = _config.make_cfg(task, **kwargs)
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create(C.byref(self.cmodel), _p(self.verts), C.byref(self.cfg)))
        self.nq, self.nv = self.cmodel.nq, self.cmodel.nv
        self.na, self.no = _config.action_dim(self.cfg), _config.obs_dim(task)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    # the choices that could not be checked against MuJoCo's source, as named switches (see OrcSim in lcr_oracle.c)
    SWITCHES = {"plane_hull_tilt": 1e-3, "implicit_kv_when_clamped": 1, "impratio": 0}

    def set_switch(self, name, value):
        self.L.orc_set_switch.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        if self.L.orc_set_switch(self.h, name.encode(), float(value)) != 0:
            raise KeyError(f"unknown switch {name!r}; available: {sorted(self.SWITCHES)}")

    def seed(self, seed):
        st = pcg64_state(seed)
        self.L.orc_seed(self.h, _p(st))

    def reset(self, seed=None):
        if seed is not None:
            self.seed(seed)
        obs = np.zeros(self.no, np.float32)
        self.L.orc_reset(self.h, _p(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, np.float32)
        if a.shape != (self.na,):
            raise ValueError("Action dimension mismatch")
        obs = np.zeros(self.no, np.float32)
        r = np.zeros(1, np.float32)
        f = np.zeros(3, np.uint8)
        self.L.orc_step(self.h, _p(a), _p(obs), _p(r), _p(f[0:1]), _p(f[1:2]), _p(f[2:3]))
        return obs, float(r[0]), bool(f[0]), bool(f[1]), bool(f[2])

    def forward(self):
        self.L.orc_forward(self.h)

    def substep(self, n=1):
        self.L.orc_substeps(self.h, int(n))

    def ik(self, target):
        t = np.ascontiguousarray(target, np.float32)
        q = np.zeros(6, np.float32)
        self.L.orc_ik(self.h, _p(t), _p(q))
        return q

    def get_state(self):
        qpos, qvel = np.zeros(self.nq), np.zeros(self.nv)
        ctrl, warm, aux = np.zeros(6), np.zeros(self.nv), np.zeros(_model.NAUX)
        ints = np.zeros(_model.NINT, np.int32)
        rng = np.zeros(4, np.uint64)
        self.L.orc_get_state(self.h, _p(qpos), _p(qvel), _p(ctrl), _p(warm), _p(aux), _p(ints))
        self.L.orc_get_rng(self.h, _p(rng))
        return dict(qpos=qpos, qvel=qvel, ctrl=ctrl, warm=warm, aux=aux, ints=ints, rng=rng)

    def set_state(self, qpos=None, qvel=None, ctrl=None, warm=None, aux=None, ints=None, rng=None):
        f = lambda a, t=np.float64: None if a is None else np.ascontiguousarray(a, t)
        args = [f(qpos), f(qvel), f(ctrl), f(warm), f(aux), f(ints, np.int32)]
        self.L.orc_set_state(self.h, *[_p(a) for a in args])
        if rng is not None:
            self.L.orc_seed(self.h, _p(np.ascontiguousarray(rng, np.uint64)))

    def peaks(self):
        """lifetime maxima (ncon, nefc, convex candidates of one substep) and the total of dropped contacts"""
        d = np.zeros(4, np.int32)
        self.L.orc_get_peaks(self.h, _p(d))
        return dict(zip(("ncon", "nefc", "ncand", "overflow"), d.tolist()))

    def get(self, name):
        buf = np.zeros(_model.MAXEFC * _model.MAXNV)
        n = self.L.orc_get(self.h, name.encode(), _p(buf), buf.size)
        if n < 0:
            raise KeyError(name)
        return buf[:n].copy()

    def diag(self):
        d = np.zeros(_model.NDIAG, np.int32)
        self.L.orc_get_diag(self.h, _p(d))
        return dict(zip(("ncon", "nefc", "niter", "max_nefc", "overflow", "nan_resets"), d.tolist()))

    def rng_double(self):
        return float(self.L.orc_rng_double(self.h))
