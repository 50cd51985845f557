"""ctypes binding of liblcrsim.so (C-ABI in ``include/lcrsim.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C gym_lowcostrobot_b200/csrc``.
There is no CPU fallback: if the library is missing or CUDA is unavailable the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

from . import model as _model

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LCR_LIB") or os.path.join(_HERE, "liblcrsim.so")  # LCR_LIB: alternative build (experiments)
F32, F64 = 0, 1

# every symbol declared in include/lcrsim.h
SYMBOLS = (
    "lcr_obs_dim", "lcr_action_dim", "lcr_create", "lcr_destroy", "lcr_seed", "lcr_reset", "lcr_step", "lcr_step_rec", "lcr_pack_outputs", "lcr_record_append", "lcr_pose_slots", "lcr_body_poses", "lcr_render",
    "lcr_get_state", "lcr_set_state", "lcr_substeps", "lcr_ik", "lcr_get_diag", "lcr_debug_contacts", "lcr_debug_phase_clocks",
    "lcr_debug_flow_stats", "lcr_flow_status", "lcr_n_envs", "lcr_kernel_launches", "lcr_last_error", "lcr_version", "lcr_sizeof_model",
    "lcr_sizeof_cfg",
)

_LIB = None


class LcrError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LcrError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C gym_lowcostrobot_b200/csrc` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, i = C.c_void_p, C.c_int
        L.lcr_create.argtypes = [vp, vp, vp, i, i, i, C.POINTER(vp)]
        L.lcr_destroy.argtypes = [vp]
        L.lcr_seed.argtypes = [vp, vp, vp, vp]
        L.lcr_reset.argtypes = [vp, vp, vp, vp]
        L.lcr_step.argtypes = [vp] * 8
        L.lcr_step_rec.argtypes = [vp] * 9
        L.lcr_pack_outputs.argtypes = [vp] * 8
        L.lcr_record_append.argtypes = [vp, i, vp, i, vp, vp, i, i, vp, vp, vp, vp, vp, i, vp]
        L.lcr_pose_slots.argtypes = [vp]
        L.lcr_body_poses.argtypes = [vp, vp, vp]
        L.lcr_render.argtypes = [vp, i, i, vp, i, vp, vp, i, i, i, vp, vp]
        L.lcr_get_state.argtypes = [vp] * 9
        L.lcr_set_state.argtypes = [vp] * 9
        L.lcr_substeps.argtypes = [vp, i, vp]
        L.lcr_ik.argtypes = [vp, vp, vp, vp]
        L.lcr_get_diag.argtypes = [vp, vp, vp]
        L.lcr_debug_contacts.argtypes = [vp, vp, vp, vp]
        L.lcr_debug_phase_clocks.argtypes = [vp, vp]
        L.lcr_debug_flow_stats.argtypes = [vp, vp]
        L.lcr_flow_status.argtypes = [vp, vp]
        L.lcr_n_envs.argtypes = [vp]
        L.lcr_kernel_launches.argtypes = [vp]
        L.lcr_action_dim.argtypes = [vp]
        L.lcr_last_error.restype = C.c_char_p
        L.lcr_version.restype = C.c_char_p
        if L.lcr_sizeof_model() != C.sizeof(_model.LcrModel) or L.lcr_sizeof_cfg() != C.sizeof(_model.LcrEnvCfg):
            raise LcrError("ctypes mirrors of LcrModel/LcrEnvCfg do not match the compiled library")
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise LcrError(lib().lcr_last_error().decode())
