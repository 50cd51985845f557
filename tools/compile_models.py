#!/usr/bin/env python
"""Compile the reference's MJCF + STL assets into gym_lowcostrobot_b200/assets/<task>.npz.

Usage: python tools/compile_models.py [ASSETS_DIR]
ASSETS_DIR defaults to /root/reference/gym_lowcostrobot/assets/low_cost_robot_6dof (any checkout of
perezjln/gym-lowcostrobot works).  The .npz files hold numbers only (tree, inertias, actuator gains,
contact parameters, convex-hull vertices); no reference source text is copied.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gym_lowcostrobot_b200 import mjcf, model  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/gym_lowcostrobot/assets/low_cost_robot_6dof"
os.makedirs(model.ASSETS_DIR, exist_ok=True)
for task in mjcf.TASK_XML:
    m = mjcf.compile_model(src, task)
    m["verts"] = m["verts"].astype(np.float64)
    out = os.path.join(model.ASSETS_DIR, f"{task}.npz")
    np.savez_compressed(out, **m)
    print(task, "->", out, os.path.getsize(out), "bytes; hull vertices", len(m["verts"]), "mesh pairs", len(m["pair_g1"]))
