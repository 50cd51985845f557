"""GPU diagnostic: flow execution mode against the fused kernel (bitwise), queue health, phase clocks, throughput."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gym_lowcostrobot_b200 as glr
from gym_lowcostrobot_b200 import capi

IDS = {"reach": "ReachCube-v0", "push": "PushCube-v0", "lift": "LiftCube-v0", "pick_place": "PickPlaceCube-v0",
       "stack": "StackTwoCubes-v0", "push_loop": "PushCubeLoop-v0"}


def compare(task, mode, n, steps, ref_mode="fused", precision="float32", max_episode_steps=7):
    envs = [glr.make(IDS[task], num_envs=n, action_mode=mode, autoreset=True, max_episode_steps=max_episode_steps, exec_mode=em,
                     precision=precision) for em in (ref_mode, "flow")]
    for e in envs:
        e.reset(seed=3)
    gen = torch.Generator(device="cuda").manual_seed(0)
    bad = 0
    mx = 0
    for t in range(steps):
        a = torch.rand(n, envs[0].action_dim, generator=gen, device="cuda") * 2 - 1
        ra = envs[0].step_packed(a).clone()
        rb = envs[1].step_packed(a).clone()
        torch.cuda.synchronize()
        st = envs[1].flow_status()
        if st != (0, 0):
            print(f"  step {t}: flow status {st} debug {envs[1].flow_debug}")
            return False
        if not torch.equal(ra, rb):
            d = (ra != rb).any(1).nonzero().flatten()
            bad += 1
            print(f"  step {t}: {len(d)} envs differ, first {d[:8].tolist()} max |diff| {(ra - rb).abs().max().item():.3e}")
            if bad > 2:
                break
        mx = max(mx, int(envs[0].diagnostics()["max_nefc"].max()))
    sa, sb = envs[0].get_state(), envs[1].get_state()
    same = all(torch.equal(sa[k], sb[k]) for k in sa)
    ovf = int(envs[1].diagnostics()["overflow"].sum())
    print(f"{task:10s} {mode:5s} n={n:6d} {precision}: outputs {'OK' if bad == 0 else 'DIFF'} state {'OK' if same else 'DIFF'}  max nefc {mx}  dropped {ovf}")
    for e in envs:
        e.close()
    return bad == 0 and same


def speed(task, mode, n, steps=12, warm=60, exec_mode="flow", stats=False):
    env = glr.make(IDS[task], num_envs=n, action_mode=mode, autoreset=True, exec_mode=exec_mode)
    env.reset(seed=0)
    gen = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand(warm + steps, n, env.action_dim, generator=gen, device="cuda") * 2 - 1
    for t in range(warm):
        env.step_flat(acts[t])
        if t == 0 and exec_mode == "flow" and env.flow_status() != (0, 0):
            print(f"{task} n={n}: flow status {env.flow_status()} after the first step -- abort", flush=True)
            env.close()
            return None
    torch.cuda.synchronize()
    st = None
    if stats and exec_mode == "flow":
        st = torch.zeros(8, dtype=torch.int64, device="cuda")
        capi.check(env._L.lcr_debug_flow_stats(env._h, st.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        env.step_flat(acts[warm + t])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    line = f"{task:10s} {mode:5s} n={n:6d} {exec_mode:8s}: {ms:8.3f} ms/step  {n / ms * 1e3:10.0f} env-steps/s"
    if exec_mode == "flow":
        line += f"  status {env.flow_status()}"
    if st is not None:
        v = st.cpu().numpy().astype(np.float64)
        tot = v.sum()
        line += "  clocks% " + " ".join(f"{k}={100 * x / tot:.1f}" for k, x in zip(("BEG", "DYN", "JOB", "COL", "SOL", "END", "BIG", "idle"), v))
    d = env.diagnostics()
    line += f"  max nefc {int(d['max_nefc'].max())} dropped {int(d['overflow'].sum())}"
    print(line, flush=True)
    env.close()
    return ms


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    t0 = time.time()
    if what in ("all", "check"):
        ok = True
        for task, mode, n, steps in (("reach", "joint", 67, 10), ("reach", "joint", 1500, 16), ("stack", "joint", 700, 12), ("pick_place", "ee", 600, 10),
                                     ("push_loop", "joint", 900, 12), ("push", "joint", 4096, 12)):
            ok &= compare(task, mode, n, steps)
        ok &= compare("push", "joint", 300, 6, precision="float64")
        print("CHECK", "PASS" if ok else "FAIL", f"({time.time() - t0:.0f}s)", flush=True)
    if what == "one":  # one speed line: task mode n exec_mode warm steps
        task, mode, n, em = sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
        speed(task, mode, n, steps=int(sys.argv[7]), warm=int(sys.argv[6]), exec_mode=em, stats=True)
    if what in ("all", "speed"):
        for task, mode, n in (("reach", "joint", 4096), ("push", "joint", 16384), ("pick_place", "ee", 8192), ("stack", "joint", 8192)):
            for em in ("flow", "lockstep" if n <= 4096 else "phased"):
                speed(task, mode, n, exec_mode=em, stats=True)
