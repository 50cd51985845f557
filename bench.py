#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched env.step() hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--task push] [--envs 16384]

A "step" is one env.step() over the whole batch of envs (20 physics substeps each).  The N=1 workload is the largest
single-GPU configuration of BASELINE.json, configs[2]: PushCube-v0, 16 384 envs, state obs, joint actions; the other
GPU configurations (Reach 4 096 and the per-GPU shares of configs 3 / 4: PickPlace-ee 8 192, Stack 8 192) are measured
after it and reported in `configs`.  Under torchrun every rank runs `--envs` envs (weak scaling) and every step ends
with one NCCL all-gather of the packed output batch; at 4 / 8 GPUs the sharded north-star configurations (PickPlace-ee
32 768 on 4, Stack 65 536 on 8) are run as well (`extra`).

Stationary window: the cost of a step grows through an episode (contacts build up), and with a common TimeLimit all
envs of a batch would reset together -- any window would sample one episode phase.  Both arms therefore stagger the
episode clocks (env i starts with elapsed_steps = i mod 50) and roll the envs for PREROLL untimed env.steps (two
episode lengths) before anything is timed: the batch then holds every episode phase in equal parts, in every window.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import queue
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

IDS = {"reach": "ReachCube-v0", "push": "PushCube-v0", "lift": "LiftCube-v0", "pick_place": "PickPlaceCube-v0",
       "stack": "StackTwoCubes-v0", "push_loop": "PushCubeLoop-v0"}
# algorithmic bytes per env-step: persistent state read + written, action in, obs/reward/flags out
ALGO_BYTES = {"reach": 446, "lift": 450, "push": 482, "pick_place": 478, "stack": 614, "push_loop": 454}
METRIC = "env-steps/sec"
PREROLL = 100
# the other GPU configurations of BASELINE.json, per GPU: (task, envs, action mode)
OTHER_CONFIGS = [("reach", 4096, "joint"), ("pick_place", 8192, "ee"), ("stack", 8192, "joint")]
# north-star sharded configurations by GPU count: (task, envs per GPU, action mode)
SHARDED = {4: ("pick_place", 8192, "ee"), 8: ("stack", 8192, "joint")}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class CpuRollout:
    """Oracle port timed on the host cores.  The envs are cut into small chunks that the worker threads pull from a
    queue (the C rollout releases the GIL): a thread that drew cheap envs takes more chunks, nobody waits for the
    thread that drew the expensive ones."""

    CHUNK = 4

    def __init__(self, task, action_mode, n_envs, n_threads, seed0=0):
        import ctypes as C

        from oracle.oracle import Oracle, lib

        self.C, self.L = C, lib()
        self.sims = [Oracle(task, action_mode=action_mode, autoreset=True) for _ in range(n_envs)]
        for i, s in enumerate(self.sims):
            s.reset(seed=seed0 + i)
            s.set_state(ints=np.array([i % 50, 0], np.int32))  # staggered episode clocks, see the module docstring
        self.na, self.n_envs, self.n_threads = self.sims[0].na, n_envs, n_threads
        self.rng = np.random.default_rng(1234)
        self.chunks = [np.arange(k, min(k + self.CHUNK, n_envs)) for k in range(0, n_envs, self.CHUNK)]
        self.handles = [(C.c_void_p * len(ch))(*[self.sims[i].h for i in ch]) for ch in self.chunks]

    def run(self, n_steps, n_threads=None):
        """n_steps env.steps of every env with fresh U(-1,1) actions; returns (env_steps_per_s, seconds)"""
        C, L = self.C, self.L
        nt = n_threads or self.n_threads
        acts = [self.rng.uniform(-1, 1, size=(n_steps, len(ch), self.na)).astype(np.float32) for ch in self.chunks]
        todo = queue.SimpleQueue()
        for k in range(len(self.chunks)):
            todo.put(k)

        def work():
            while True:
                try:
                    k = todo.get_nowait()
                except queue.Empty:
                    return
                L.orc_rollout(self.handles[k], len(self.chunks[k]), n_steps, acts[k].ctypes.data_as(C.c_void_p), None, None)

        threads = [threading.Thread(target=work) for _ in range(nt)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        return self.n_envs * n_steps / dt, dt


def cpu_sample(task, action_mode, n_envs, cores):
    """(CpuRollout after the pre-roll, sample size): the same bounded sample for the reference arm and cpu_baseline"""
    sample = min(n_envs, 64 * cores)
    cpu = CpuRollout(task, action_mode, sample, cores)
    cpu.run(PREROLL)
    return cpu, sample


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation is MuJoCo, which is not installable here, so
    this times the float64 oracle port (oracle/lcr_oracle.c) on all host cores, same config and metric."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = args.envs
    from oracle.oracle import build

    build()
    # bounded sample: each "step" = 10 env.steps of min(envs, 64*cores) envs on all cores, scaled to the full batch
    cpu, sample = cpu_sample(args.task, args.action_mode, n_envs, cores)
    per_step = []
    for k in range(args.warmup + args.steps):
        rate, dt = cpu.run(10)
        if k >= args.warmup:
            per_step.append(rate)
    value = float(np.mean(per_step))
    one, _ = CpuRollout(args.task, args.action_mode, 32, 1).run(20)  # (fresh envs: an upper bound of the mid-episode per-thread rate)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_envs / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.task, n_envs, args.action_mode) + " (CPU)", "envs": n_envs, "action_mode": args.action_mode,
                   "preroll_steps": PREROLL, "episode_clocks": "staggered (env i starts at elapsed = i mod 50)"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "per_thread_1t": one,
                         "sample": f"{sample} envs x 10 env.steps per timed step on {cores} threads (dynamic queue of {CpuRollout.CHUNK}-env chunks) after "
                                   f"{PREROLL} untimed steps, scaled to {n_envs} envs; MuJoCo itself is not installable in this image, the port is "
                                   "oracle/lcr_oracle.c"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(task, n, mode):
    return f"{IDS[task]} {n} envs/GPU, state obs, {mode} action, 20 substeps, random U(-1,1) actions, next-step autoreset (TimeLimit 50)"


class GpuRun:
    """One workload on this rank's GPU: pre-roll, device-timed steps, host-buffer (e2e) steps."""

    def __init__(self, task, n_local, action_mode, exec_mode, rank, world, local_rank, K, W):
        import torch

        import gym_lowcostrobot_b200 as glr

        self.torch, self.task, self.n_local, self.world, self.rank, self.K, self.W = torch, task, n_local, world, rank, K, W
        self.dev = torch.device("cuda", local_rank)
        self.env = glr.make(IDS[task], num_envs=n_local, device=f"cuda:{local_rank}", action_mode=action_mode, autoreset=True,
                            env_offset=rank * n_local, exec_mode=exec_mode)
        self.env.reset(seed=0)
        ints = torch.zeros(n_local, 2, dtype=torch.int32, device=self.dev)
        ints[:, 0] = (torch.arange(n_local, device=self.dev) + rank * n_local) % 50  # staggered episode clocks, see the module docstring
        self.env.set_state(ints=ints)
        self.A, self.O = self.env.action_dim, self.env.obs_dim
        gen = self.gen = torch.Generator(device=self.dev).manual_seed(1234 + rank)
        self.pre = torch.rand(PREROLL, n_local, self.A, generator=gen, device=self.dev) * 2 - 1
        self.actions = torch.rand(W + K, n_local, self.A, generator=gen, device=self.dev) * 2 - 1  # resident in HBM
        self.rec = torch.empty(n_local, self.O + 4, dtype=torch.float32, device=self.dev)
        self.full = torch.empty(n_local * world, self.O + 4, dtype=torch.float32, device=self.dev) if world > 1 else None
        for t in range(PREROLL):
            self.env.step_flat(self.pre[t])
        torch.cuda.synchronize()

    def step(self, a):
        """the public step: packed record written by the step kernels, all-gathered when sharded"""
        import torch.distributed as dist

        self.env.step_packed(a, out=self.rec)
        if self.world > 1:
            dist.all_gather_into_tensor(self.full, self.rec)
            return self.full
        return self.rec

    def check_gather(self):
        """once: the gathered batch is the concatenation of the ranks' records (rank 0 checks against point-to-point copies)"""
        import torch.distributed as dist

        torch = self.torch
        if self.world == 1:
            return True
        parts = [torch.empty_like(self.rec) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(self.rec, parts, dst=0)
        ok = torch.tensor([1], device=self.dev)
        if self.rank == 0:
            ok[0] = int(torch.equal(torch.cat(parts), self.full))
        dist.broadcast(ok, 0)
        return bool(ok.item())

    def timed(self, flush):
        import torch.distributed as dist

        torch, env, K, W = self.torch, self.env, self.K, self.W
        for t in range(W):
            self.step(self.actions[t])
        torch.cuda.synchronize()
        gather_ok = self.check_gather()
        if self.world > 1:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        l0 = env.kernel_launches
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()  # (ncu --profile-from-start off: only the timed steps are captured)
        for t in range(K):
            flush.zero_()  # L2 flush between timed iterations (outside the event pairs)
            ev[t][0].record()
            kev[t][0].record()
            env.step_packed(self.actions[W + t], out=self.rec)
            kev[t][1].record()
            if self.world > 1:
                dist.all_gather_into_tensor(self.full, self.rec)
            ev[t][1].record()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        if self.world > 1:
            dist.barrier()
        launches = env.kernel_launches - l0
        step_ms = sum(a.elapsed_time(b) for a, b in ev)
        kern_ms = sum(a.elapsed_time(b) for a, b in kev) / K
        tmax = torch.tensor([step_ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
        return dict(value=self.n_local * self.world * K / (total_ms * 1e-3), ms_per_step=total_ms / K, kernel_ms=kern_ms, launches=launches,
                    gather_ok=gather_ok)

    def e2e(self):
        """public API with HOST buffers: H2D of the actions, the step (+ all-gather when sharded) and D2H of the result
        batch inside the timed region, the caller reads the result every step"""
        import torch.distributed as dist

        torch, K, W = self.torch, self.K, self.W
        # fresh i.i.d. actions (replaying the timed window's actions would push the same way again and again: the arms drift into the
        # floor / their limits and the steps get heavier -- measured: +40 % after three replays on PickPlace-ee)
        h_act = torch.empty(K, self.n_local, self.A, dtype=torch.float32).pin_memory()
        h_act.copy_((torch.rand(K, self.n_local, self.A, generator=self.gen, device=self.dev) * 2 - 1).cpu())
        n_out = self.n_local * self.world
        h_out = torch.empty(n_out, self.O + 4, dtype=torch.float32).pin_memory()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(K):
            a = h_act[t].to(self.dev, non_blocking=True)
            out = self.step(a)
            h_out.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return dict(value=n_out * K / (float(ms.item()) * 1e-3), unit="env-steps/s", h2d_bytes_per_step=self.n_local * self.A * 4,
                    d2h_bytes_per_step=n_out * (self.O + 4) * 4)

    def close(self):
        self.env.close()


def profile_facts(task, n_local, exec_mode):
    """per-launch DRAM traffic and executed warp instructions of the dominant kernel from the committed ncu captures"""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return {}
    return json.load(open(tp)).get(f"{task}_{n_local}_{exec_mode}", {})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--task", default="push", choices=list(IDS))
    ap.add_argument("--envs", type=int, default=16384, help="envs per GPU")
    ap.add_argument("--action-mode", dest="action_mode", default="joint", choices=["joint", "ee"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline workload only")
    ap.add_argument("--exec-mode", dest="exec_mode", default="auto", choices=["auto", "fused", "phased", "lockstep", "flow"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, args.warmup
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    head = GpuRun(args.task, args.envs, args.action_mode, args.exec_mode, rank, world, local_rank, K, W)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    r = head.timed(flush)
    clocks = sampler.stop() if rank == 0 else None
    e2e = head.e2e()
    exec_mode = head.env.exec_mode  # what "auto" resolved to
    A, O = head.A, head.O
    head.close()

    others, extra = [], None
    if not args.no_configs:
        todo = [(t, n, m) for t, n, m in OTHER_CONFIGS if (t, n, m) != (args.task, args.envs, args.action_mode)]
        for t, n, m in todo:
            g = GpuRun(t, n, m, args.exec_mode, rank, world, local_rank, max(8, K // 2), 3)
            rr, ee = g.timed(flush), g.e2e()
            others.append({"workload": workload_name(t, n, m), "envs_per_gpu": n, "value": rr["value"], "ms_per_step": rr["ms_per_step"],
                           "e2e": ee["value"], "exec_mode": g.env.exec_mode, "gather_ok": rr["gather_ok"],
                           "hbm_frac": ALGO_BYTES[t] * n / (rr["kernel_ms"] * 1e-3) / 1e9 / peaks()[0]})
            g.close()
        if world in SHARDED:  # the north-star configuration of this GPU count, sharded over all ranks
            t, n, m = SHARDED[world]
            extra = next((o for o in others if o["envs_per_gpu"] == n and IDS[t] in o["workload"]), None)
            if extra is not None:
                extra = dict(extra, north_star=f"{IDS[t]} {n * world} envs, {m} action, {world}x B200 shard")

    if rank == 0:
        peak, peak_src = peaks()
        achieved = ALGO_BYTES[args.task] * args.envs / (r["kernel_ms"] * 1e-3) / 1e9
        pf = profile_facts(args.task, args.envs, exec_mode)
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        issue_frac = pf["inst_executed"] / (4 * 148 * sm_hz * r["kernel_ms"] * 1e-3) if pf.get("inst_executed") else None
        kernel = {"lockstep": "k_step_ls (+ k_sched, BIG redo pass)", "fused": "k_step (+ BIG redo pass)", "flow": "k_flow (+ k_sched_flow)",
                  "phased": "k_ph_* chain (2 + 4 x 20 launches per env group, replayed as one CUDA graph; envs beyond the fast caps migrate to BIG passes beside it)"}[exec_mode]
        line = {
            "metric": METRIC, "value": r["value"], "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.task, args.envs, args.action_mode), "envs_per_gpu": args.envs, "envs_total": args.envs * world,
                       "action_mode": args.action_mode, "exec_mode": exec_mode, "preroll_steps": PREROLL,
                       "episode_clocks": "staggered (env i starts at elapsed = i mod 50)",
                       "l2": "256 MiB memset between timed steps (outside the per-step event pairs)",
                       "parallelism": f"env-index shard x{world}" + (", 1 NCCL all-gather of the output batch per step" if world > 1 else "")},
            "e2e": e2e,
            "gpu_launches": r["launches"],
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": pf.get("traffic"),
                         "traffic_capture": pf.get("capture"), "issue_frac": issue_frac, "peak_source": peak_src, "kernel": kernel,
                         "kernel_ms": r["kernel_ms"],
                         "note": "per-step device time of the step kernel(s); the path is latency bound (one warp walks 20 substeps of small dense algebra and "
                                 "collision per ~0.45 KB of state), not HBM bound; issue_frac = executed warp instructions / (4 per clock x 148 SMs x time); "
                                 "see DESIGN.md 4"},
            "configs": others,
        }
        if world > 1:
            line["gather_ok"] = r["gather_ok"]
        if extra is not None:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cpu, sample = cpu_sample(args.task, args.action_mode, args.envs, cores)
            rate, dt = cpu.run(150)
            one, _ = CpuRollout(args.task, args.action_mode, 32, 1).run(20)
            line["cpu_baseline"] = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port", "per_thread_1t": one,
                                    "sample": f"{sample} envs x 150 env.steps on {cores} threads (dynamic queue of {CpuRollout.CHUNK}-env chunks, {dt:.1f} s) after "
                                              f"{PREROLL} untimed steps; float64 oracle port, MuJoCo not installable here; per_thread_1t = 32 fresh envs x 20 steps "
                                              "on one thread"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
