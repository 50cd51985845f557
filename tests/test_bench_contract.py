"""The reference arm of bench.py (`--impl reference`: the CPU port on the host cores) runs without a GPU: check here that its
one JSON line carries the keys of the measurement contract, and that the GPU arm's static tables are consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--envs", "64",
                          "--task", "push"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env-steps/sec" and line["unit"] == "env-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"]["workload"].startswith("PushCube-v0 64 envs") and line["config"]["preroll_steps"] == 100
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb and cb["per_thread_1t"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None


def test_static_tables_of_the_gpu_arm():
    sys.path.insert(0, ROOT)
    import bench

    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert base["metric"].lower().replace("-", " ").startswith("env steps") or "env-steps" in base["metric"].lower()
    assert set(bench.ALGO_BYTES) == set(bench.IDS)
    # the headline default = the largest single-GPU configuration, the others are reported in `configs`
    ap_defaults = dict(task="push", envs=16384)
    assert ("reach", 4096, "joint") in bench.OTHER_CONFIGS and bench.SHARDED[4][0] == "pick_place" and bench.SHARDED[8][0] == "stack"
    facts = bench.profile_facts(ap_defaults["task"], ap_defaults["envs"], "phased")
    assert facts["traffic"] > 1e8 and facts["inst_executed"] > 1e9 and "profiles/" in facts["capture"]
