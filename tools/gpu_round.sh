#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines of the BASELINE configs, ncu launch list and one full capture of the
# step kernel (outputs in gpurun_out/; tools/summarize_profiles.py <tag> turns them into the tracked files of profiles/).
# usage (from the repo root):  gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [quick]'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
B="timeout 600 python bench.py --steps 30 --warmup 5"
$B > gpurun_out/bench_reach4096.json 2> gpurun_out/bench_reach4096.err
$B --no-cpu-baseline --task push --envs 16384 > gpurun_out/bench_push16384.json 2> gpurun_out/bench_push16384.err
$B --no-cpu-baseline --task pick_place --action-mode ee --envs 8192 > gpurun_out/bench_pickplace_ee8192.json 2> gpurun_out/bench_pickplace_ee8192.err
$B --no-cpu-baseline --task stack --envs 8192 > gpurun_out/bench_stack8192.json 2> gpurun_out/bench_stack8192.err
$B --no-cpu-baseline --task push_loop > gpurun_out/bench_pushloop4096.json 2> gpurun_out/bench_pushloop4096.err
$B --no-cpu-baseline --task push_loop --envs 16384 > gpurun_out/bench_pushloop16384.json 2> gpurun_out/bench_pushloop16384.err
$B --no-cpu-baseline --envs 16384 > gpurun_out/bench_reach16384.json 2> gpurun_out/bench_reach16384.err
$B --no-cpu-baseline --envs 65536 --steps 10 > gpurun_out/bench_reach65536.json 2> gpurun_out/bench_reach65536.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
[ "$1" = quick ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_phased16k.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --envs 16384 > gpurun_out/ncu_list_phased.log 2>&1
[ "$1" = lists ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lockstep.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_ls -s 20 -c 1 -o gpurun_out/prof_lockstep python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 300 python tools/phase_clocks.py ReachCube-v0 4096 25 > gpurun_out/phase_clocks.txt 2>&1
ls -la gpurun_out
