#!/bin/bash
mkdir -p gpurun_out
P=$PWD/gym_lowcostrobot_b200
: > gpurun_out/sweep16.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep16.jsonl; shift; env "$@" >> gpurun_out/sweep16.jsonl 2>> gpurun_out/sweep16.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach auto" $B
run "reach auto unroll16" LCR_LIB=$P/liblcrsim_xunroll16.so $B
run "reach16k auto" $B --envs 16384
run "reach16k auto unroll16" LCR_LIB=$P/liblcrsim_xunroll16.so $B --envs 16384
run "reach phased G8" LCR_GROUPS=8 $B --exec-mode phased
run "reach16k phased G8" LCR_GROUPS=8 $B --exec-mode phased --envs 16384
run "reach16k phased G2" LCR_GROUPS=2 $B --exec-mode phased --envs 16384
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_ls -s 20 -c 1 -o gpurun_out/prof_ls4 python bench.py --exec-mode lockstep --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_ls4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lockstep.csv python bench.py --exec-mode lockstep --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ls_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_phased16k.csv python bench.py --exec-mode phased --envs 16384 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ph_list.log 2>&1
