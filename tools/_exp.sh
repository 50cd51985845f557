#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep33.jsonl
run() { echo "{\"label\": \"$1\"}" >> gpurun_out/sweep33.jsonl; shift; env "$@" >> gpurun_out/sweep33.jsonl 2>> gpurun_out/sweep33.err; }
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
run "reach5k auto" $B --envs 5120
run "reach5k lockstep" $B --envs 5120 --exec-mode lockstep
run "reach8k auto" $B --envs 8192
run "reach8k G3" LCR_GROUPS=3 $B --envs 8192
run "reach16k G3" LCR_GROUPS=3 $B --envs 16384
run "reach64k auto" $B --envs 65536 --steps 10
run "push16k auto" $B --task push --envs 16384
run "pp8k ee auto" $B --task pick_place --action-mode ee --envs 8192
run "stack8k auto" $B --task stack --envs 8192
run "loop16k auto" $B --task push_loop --envs 16384
tail -3 gpurun_out/sweep33.err
