#!/bin/bash
# Lockstep-kernel tuning sweep on one B200: parity first, then bench lines per (CTA width, barrier flags).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lockstep or phased" > gpurun_out/pytest_ls.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ls.log
tail -3 gpurun_out/pytest_ls.log
: > gpurun_out/sweep.jsonl
for task_envs in "reach 4096" "stack 8192"; do
  set -- $task_envs
  for W in 16 8 4; do
    for F in 31 23 15 7 0; do
      echo "{\"task\": \"$1\", \"W\": $W, \"F\": $F}" >> gpurun_out/sweep.jsonl
      LCR_LS_WARPS=$W LCR_LS_FLAGS=$F timeout 300 python bench.py --task $1 --envs $2 --exec-mode lockstep --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
    done
  done
done
timeout 600 compute-sanitizer --tool racecheck python tools/ls_small.py > gpurun_out/racecheck.log 2>&1
tail -5 gpurun_out/racecheck.log
