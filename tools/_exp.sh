#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sanitizer.log
for cfg in "PushCubeLoop-v0 phased" "PushCubeLoop-v0 lockstep" "PushCubeLoop-v0 fused" "StackTwoCubes-v0 phased"; do
  echo "### memcheck $cfg" >> gpurun_out/sanitizer.log
  timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python tools/san_small.py $cfg 24 3 2>&1 | grep -v "^$" | tail -6 >> gpurun_out/sanitizer.log
done
echo "### racecheck PushCubeLoop-v0 lockstep" >> gpurun_out/sanitizer.log
timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python tools/san_small.py PushCubeLoop-v0 lockstep 24 2 2>&1 | grep -v "^$" | tail -8 >> gpurun_out/sanitizer.log
cat gpurun_out/sanitizer.log
