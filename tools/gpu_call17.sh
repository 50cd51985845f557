#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi2.txt
timeout 600 python -m pytest tests/test_vec_and_recorder.py -m gpu -q > gpurun_out/pytest_vec.log 2>&1; tail -2 gpurun_out/pytest_vec.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 600 gpurun_out/bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --task stack --envs 8192 > gpurun_out/bench_2gpu_stack.json 2> gpurun_out/bench_2gpu_stack.err; tail -c 300 gpurun_out/bench_2gpu_stack.json
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 900 gpurun_out/bench_1gpu.json
tail -3 gpurun_out/bench_2gpu.err
