#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_loop2.py > gpurun_out/debug_loop2.txt 2>&1
tail -5 gpurun_out/debug_loop2.txt
