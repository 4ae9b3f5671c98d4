#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_check.py 10=6 > gpurun_out/ab_r02y.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 default 10=6 > gpurun_out/kbench_r02y.txt 2>&1
cat gpurun_out/ab_r02y.txt gpurun_out/kbench_r02y.txt
