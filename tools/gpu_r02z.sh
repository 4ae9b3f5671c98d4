#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_check.py 8=1 > gpurun_out/ab_r02z5.txt 2>&1
timeout 900 python tools/kbench.py --blocks 32 default 8=1 > gpurun_out/kbench_r02z5.txt 2>&1
timeout 900 python tools/kbench.py --config cfg4 --blocks 32 default >> gpurun_out/kbench_r02z5.txt 2>&1
cat gpurun_out/ab_r02z5.txt gpurun_out/kbench_r02z5.txt
