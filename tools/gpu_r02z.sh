#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kbench.py --blocks 32 default 10=6 13=4 > gpurun_out/kbench_r02z9.txt 2>&1
timeout 600 python tools/ab_check.py 13=4 > gpurun_out/ab_r02z9.txt 2>&1
cat gpurun_out/kbench_r02z9.txt; head -3 gpurun_out/ab_r02z9.txt
