#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_check.py 9=1 > gpurun_out/ab_r02z8.txt 2>&1
timeout 900 python tools/kbench.py --blocks 32 default 9=1 > gpurun_out/kbench_r02z8.txt 2>&1
head -3 gpurun_out/ab_r02z8.txt; cat gpurun_out/kbench_r02z8.txt
