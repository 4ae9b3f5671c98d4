#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kbench.py --blocks 32 default 12=4 12=8 12=16 12=8,15=-1 > gpurun_out/kbench_r02z4.txt 2>&1
cat gpurun_out/kbench_r02z4.txt
