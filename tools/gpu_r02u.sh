#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py 10=5 10=3 6=1 > gpurun_out/ab_r02u.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default 10=5 10=3 6=1 > gpurun_out/kbench_r02u.txt 2>&1
echo done
