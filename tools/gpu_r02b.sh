#!/bin/bash
# round-2 call B: 16-column cols tile experiment
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py 13=1 > gpurun_out/ab_r02b.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default 13=1 > gpurun_out/kbench_r02b.txt 2>&1
echo done
