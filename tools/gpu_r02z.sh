#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02z7.txt 2>&1
tail -2 gpurun_out/pytest_r02z7.txt
timeout 600 python tools/ab_check.py 13=4 10=6 > gpurun_out/ab_r02z7.txt 2>&1
timeout 900 python tools/kbench.py --blocks 32 default 10=6 13=4 > gpurun_out/kbench_r02z7.txt 2>&1
cat gpurun_out/ab_r02z7.txt gpurun_out/kbench_r02z7.txt
