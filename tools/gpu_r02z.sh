#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_check.py 13=4 > gpurun_out/ab_r02z6.txt 2>&1
timeout 900 python tools/kbench.py --blocks 32 default 13=4 > gpurun_out/kbench_r02z6.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row_kernel_variants or full_size" > gpurun_out/pytest_r02z6.txt 2>&1
tail -3 gpurun_out/pytest_r02z6.txt
cat gpurun_out/ab_r02z6.txt gpurun_out/kbench_r02z6.txt
