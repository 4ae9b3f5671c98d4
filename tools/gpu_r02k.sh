#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py 13=3 > gpurun_out/ab_r02k.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default 13=3 > gpurun_out/kbench_r02k.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -q > gpurun_out/r02k_pytest.txt 2>&1
echo done
