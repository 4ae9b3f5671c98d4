#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py 13=5 13=4 > gpurun_out/ab_r02m.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default 13=5 13=4 > gpurun_out/kbench_r02m.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q > gpurun_out/r02m_pytest.txt 2>&1
echo done
