#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -x -q > gpurun_out/r02d_variants.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_variants.py > gpurun_out/r02d_pytest.txt 2>&1
timeout 300 python tools/ab_check.py 13=2 > gpurun_out/ab_r02d.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default 13=2 > gpurun_out/kbench_r02d.txt 2>&1
echo done
