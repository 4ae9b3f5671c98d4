#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -q > gpurun_out/r02g_variants.txt 2>&1
timeout 900 python -m pytest tests/test_filter_abi.py -m gpu -q > gpurun_out/r02g_abi.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_configs.py -q > gpurun_out/r02g_configs.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_variants.py --deselect tests/test_filter_abi.py --deselect tests/test_gpu_configs.py > gpurun_out/r02g_pytest.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default > gpurun_out/kbench_r02g.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err
echo done
