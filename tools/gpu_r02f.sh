#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -q > gpurun_out/r02f_variants.txt 2>&1
timeout 900 python -m pytest tests/test_filter_abi.py -m gpu -q > gpurun_out/r02f_abi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_variants.py --deselect tests/test_filter_abi.py > gpurun_out/r02f_pytest.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default > gpurun_out/kbench_r02f.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
echo done
