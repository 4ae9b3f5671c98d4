#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_check.py static=0 > gpurun_out/ab_r02x.txt 2>&1
timeout 600 python tools/kbench.py --config cfg4 --blocks 32 default static=0 > gpurun_out/kbench_r02x.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_r02x.txt 2>&1
timeout 600 python bench.py --config cfg4 > gpurun_out/bench_r02x_cfg4.json 2> gpurun_out/bench_r02x_cfg4.err
tail -3 gpurun_out/pytest_r02x.txt; cat gpurun_out/ab_r02x.txt gpurun_out/kbench_r02x.txt
