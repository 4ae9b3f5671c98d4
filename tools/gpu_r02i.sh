#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_lin.py > gpurun_out/r02i_diag.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err
echo done
