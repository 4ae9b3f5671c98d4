#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default > gpurun_out/kbench_r02h.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
for c in cfg3 cfg4; do
  timeout 900 python bench.py --config $c --no-filter-h > gpurun_out/bench_r02h_$c.json 2> gpurun_out/bench_r02h_$c.err
done
echo done
