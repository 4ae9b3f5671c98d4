#!/bin/bash
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/kbench.py --blocks 2 --iters 1 --rounds 1 default > gpurun_out/sanitizer_cfg2.txt 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_cfg2.txt
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/kbench.py --config cfg4 --blocks 2 --iters 1 --rounds 1 default > gpurun_out/sanitizer_cfg4.txt 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_cfg4.txt
tail -6 gpurun_out/sanitizer_cfg2.txt; tail -6 gpurun_out/sanitizer_cfg4.txt
