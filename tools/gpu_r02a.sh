#!/bin/bash
# round-2 call A: baseline of the shipped build (tests, bench, ncu --set full at 32 blocks/launch, sub-batch sweep)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
(ldconfig -p | grep -i fftw; ls /usr/lib/x86_64-linux-gnu | grep -i fftw) > gpurun_out/r02a_fftw.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.txt
timeout 600 python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default 12=4 12=8 12=16 > gpurun_out/kbench_r02a.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd_cols_v2|fwd_rows_v2|chan_v2' -s 6 -c 3 -o gpurun_out/prof_r02a -f \
    python tools/kbench.py --blocks 32 --iters 2 --rounds 1 default > gpurun_out/ncu_r02a.log 2>&1
echo done
