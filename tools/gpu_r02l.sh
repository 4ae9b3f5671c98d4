#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ab_check.py 13=5 13=4 > gpurun_out/ab_r02l.txt 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default 13=5 13=4 12=4 12=8 > gpurun_out/kbench_r02l.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd_cols_r36|fwd_rows_v2|chan_v2' -s 6 -c 3 -o gpurun_out/prof_r02l -f \
    python tools/kbench.py --blocks 32 --iters 2 --rounds 1 default > gpurun_out/ncu_r02l.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest.txt 2>&1
echo done
