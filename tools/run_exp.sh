timeout 120 python tools/ab_check.py 5=1 8=1 2>&1 | tail -8
timeout 300 python tools/kbench.py --blocks 32 --iters 10 default 8=1 > gpurun_out/kbench_v3.txt 2>&1
cat gpurun_out/kbench_v3.txt
