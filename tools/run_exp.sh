timeout 120 python tools/ab_check.py 9=1 2>&1 | grep -v "^complex\|^real" | head -4
timeout 300 python tools/kbench.py --blocks 32 --iters 10 default 9=1 > gpurun_out/kbench_c3.txt 2>&1
cat gpurun_out/kbench_c3.txt
