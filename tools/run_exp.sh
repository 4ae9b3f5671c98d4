timeout 120 python tools/ab_check.py 11=1 2>&1 | grep -i "variant\|error\|Traceback" | head
timeout 300 python tools/kbench.py --blocks 32 --iters 10 default 11=1 > gpurun_out/kbench_tma.txt 2>&1
cat gpurun_out/kbench_tma.txt | tail -4
