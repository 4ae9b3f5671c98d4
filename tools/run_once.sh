timeout 150 python tools/ab_check.py 12=2 6=2 2>&1 | grep -i "variant\|error" | head -6
timeout 300 python tools/kbench.py --blocks 32 --iters 10 default 6=2 12=4 12=8 2>&1 | tail -4
