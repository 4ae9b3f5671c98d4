python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kbench.py --blocks 32 --iters 10 default 2=1 > gpurun_out/kbench_a.txt 2>&1
cat gpurun_out/kbench_a.txt
