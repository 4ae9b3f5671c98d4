#!/bin/bash
# evidence pass: tests, bench lines of every single-GPU configuration + the reference arm, launch list, one ncu --set full capture
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02v_pytest.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_r02_cfg2.json 2> gpurun_out/bench_r02_cfg2.err
for c in cfg3 cfg4; do
  timeout 900 python bench.py --config $c --no-filter-h > gpurun_out/bench_r02_$c.json 2> gpurun_out/bench_r02_$c.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fwd_|chan_|notch' -s 6 -c 60 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 4 --warmup 3 --quick --no-cpu-baseline > gpurun_out/launches_r02.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd_cols_r36|fwd_rows_r50|chan_v2' -s 6 -c 3 -o gpurun_out/prof_r02_final -f \
    python tools/kbench.py --blocks 32 --iters 2 --rounds 1 default > gpurun_out/ncu_r02_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd_cols_2s|fwd_rows_2s' -s 4 -c 2 -o gpurun_out/prof_r02_cfg4 -f \
    python tools/kbench.py --config cfg4 --blocks 32 --iters 2 --rounds 1 default > gpurun_out/ncu_r02_cfg4.log 2>&1
timeout 600 python tools/kbench.py --blocks 32 --iters 10 --rounds 3 default 10=6 13=4 14=1,15=-1 > gpurun_out/kbench_r02_final.txt 2>&1
timeout 600 python tools/kbench.py --config cfg4 --blocks 32 --iters 10 --rounds 3 default static=0 > gpurun_out/kbench_r02_cfg4.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.txt 2>&1
tail -2 gpurun_out/r02v_pytest.txt; cat gpurun_out/kbench_r02_final.txt; tail -2 gpurun_out/smoke_r02.txt
echo done
