python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err; tail -c 300 gpurun_out/bench_r01i.err
python tools/benchsum.py gpurun_out/bench_r01i.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01i_ref.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r01i_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"fwd_cols|fwd_rows|chan_v2" -c 40 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"fwd_cols|fwd_rows|chan_v2" --launch-skip 12 --launch-count 5 -o gpurun_out/prof_r01i -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i.log 2>&1
tail -2 gpurun_out/ncu_i.log
