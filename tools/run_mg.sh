# usage: run_mg.sh N mode [mode ...]   -> gpurun_out/scale_<N>_<mode>.json / .log
N=$1; shift
port=29540
for mode in "$@"; do
  port=$((port+1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 10 --warmup 3 --mg-mode $mode > gpurun_out/scale_${N}_${mode}.log 2>&1
  echo "== $mode rc=$?"
  grep '^{' gpurun_out/scale_${N}_${mode}.log | tail -1 > gpurun_out/scale_${N}_${mode}.json
  python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/scale_${N}_${mode}.json").read())
    print("$mode", d["n_gpus"], "value", round(d["value"]), "stream", round(d["config"]["stream_msps"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
    print("   ", d["config"]["parallelism"][:230])
    print("   ", {k: round(v["avg_ms"],4) for k,v in d["kernels"].items()})
except Exception as ex:
    print("$mode FAILED", ex)
    import subprocess; print(subprocess.run("grep -v Warning gpurun_out/scale_${N}_${mode}.log | tail -25", shell=True, capture_output=True, text=True).stdout[-3000:])
EOF
done
