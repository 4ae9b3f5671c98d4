# usage: run_mg.sh N mode[:blocks[:depth]] ...   -> gpurun_out/scale_<N>_<tag>.json / .log
N=$1; shift
port=29540
for spec in "$@"; do
  port=$((port+1))
  IFS=: read mode blocks depth <<< "$spec"
  blocks=${blocks:-32}; depth=${depth:-2}
  tag=${mode}_b${blocks}_d${depth}${MG_TAG}
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 10 --warmup 3 --mg-mode $mode --blocks-per-step $blocks --depth $depth $MG_EXTRA > gpurun_out/scale_${N}_${tag}.log 2>&1
  echo "== $tag rc=$?"
  grep '^{' gpurun_out/scale_${N}_${tag}.log | tail -1 > gpurun_out/scale_${N}_${tag}.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${N}_${tag}.json").read())
    print("$tag", d["n_gpus"], "value", round(d["value"]), "stream", round(d["config"]["stream_msps"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "parity", d["parity"].get("max_rel_err"), d["parity"].get("checked"))
    print("   ", d["config"]["parallelism"][:230])
    print("   ", {k: round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
except Exception as ex:
    print("$tag FAILED", ex)
    import subprocess; print(subprocess.run("grep -v Warning gpurun_out/scale_${N}_${tag}.log | tail -25", shell=True, capture_output=True, text=True).stdout[-3000:])
PY
done
