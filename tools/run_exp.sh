for c in 8 16 32; do
  export KA9Q_MC_CTAS=$c
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29560+c)) bench.py --gpus 2 --steps 10 --warmup 3 --mg-mode spectrum > gpurun_out/mc_$c.log 2>&1
  grep '^{' gpurun_out/mc_$c.log | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ctas $c value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k: round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done
