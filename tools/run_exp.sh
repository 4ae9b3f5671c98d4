p=29600
for ch in 8 16; do for mode in spectrum input; do
  p=$((p+1))
  NCCL_MAX_NCHANNELS=$ch NCCL_MIN_NCHANNELS=$ch timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 10 --warmup 3 --mg-mode $mode > gpurun_out/nccl_${ch}_$mode.log 2>&1
  grep '^{' gpurun_out/nccl_${ch}_$mode.log | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('channels $ch $mode value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), {k: round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done; done
