#!/bin/bash
mkdir -p gpurun_out
for k in 1 2 3; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --filter-h-blocks-per-write $k > gpurun_out/fh_k$k.json 2> gpurun_out/fh_k$k.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/fh_k$k.json").read().strip().splitlines()[-1])
f=d["e2e"]["filter_h"]; print("k=$k", round(f["value"]), f["ms_per_block"], f["latency_ms_mean"], f["latency_ms_max"], f["dropped_blocks"], f.get("parity"), round(f["with_ring_memcpy"]["value"]))
PY
done
