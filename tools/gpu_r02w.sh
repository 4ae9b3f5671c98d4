#!/bin/bash
# multi-GPU line of the default hand-off (run under gpurun --gpus N): tools/run_mg.sh N a2a:32:2
mkdir -p gpurun_out
N=${MG_N:-8}
MG_TAG=_final bash tools/run_mg.sh $N a2a:32:2 > gpurun_out/r02w_mg${N}_final.txt 2>&1
cat gpurun_out/r02w_mg${N}_final.txt
echo done
