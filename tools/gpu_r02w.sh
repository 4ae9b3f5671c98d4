#!/bin/bash
mkdir -p gpurun_out
N=${MG_N:-4}
MG_EXTRA=--quick bash tools/run_mg.sh $N a2a:32:3 a2a:32:4 a2a:64:2 > gpurun_out/r02w_mg${N}_depth.txt 2>&1
cat gpurun_out/r02w_mg${N}_depth.txt
echo done
