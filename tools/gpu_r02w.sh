#!/bin/bash
mkdir -p gpurun_out
N=${MG_N:-4}
MG_TAG=_full bash tools/run_mg.sh $N a2a:32:2 > gpurun_out/r02w_mg$N.txt 2>&1
cat gpurun_out/r02w_mg$N.txt
echo done
