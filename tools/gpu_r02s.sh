#!/bin/bash
mkdir -p gpurun_out
bash tools/run_mg.sh 2 a2a:32:2 allgather:32:2 > gpurun_out/r02s_mg2.txt 2>&1
echo done
