#!/bin/bash
mkdir -p gpurun_out
bash tools/run_mg.sh 2 spectrum:32:2 spectrum:16:3 spectrum:8:4 allgather:32:2 slices:32:2 input:32:2 > gpurun_out/r02j_mg2.txt 2>&1
echo done
