#!/bin/bash
mkdir -p gpurun_out
bash tools/run_mg.sh 8 allgather:32:2 spectrum:32:2 slices:32:2 > gpurun_out/r02o_mg8.txt 2>&1
echo done
