#!/bin/bash
mkdir -p gpurun_out
bash tools/run_mg.sh 2 spectrum:32:2 allgather:32:2 slices:32:2 > gpurun_out/r02n_mg2_hipri.txt 2>&1
KA9Q_NCCL_HIPRI=0 bash tools/run_mg.sh 2 spectrum:32:2 > gpurun_out/r02n_mg2_lopri.txt 2>&1
echo done
