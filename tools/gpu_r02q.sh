#!/bin/bash
mkdir -p gpurun_out
export MG_EXTRA=--quick
MG_TAG=_nvls NCCL_ALGO=NVLS bash tools/run_mg.sh 8 allgather:32:2 > gpurun_out/r02q_mg8.txt 2>&1
MG_TAG=_ch32 NCCL_MIN_NCHANNELS=32 bash tools/run_mg.sh 8 allgather:32:2 >> gpurun_out/r02q_mg8.txt 2>&1
MG_TAG=_d3 bash tools/run_mg.sh 8 allgather:32:3 slices:32:2 >> gpurun_out/r02q_mg8.txt 2>&1
echo done
