#!/bin/bash
mkdir -p gpurun_out
MG_EXTRA=--quick bash tools/run_mg.sh 8 a2a:32:2 > gpurun_out/r02w_mg8.txt 2>&1
MG_TAG=_full bash tools/run_mg.sh 8 a2a:32:2 >> gpurun_out/r02w_mg8.txt 2>&1
echo done
