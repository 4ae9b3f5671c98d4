#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_floor.py default 10=6 10=3 13=4 13=4,10=6 static=0 > gpurun_out/diag_floor.txt 2>&1
cat gpurun_out/diag_floor.txt
