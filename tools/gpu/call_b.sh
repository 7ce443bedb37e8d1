#!/bin/bash
# ncu --set full of the CUDA-core kernels that sit far above their byte floor in the step's launch list
set -x
mkdir -p gpurun_out/r2c
SPECS="swgrad:64,32,32,2,64,3 sfwd:64,32,32,64,2,3 swgrad:64,128,128,128,2,1 sfwd:64,64,64,192,2,1 sfwd:64,32,32,2,64,3 swgrad:64,2,2,192,2,3 swgrad:64,16,16,2,64,3 sfwd:64,128,128,128,2,1"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2c/small_r02 python tools/ncu_shapes.py $SPECS > gpurun_out/r2c/ncu.log 2>&1
tail -3 gpurun_out/r2c/ncu.log
python tools/ncu_digest.py gpurun_out/r2c/small_r02.ncu-rep "$(echo $SPECS | sed 's/,/_/g; s/ /,/g')" --source 14 > gpurun_out/r2c/small_r02.txt 2>&1
ls -la gpurun_out/r2c/
head -60 gpurun_out/r2c/small_r02.txt
