#!/bin/bash
# round 2, call ZA: tile variants on example-sized lattices (FP16C / FP16S LES) and on a large FP32 LES lattice
mkdir -p gpurun_out
: > gpurun_out/r2za_sweeps.txt
timeout 80 python dev/variant_sweep.py profile256_fp16c 0,4,5,6,7,d 300 30 >> gpurun_out/r2za_sweeps.txt 2> gpurun_out/r2za.err
timeout 80 python dev/variant_sweep.py profile256_fp16s 0,4,5,6,7,d 300 30 >> gpurun_out/r2za_sweeps.txt 2>> gpurun_out/r2za.err
timeout 120 python dev/variant_sweep.py urban512_fp32 0,1,5,d 60 10 >> gpurun_out/r2za_sweeps.txt 2>> gpurun_out/r2za.err
cat gpurun_out/r2za_sweeps.txt; tail -3 gpurun_out/r2za.err
