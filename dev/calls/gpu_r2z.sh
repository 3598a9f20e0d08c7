#!/bin/bash
# round 2, call Z: FP32 LES step (C1-sized lattice) per tile variant: single pass (V0, the FP32 default), two-pass (V1), lean loop (V5)
mkdir -p gpurun_out
timeout 150 python dev/variant_sweep.py profile256_fp32 0,1,5,d 300 30 > gpurun_out/r2z_sweep_c1_fp32.txt 2> gpurun_out/r2z_sweep.err; echo "rc=$?"; cat gpurun_out/r2z_sweep_c1_fp32.txt; tail -3 gpurun_out/r2z_sweep.err
