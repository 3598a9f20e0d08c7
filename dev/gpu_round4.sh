#!/bin/bash
export LUW_VERBOSE=1
for v in 0 1 2; do LUW_TILE_VARIANT=$v LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py 1 0 2>&1 | grep -v "^[0-9]"; done
for v in 0 1 2; do LUW_TILE_VARIANT=$v QB_PRECS=1 timeout 600 python tests/quickbench_dev.py 2>&1 | grep "arith=1\|variant\|luw"; done
