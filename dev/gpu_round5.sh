#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
export LUW_VERBOSE=1
TRACE_TILES_X=4 LUW_TILE_VARIANT=0 LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py 1 0 2>&1
TRACE_TILES_X=2 LUW_TILE_VARIANT=1 LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py 1 0 2>&1
TRACE_TILES_X=8 LUW_TILE_VARIANT=2 LUW_CUDA_LIB=latticeurbanwind_b200/lib_trace/libluw_cuda.so python dev/trace_tiles.py 1 0 2>&1
unset LUW_VERBOSE
for v in 0 1 2; do LUW_TILE_VARIANT=$v QB_PRECS=1,0 timeout 600 python tests/quickbench_dev.py 2>&1 | grep "arith=1\|variant"; done
