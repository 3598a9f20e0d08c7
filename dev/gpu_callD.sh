#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fast or tiled" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_D.log
run() { # lib variant workload
LUW_CUDA_LIB=$1 LUW_VERBOSE=1 LUW_TILE_VARIANT=$2 timeout 300 python bench.py --no-cpu --no-e2e --steps 60 --warmup 10 --workload $3 2> gpurun_out/err.log | python -c "import json,sys; d=json.load(sys.stdin); print('$1 v$2 $3', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))" | tee -a gpurun_out/variants_D.txt
grep "luw" gpurun_out/err.log | head -1 | cut -c1-110
}
for v in 0 1 3 4; do for w in urban_fp16s channel512_fp16s; do run latticeurbanwind_b200/lib/libluw_cuda.so $v $w; done; done
for v in 1 4; do for w in urban_fp16s channel512_fp16s; do run latticeurbanwind_b200/lib_nb/libluw_cuda.so $v $w; done; done
run latticeurbanwind_b200/lib/libluw_cuda.so 1 urban_fp16s_nz
run latticeurbanwind_b200/lib/libluw_cuda.so 4 urban_fp16s_uf
run latticeurbanwind_b200/lib/libluw_cuda.so 1 urban_fp16s_uf
