#!/bin/bash
# GPU call B: parity tests with the two-pass kernel as default, then A/B of the tile variants on the bench workloads.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_B.log
for v in 0 1 3; do for w in urban_fp16s channel512_fp16s urban_fp16s_uf; do
LUW_VERBOSE=1 LUW_TILE_VARIANT=$v timeout 300 python bench.py --no-cpu --no-e2e --steps 60 --warmup 10 --workload $w 2> gpurun_out/err_${v}_$w.log | python -c "import json,sys; d=json.load(sys.stdin); print('v$v $w', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))" | tee -a gpurun_out/variants_B.txt
grep "luw" gpurun_out/err_${v}_$w.log | head -2
done; done
