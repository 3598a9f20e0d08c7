#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1 2; do LUW_TILE_VARIANT=$v QB_PRECS=1 timeout 600 python tests/quickbench_dev.py 2>&1 | grep "arith=1\|variant"; done
for v in 0 2; do LUW_TILE_VARIANT=$v timeout 300 python bench.py --no-cpu --no-e2e --steps 100 | python -c "import json,sys; d=json.load(sys.stdin); print('channel512_fp16s variant $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
LUW_TILE_VARIANT=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 100 --workload urban_fp16s | python -c "import json,sys; d=json.load(sys.stdin); print('urban', d['value'], d['ms_per_step'], d['roofline']['frac'])"
