#!/bin/bash
for v in 0 1 2; do for w in channel512_fp16s urban_fp16s urban_fp16s_uf channel512_fp32; do LUW_TILE_VARIANT=$v timeout 300 python bench.py --no-cpu --no-e2e --steps 60 --warmup 10 --workload $w | python -c "import json,sys; d=json.load(sys.stdin); print('v$v $w', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"; done; done
