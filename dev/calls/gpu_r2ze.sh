#!/bin/bash
# round 2, call ZE (2 GPUs): the multi-rank end-to-end leg of bench.py (per-step boundary upload and probe read-back on every rank) on a small workload
mkdir -p gpurun_out
timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 60 --warmup 10 --workload profile256_fp16s > gpurun_out/r2ze_n2_e2e.json 2> gpurun_out/r2ze_n2_e2e.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2ze_n2_e2e.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'e2e', d['e2e'])
except Exception as e:
    print('no line:', e); print(open('gpurun_out/r2ze_n2_e2e.err').read()[-1500:])
PY
