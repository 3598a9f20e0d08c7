#!/bin/bash
# round 2, call N (8 GPUs): weak scaling with the edge-warp change (x-decomposed layouts no longer send whole first / last tiles through the general body)
mkdir -p gpurun_out
tr() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 "$@"; }
tr --steps 60 --warmup 10 --workload urban_fp16s > gpurun_out/r2n_n8_urban.json 2> gpurun_out/r2n_n8_urban.err
tr --steps 100 --warmup 10 --workload channel512_fp16s > gpurun_out/r2n_n8_channel.json 2> gpurun_out/r2n_n8_channel.err
timeout 300 python bench.py --workload urban_fp16s --steps 60 --warmup 10 --no-cpu --no-e2e --traffic off --also 'channel512_fp16s' --sustain 0 > gpurun_out/r2n_n1.json 2>/dev/null
python - <<'PY'
import json
for n in ('urban','channel'):
    try:
        d=json.load(open(f'gpurun_out/r2n_n8_{n}.json'))
        print(n, d['config']['decomposition'], round(d['value']), 'ms', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'exposed', round(d['halo']['exposed_ms_per_step'],4), d['halo'].get('overlapped_with_the_step'))
        for a in d.get('also',[]): print('   also', a['decomposition'], round(a['value']), round(a['ms_per_step'],3), 'kernel', round(a['kernel_ms'],3), round(a['halo']['exposed_ms_per_step'],4))
    except Exception as e: print(n,'failed',e)
d=json.load(open('gpurun_out/r2n_n1.json')); print('N=1 urban', round(d['value']), 'channel', round(d['also'][0]['value']))
PY
