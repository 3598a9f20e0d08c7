#!/bin/bash
# round 2, call J (8 GPUs): weak scaling at 8 GPUs with the README layouts under `also`, the 10 G-cell configuration, dataset generation as replicas with the real driver,
# the reference driver with n_gpu = [2, 2, 2]
mkdir -p gpurun_out
tr() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 "$@"; }
tr --steps 60 --warmup 10 --workload urban_fp16s > gpurun_out/r2j_n8_urban.json 2> gpurun_out/r2j_n8_urban.err
tr --steps 100 --warmup 10 --workload channel512_fp16s > gpurun_out/r2j_n8_channel.json 2> gpurun_out/r2j_n8_channel.err
tr --steps 20 --warmup 5 --workload city10g_fp16s --decomp 2,2,2 > gpurun_out/r2j_n8_city10g_222.json 2> gpurun_out/r2j_n8_city10g_222.err
python - <<'PY'
import json
for n in ('urban','channel','city10g_222'):
    try:
        d=json.load(open(f'gpurun_out/r2j_n8_{n}.json'))
        print(n, d['config']['decomposition'], d['config']['lattice'], round(d['value']), 'ms', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'exposed', round(d['halo']['exposed_ms_per_step'],4), d['halo'].get('overlapped_with_the_step'))
        for a in d.get('also',[]): print('   also', a['decomposition'], round(a['value']), round(a['ms_per_step'],3), round(a['halo']['exposed_ms_per_step'],4))
    except Exception as e: print(n,'failed',e)
PY
timeout 300 python bench.py --workload urban_fp16s --steps 60 --warmup 10 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2j_n1_urban.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/r2j_n1_urban.json')); print('N=1 urban', round(d['value']), round(d['ms_per_step'],3))"
timeout 900 python dev/dataset_run.py --gpus 8 --seq-cases 4 > gpurun_out/r2j_dataset.json 2> gpurun_out/r2j_dataset.err
cat gpurun_out/r2j_dataset.json | cut -c1-900; tail -3 gpurun_out/r2j_dataset.err
rm -rf /tmp/case222; cp -r baseline/_ref/case_profile /tmp/case222
(cd /tmp/case222 && LUW_VERBOSE=1 timeout 600 /root/repo/baseline/_ref/luw_reference_driver /tmp/case222/conf_222.luwpf < /dev/null > /root/repo/gpurun_out/r2j_driver_222.log 2> /root/repo/gpurun_out/r2j_driver_222.err; echo "driver [2,2,2] exit $?")
grep -E "Grid Resolution|MLUPs|Task finished|GPU Estimate|Device" gpurun_out/r2j_driver_222.log | head -12 | cut -c1-150
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
