#!/bin/bash
# round 2, call I (2 GPUs): overlapped halo exchange across two devices (IPC over NVLink), 2-rank parity tests, weak-scaling lines with the overlap on and off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q -rfEs -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2i_pytest.log
tail -4 gpurun_out/r2i_pytest.log
# the reference API's multi-GPU path on physical devices: the C++ LBM (one thread, D domains, device d % ndev) with the direct (remote-store) in-process exchange
timeout 900 python -m pytest tests/test_cpp_host.py tests/test_gpu_parity.py tests/test_voxelize.py tests/test_thermal_gpu.py tests/test_stats.py -m gpu -q -rfEs -p no:cacheprovider -k "decomposed or domains or cpp or 2x" 2>&1 | tail -12 > gpurun_out/r2i_pytest_inprocess.log
tail -4 gpurun_out/r2i_pytest_inprocess.log
LUW_HALO_DIRECT=0 timeout 900 python -m pytest tests/test_cpp_host.py -m gpu -q -rfEs -p no:cacheprovider -k "decomposed or domains or 2x" 2>&1 | tail -4 > gpurun_out/r2i_pytest_inprocess_staged.log
tail -2 gpurun_out/r2i_pytest_inprocess_staged.log
run() { # name, env, args
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 $3 > gpurun_out/r2i_$1.json 2> gpurun_out/r2i_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2i_$1.json'))
    print('$1', d['config']['decomposition'], round(d['value']), 'ms', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'exposed', round(d['halo']['exposed_ms_per_step'],4), 'overlapped', d['halo'].get('overlapped_with_the_step'))
except Exception as e: print('$1 failed', e)
PY
}
run urban_on LUW_X=1 "--workload urban_fp16s"
run urban_off LUW_HALO_OVERLAP=0 "--workload urban_fp16s"
run channel_z_on LUW_X=1 "--workload channel512_fp16s"
run channel_z_off LUW_HALO_OVERLAP=0 "--workload channel512_fp16s"
run channel_y_on LUW_X=1 "--workload channel512_fp16s --decomp 1,2,1"
run channel_x LUW_X=1 "--workload channel512_fp16s --decomp 2,1,1"
timeout 300 python bench.py --workload urban_fp16s --steps 100 --warmup 10 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2i_urban_n1.json 2>/dev/null
timeout 300 python bench.py --workload channel512_fp16s --steps 100 --warmup 10 --no-cpu --no-e2e --traffic off --also '' --sustain 0 > gpurun_out/r2i_channel_n1.json 2>/dev/null
python - <<'PY'
import json
for n in ('urban','channel'):
    d=json.load(open(f'gpurun_out/r2i_{n}_n1.json')); print(n,'N=1', round(d['value']), round(d['ms_per_step'],4))
PY
