#!/bin/bash
# 2 GPUs: everything that needs two devices once more with the final library, and the N=2 lines for the record
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "decomposed or nccl or halo or full_size" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_U.log
run() { # workload decomp
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --workload $1 $2 > gpurun_out/bench_n2_U.json 2> gpurun_out/bench_n2_U.err
echo "exit=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n2_U.err | tail -3
python -c "import json; d=json.loads(open('gpurun_out/bench_n2_U.json').read().strip().splitlines()[-1]); print('N2 $1 $2', d['config']['decomposition'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['halo'], d['gpu_launches'])" | tee -a gpurun_out/n2_U.txt
cp gpurun_out/bench_n2_U.json gpurun_out/bench_n2_final_$3.json
}
run channel512_fp16s "" channel
run urban_fp16s "" urban
run channel512_fp16s "--decomp 2,1,1" channel_xsplit
timeout 300 python bench.py --no-cpu --no-e2e --steps 100 --warmup 10 --workload channel512_fp16s | python -c "import json,sys; d=json.load(sys.stdin); print('N1 channel', round(d['value']), round(d['ms_per_step'],3))" | tee -a gpurun_out/n2_U.txt
