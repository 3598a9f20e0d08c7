#!/bin/bash
# round 2, call U (8 GPUs): dataset generation with production-like case lengths (3000 steps), staggered and unstaggered starts; the distributed parity tests on real devices (NCCL incl. 4 ranks)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q -rfEs -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/r2u_pytest.log; tail -2 gpurun_out/r2u_pytest.log
timeout 600 python dev/dataset_run.py --gpus 8 --seq-cases 4 --steps 3000 --stagger 1.1 > gpurun_out/r2u_dataset_stagger.json 2> gpurun_out/r2u_dataset.err; cut -c1-700 gpurun_out/r2u_dataset_stagger.json
timeout 600 python dev/dataset_run.py --gpus 8 --seq-cases 2 --steps 3000 --stagger 0 > gpurun_out/r2u_dataset_nostagger.json 2>> gpurun_out/r2u_dataset.err; cut -c1-700 gpurun_out/r2u_dataset_nostagger.json
tail -3 gpurun_out/r2u_dataset.err
