#!/bin/bash
# round 2, call M: full GPU suite after the edge-warp / FP16C-default / TYPE_E changes; thermal and FP16C workloads with the new defaults
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1800 python -m pytest tests -m gpu -q -rfE -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r2m_pytest.log
tail -5 gpurun_out/r2m_pytest.log
for w in urban_fp16s_thermal urban_fp16c_thermal urban_fp16c_uf; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e --traffic off --also '' --sustain 0 2>> gpurun_out/r2m_sweep.err | tee -a gpurun_out/r2m_bench.txt | cut -c1-120
done
for w in urban_fp16s:d channel512_fp16c:d channel512_fp16s:d; do
  timeout 300 python dev/variant_sweep.py ${w%%:*} ${w##*:} 40 10 2>> gpurun_out/r2m_sweep.err | tee -a gpurun_out/r2m_sweep.txt
done
grep -E "^\|  [0-9]+ " gpurun_out/reference_driver.log | tail -2
