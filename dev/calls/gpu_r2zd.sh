#!/bin/bash
# round 2, call ZD: every GPU test that runs FP32 (the precision whose default tile variants changed), on the tree as committed; smoke()
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 95 python -m pytest tests -m gpu -q -k "fp32" -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r2zd_pytest_fp32.log; tail -3 gpurun_out/r2zd_pytest_fp32.log
