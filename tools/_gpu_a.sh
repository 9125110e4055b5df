#!/bin/bash
cd /root/repo
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "match" 2>&1 | tail -5 > gpurun_out/c_tests.log
for v in "" "RTR_MATCH_KEEP=8" "RTR_MATCH_KEEP=16"; do
  env $v timeout 300 python tools/bench_match_small.py >> gpurun_out/c_match_small.log 2>&1
done
timeout 900 python tools/bench_match_scale.py --check 256 --variants "RTR_MATCH_CLUSTER=1" > gpurun_out/c_match_scale.log 2>&1
timeout 900 python -m pytest tests/test_gpu_many.py tests/test_gpu_parity.py -x -q -k "register" 2>&1 | tail -5 > gpurun_out/c_tests_reg.log
echo done
