#!/bin/bash
cd /root/repo
export PYTHONUNBUFFERED=1
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "wide_hierarchy or register_prepared or icp_cap_sized or icp_parity" > gpurun_out/k_memcheck.log 2>&1
echo "rc $?" >> gpurun_out/k_memcheck.log
RTR_MATCH_TC=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "wide_hierarchy" > gpurun_out/k_racecheck.log 2>&1
echo "rc $?" >> gpurun_out/k_racecheck.log
tail -4 gpurun_out/k_memcheck.log gpurun_out/k_racecheck.log
