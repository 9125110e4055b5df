#!/bin/bash
cd /root/repo
export PYTHONUNBUFFERED=1
for f in 0.7 1.0 1.4; do
RTR_ICP_CELL_FACTOR=$f timeout 300 python tools/bench_icp.py --reps 3 >> gpurun_out/f_icp_cell.log 2>&1
done
echo done
