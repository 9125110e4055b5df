#!/bin/bash
cd /root/repo
export PYTHONUNBUFFERED=1
N=$1
export RTR_COMM_VERBOSE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/check_comm.py > gpurun_out/p_check_$N.log 2>&1
echo "rc $?" >> gpurun_out/p_check_$N.log
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/p_check_$N.log | tail -n 6
