#!/bin/bash
cd /root/repo
export PYTHONUNBUFFERED=1
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m_scale_$N.json 2> gpurun_out/m_scale_$N.err
echo "rc $?" >> gpurun_out/m_scale_$N.err
tail -n 3 gpurun_out/m_scale_$N.err
