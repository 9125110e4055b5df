#!/bin/bash
cd /root/repo
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/h_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_ref.json 2> gpurun_out/h_bench_ref.err
echo done
