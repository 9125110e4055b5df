#!/bin/bash
cd /root/repo
export PYTHONUNBUFFERED=1
# 1. launch list of the bench command
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/l_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --no-sustained > gpurun_out/l_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/l_launches.csv gpurun_out/l_launches_summary.md > /dev/null 2>&1
gzip -f gpurun_out/l_launches.csv
# 2. full capture of one online step (prepared database)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/prepared python tools/run_many.py --prepared --steps 1 > gpurun_out/l_prepared_ncu.log 2>&1
ncu -i /tmp/prepared.ncu-rep --page raw --csv > gpurun_out/l_prepared_raw.csv 2>/dev/null
# 3. the bench lines
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/l_bench_ref.json 2> gpurun_out/l_bench_ref.err
echo done
