#!/bin/sh
# Run under gpurun (one GPU): the default bench line, other sizes, the ncu launch list.
#   sh profiles/run_final.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench_1gpu.json
python bench.py --particles 1000000 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_1M_1gpu.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --particles 100000 --steps 100 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_100k_1gpu.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --fluid --particles 12500000 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_fluid_12M_1gpu.json 2>> gpurun_out/${TAG}_bench.err
for f in 1M 100k fluid_12M; do python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${f}_1gpu.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['value'], d['config']['list_reuse']['builds_in_timed_steps'])
"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
tail -c 300 gpurun_out/${TAG}_launches.log
