#!/bin/sh
# Final verification of round 2 under gpurun (one GPU): GPU tests, smoke, the default bench line, the ncu launch list.
TAG=${1:-r02z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_gputests.log; cat gpurun_out/${TAG}_gputests.log
python __graft_entry__.py smoke 2>&1 | tail -4 > gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
tail -c 200 gpurun_out/${TAG}_launches.log
