#!/bin/sh
# Run under gpurun (one GPU). Produces the launch list and one full capture of the pair kernel for bench.py's workload.
#   sh profiles/run_ncu.sh <tag> [particles] [extra bench args]
TAG=${1:-r01}; N=${2:-1000000}; shift; shift
mkdir -p gpurun_out
BENCH="python bench.py --particles $N --steps 2 --warmup 3 --no-e2e --no-cpu-baseline $*"
# every launch of two timed steps with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
# the dominant kernel, full set, source-level
ncu --set full --clock-control none --import-source on -k regex:k_pair_sum -s 1 -c 1 -f -o gpurun_out/${TAG}_pair $BENCH > gpurun_out/${TAG}_pair.log 2>&1
ls -la gpurun_out/${TAG}_*
