#!/bin/sh
# Run under gpurun (one GPU): full ncu capture of the two pair kernels for bench.py's workload (k_pair_sum: second launch,
# particles already moving; k_pair_lists: first launch = a list build, later launches return at once while the lists are reused).
#   sh profiles/run_ncu2.sh <tag> [particles] [extra bench args]
TAG=${1:-r02}; N=${2:-10000000}; shift; shift
mkdir -p gpurun_out
BENCH="python bench.py --particles $N --steps 2 --warmup 3 --no-e2e --no-cpu-baseline $*"
ncu --set full --clock-control none --import-source on -k regex:k_pair_sum -s 1 -c 1 -f -o gpurun_out/${TAG}_pair $BENCH > gpurun_out/${TAG}_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_lists -s 0 -c 1 -f -o gpurun_out/${TAG}_lists $BENCH > gpurun_out/${TAG}_lists.log 2>&1
ls -la gpurun_out/${TAG}_*
