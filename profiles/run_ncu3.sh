#!/bin/sh
# Run under gpurun (one GPU): full ncu capture of the kernels beside the pair stage (solid bench workload) and of the
# fluid configuration's pair kernels.   sh profiles/run_ncu3.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'k_prologue_pack|k_predict|k_correct|k_criteria|k_bounds' -s 5 -c 5 -f -o gpurun_out/${TAG}_rest $B --particles 10000000 > gpurun_out/${TAG}_rest.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pair_sum|k_prologue_pack' -s 2 -c 2 -f -o gpurun_out/${TAG}_fluid $B --fluid --particles 12500000 > gpurun_out/${TAG}_fluid.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_lists -s 0 -c 1 -f -o gpurun_out/${TAG}_fluid_lists $B --fluid --particles 12500000 > gpurun_out/${TAG}_fluid_lists.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_lists -s 0 -c 1 -f -o gpurun_out/${TAG}_lists $B --particles 10000000 > gpurun_out/${TAG}_lists.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_fluid_launches.csv $B --fluid --particles 12500000 > gpurun_out/${TAG}_fluid_launches.log 2>&1
ls -la gpurun_out/${TAG}_*
