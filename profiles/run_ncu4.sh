#!/bin/sh
# Run under gpurun (one GPU): full ncu captures of the optional-term variants of k_pair_sum and of the component sweep at the
# bench size.   sh profiles/run_ncu4.sh <tag>
TAG=${1:-r02t}
mkdir -p gpurun_out
B="python bench.py --particles 10000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_pair_sum -s 1 -c 1 -f -o gpurun_out/${TAG}_stressav $B --no-gravity --terms stressav > gpurun_out/${TAG}_stressav.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_sum -s 1 -c 1 -f -o gpurun_out/${TAG}_deltasph $B --no-gravity --terms deltasph > gpurun_out/${TAG}_deltasph.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_comp_sweep -s 4 -c 1 -f -o gpurun_out/${TAG}_comp python profiles/run_components_ncu.py > gpurun_out/${TAG}_comp.log 2>&1
ls -la gpurun_out/${TAG}_*
