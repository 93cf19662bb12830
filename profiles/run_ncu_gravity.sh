#!/bin/sh
# Run under gpurun (one GPU): full ncu capture of the gravity walk and moments kernels at 1.06 M particles.
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_grav_walk|k_grav_moments|k_grav_tree' -s 0 -c 3 -f -o gpurun_out/${TAG}_gravity python profiles/run_gravity.py 1000000 > gpurun_out/${TAG}_gravity_ncu.log 2>&1
python profiles/ncu_summary.py gpurun_out/${TAG}_gravity.ncu-rep > gpurun_out/${TAG}_gravity_ncu_summary.csv
head -60 gpurun_out/${TAG}_gravity_ncu_summary.csv
