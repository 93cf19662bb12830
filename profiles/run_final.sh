#!/bin/sh
# Run under gpurun (one GPU): compute-sanitizer smoke, the default bench line, the ncu launch list and the full captures.
#   sh profiles/run_final.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tests/sanitizer_smoke.py > gpurun_out/${TAG}_san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "SUMMARY|SMOKE DONE" gpurun_out/${TAG}_san_$tool.log | tail -3
done
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
sh profiles/run_ncu2.sh ${TAG} 10000000
