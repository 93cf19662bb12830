mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  SMOKE_ONLY=deltasph timeout 600 compute-sanitizer --tool $tool python tests/sanitizer_smoke.py > gpurun_out/r02c_san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SMOKE DONE|hazard|Invalid|delta-SPH|artificial stress" gpurun_out/r02c_san_$tool.log | sort | uniq -c | tail -12
done
