mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tests/sanitizer_smoke.py > gpurun_out/r02b_san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SMOKE DONE|hazard|Invalid" gpurun_out/r02b_san_$tool.log | sort | uniq -c | tail -6
done
