F="pk cell,mapped,tvb,minmax,kxrcf"
for tool in memcheck racecheck; do
  timeout 330 compute-sanitizer --tool $tool --print-limit 50 python scripts/sanitize_cases.py "$F" > gpurun_out/r02h_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_cases: ok" gpurun_out/r02h_sanitize_${tool}.log | tail -3
done
