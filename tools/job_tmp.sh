for t in memcheck racecheck synccheck initcheck; do
  echo "== $t"; timeout 1200 compute-sanitizer --tool $t python tools/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|step loss|modes" | sort | uniq -c | head -12
done > gpurun_out/r2G_sanitizer.txt 2>&1
cat gpurun_out/r2G_sanitizer.txt
