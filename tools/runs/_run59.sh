mkdir -p gpurun_out
S=gpurun_out/r2_compute_sanitizer_final.txt
echo "compute-sanitizer on the code as shipped at the end of round 2 (quantise pass with the branch-free interior path, 4 groups per thread; everything else as in r2_compute_sanitizer.txt): tools/sanitizer_target.py" > $S
for tool in memcheck initcheck; do
  echo "== $tool" >> $S
  timeout 700 compute-sanitizer --tool $tool python tools/sanitizer_target.py > gpurun_out/san2_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum|done" gpurun_out/san2_$tool.log >> $S
done
grep -E "SUMMARY" $S
