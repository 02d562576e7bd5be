mkdir -p gpurun_out
S=gpurun_out/r2_compute_sanitizer.txt
echo "compute-sanitizer, round 2 (B200, tools/sanitizer_target.py): FAST, EXACT, STFT channelizer (TMA-staged IQ ring, setmaxnreg role split), STFT guard incl. the indirect FAST redo from anchors, 96/48 kHz geometries" > $S
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> $S
  timeout 900 compute-sanitizer --tool $tool python tools/sanitizer_target.py > gpurun_out/san_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum|done" gpurun_out/san_$tool.log >> $S
  grep -E "=========" gpurun_out/san_$tool.log | grep -v SUMMARY | head -20 >> $S
done
cat $S
