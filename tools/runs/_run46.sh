mkdir -p gpurun_out
S=gpurun_out/r2_compute_sanitizer.txt
echo "compute-sanitizer, round 2 final (B200, tools/sanitizer_target.py): FAST, EXACT, STFT channelizer (TMA-staged IQ ring, setmaxnreg role split, two hand-over groups on named barriers), packed hand-off, STFT guard incl. the indirect FAST redo from anchors, 96/48 kHz geometries" > $S
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool" >> $S
  timeout 900 compute-sanitizer --tool $tool python tools/sanitizer_target.py > gpurun_out/san_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum|done" gpurun_out/san_$tool.log >> $S
done
grep -E "SUMMARY" $S
grep -E "Uninitialized|at 0x|by thread" gpurun_out/san_initcheck.log | head -20
