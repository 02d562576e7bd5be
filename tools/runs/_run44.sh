mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
S=gpurun_out/r2_compute_sanitizer.txt
echo "compute-sanitizer, round 2 final (B200, tools/sanitizer_target.py): FAST, EXACT, STFT channelizer (TMA-staged IQ ring, setmaxnreg role split, two hand-over groups on named barriers), STFT guard incl. the indirect FAST redo from anchors, 96/48 kHz geometries" > $S
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> $S
  timeout 900 compute-sanitizer --tool $tool python tools/sanitizer_target.py > gpurun_out/san_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum|done" gpurun_out/san_$tool.log >> $S
done
grep -E "SUMMARY" $S
timeout 900 python bench.py > gpurun_out/r2_bench44.json 2> gpurun_out/r2_bench44.err
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench44.json') if l.startswith('{')][-1];print('bench', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['launch_ms'],4), d['roofline']['frac'], d['e2e']['value'], {k:round(v['value']) for k,v in d['other_modes'].items()}, d['clocks'])"
