mkdir -p gpurun_out
PROBE_KINDS=bench timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe45.log 2>&1
echo "probe 1024: $(grep -E 'stft_raw' gpurun_out/r2_probe45.log | cut -c40-140)"
PROF_MODE=stft PROF_SLOTS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_chan -c 1 -s 1 -o gpurun_out/r2_chan_final --force-overwrite python tools/profile_target.py > gpurun_out/ncu_chan45.log 2>&1
tail -1 gpurun_out/ncu_chan45.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"demod|quantise|guard|clear_u32|phase" -c 400 --csv --log-file gpurun_out/r2_launches45.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench45_ncu.log 2>&1
tail -2 gpurun_out/r2_launches45.csv | cut -c1-200
