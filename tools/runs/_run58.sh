mkdir -p gpurun_out
PROF_MODE=stft PROF_SLOTS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:quantise -c 1 -s 1 -o gpurun_out/r2_quant4 --force-overwrite python tools/profile_target.py > gpurun_out/ncu_quant58.log 2>&1
tail -1 gpurun_out/ncu_quant58.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"demod|quantise|guard|clear_u32|phase" -c 400 --csv --log-file gpurun_out/r2_launches58.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench58_ncu.log 2>&1
tail -2 gpurun_out/r2_launches58.csv | cut -c1-200
