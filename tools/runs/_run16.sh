set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
date +%s
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest16.log
tail -5 gpurun_out/r2_pytest16.log
date +%s
timeout 1200 python bench.py > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err
tail -c 1500 gpurun_out/r2_bench16.json; tail -3 gpurun_out/r2_bench16.err
date +%s
PROF_MODE=stft PROF_SLOTS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_chan -c 1 -s 1 -o gpurun_out/r2_chan_16 --force-overwrite python tools/profile_target.py > gpurun_out/ncu_chan16.log 2>&1
tail -3 gpurun_out/ncu_chan16.log
date +%s
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches16.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench16_ncu.log 2>&1
tail -5 gpurun_out/r2_launches16.csv | cut -c1-200
date +%s
