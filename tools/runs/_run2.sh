set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
PROBE_KINDS=bench,stress16 timeout 600 python tools/r2_probe.py > gpurun_out/r2_probe2.log 2>&1
grep -E "kernel_ms" gpurun_out/r2_probe2.log | cut -c1-420
PROF_MODE=stft PROF_SLOTS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_chan -c 1 -s 1 -o gpurun_out/r2_chan_tma --force-overwrite python tools/profile_target.py > gpurun_out/ncu_chan2.log 2>&1
tail -3 gpurun_out/ncu_chan2.log
PROF_MODE=stft PROF_SLOTS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_stft_probe.csv python tools/profile_target.py > /dev/null 2>&1
tail -12 gpurun_out/r2_launches_stft_probe.csv | cut -c1-200
