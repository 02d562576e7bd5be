set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2_pytest5.log
tail -15 gpurun_out/r2_pytest5.log
PROBE_KINDS=bench,stress16 timeout 600 python tools/r2_probe.py > gpurun_out/r2_probe5.log 2>&1
grep -E "kernel_ms" gpurun_out/r2_probe5.log | cut -c1-420
PROF_MODE=stft PROF_SLOTS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_chan -c 1 -s 1 -o gpurun_out/r2_chan_quad --force-overwrite python tools/profile_target.py > gpurun_out/ncu_chan5.log 2>&1
tail -3 gpurun_out/ncu_chan5.log
