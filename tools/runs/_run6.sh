set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "stft or stress or chan or guard" 2>&1 | tail -15 > gpurun_out/r2_pytest8.log
tail -8 gpurun_out/r2_pytest8.log
PROBE_KINDS=bench timeout 600 python tools/r2_probe.py > gpurun_out/r2_probe8.log 2>&1
grep -E "kernel_ms" gpurun_out/r2_probe8.log | cut -c1-300
PROF_MODE=stft PROF_SLOTS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_chan -c 1 -s 1 -o gpurun_out/r2_chan_quad4 --force-overwrite python tools/profile_target.py > gpurun_out/ncu_chan8.log 2>&1
tail -3 gpurun_out/ncu_chan8.log
