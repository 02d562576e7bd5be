mkdir -p gpurun_out
for v in $VARIANTS; do
export CWSL_B200_LIB=$PWD/build/libcwsl_$v.so
PROBE_KINDS=bench timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe43_$v.log 2>&1
echo "$v 1024: $(grep -E 'stft_raw' gpurun_out/r2_probe43_$v.log | cut -c40-140)"
PROBE_KINDS=bench PROBE_CHANNELS=64 timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe43_64_$v.log 2>&1
echo "$v 64: $(grep -E 'stft_raw' gpurun_out/r2_probe43_64_$v.log | cut -c40-140)"
done
CWSL_B200_LIB=$PWD/build/libcwsl_tg.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or stress_stft or streaming" 2>&1 | tail -2
