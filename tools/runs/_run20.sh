mkdir -p gpurun_out
for v in low high low high; do
export CWSL_B200_LIB=$PWD/build/libcwsl_$v.so
PROBE_KINDS=bench timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe20_$v.log 2>&1
echo "$v probe: $(grep -E 'stft_raw' gpurun_out/r2_probe20_$v.log | cut -c1-150)"
done
unset CWSL_B200_LIB
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "stft or golden or guard or stress or chan" 2>&1 | tail -3
