mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "stft or golden or guard or stress or chan" 2>&1 | tail -3
PROBE_KINDS=bench timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe28.log 2>&1
echo "probe: $(grep -E 'stft_raw' gpurun_out/r2_probe28.log | cut -c1-150)"
PROBE_KINDS=bench PROBE_CHANNELS=64 timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe28_64.log 2>&1
echo "probe64: $(grep -E 'stft_raw' gpurun_out/r2_probe28_64.log | cut -c1-150)"
