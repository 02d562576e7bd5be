set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "stft or golden" 2>&1 | tail -3
PROBE_KINDS=bench timeout 600 python tools/r2_probe.py > gpurun_out/r2_probe14.log 2>&1
grep -E "kernel_ms" gpurun_out/r2_probe14.log | cut -c1-200
