set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_pytest1.log
cat gpurun_out/r2_pytest1.log | tail -30
timeout 900 python tools/r2_probe.py > gpurun_out/r2_probe1.log 2>&1
tail -30 gpurun_out/r2_probe1.log
