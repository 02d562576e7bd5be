set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest13.log
tail -5 gpurun_out/r2_pytest13.log
PROBE_KINDS=bench,stress16 timeout 600 python tools/r2_probe.py > gpurun_out/r2_probe13.log 2>&1
grep -E "kernel_ms" gpurun_out/r2_probe13.log | cut -c1-300
timeout 1200 python bench.py > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench9.err
tail -3 gpurun_out/r2_bench13.err
