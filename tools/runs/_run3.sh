set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest3.log
tail -15 gpurun_out/r2_pytest3.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
tail -c 3000 gpurun_out/r2_bench3.json
tail -5 gpurun_out/r2_bench3.err
