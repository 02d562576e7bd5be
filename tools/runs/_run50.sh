mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_wspr_gpu.py tests/test_host_gpu.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_pytest50.log
tail -25 gpurun_out/r2_pytest50.log
