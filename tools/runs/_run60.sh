mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_proxy.py -m gpu -q -x -s -k "jt65" 2>&1 | grep -v "^$" | tail -8 > gpurun_out/r2_pytest60.log; cat gpurun_out/r2_pytest60.log
