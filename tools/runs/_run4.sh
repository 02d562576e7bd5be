set -x
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
tail -c 1500 gpurun_out/r2_bench4.json
tail -5 gpurun_out/r2_bench4.err
timeout 300 python tools/numa_probe.py > gpurun_out/r2_numa_n1.log 2>&1
tail -3 gpurun_out/r2_numa_n1.log
