mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_host_gpu.py -m gpu -q -x -k "packed or different_length or host or station or wav or shm" 2>&1 | tail -6
timeout 600 python bench.py --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench38.json 2> gpurun_out/r2_bench38.err
tail -2 gpurun_out/r2_bench38.err
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench38.json') if l.startswith('{')][-1];print('bench', round(d['value']), round(d['ms_per_step'],2), 'e2e', d['e2e']['value'], d['e2e']['d2h_gbs_per_gpu'], d['e2e']['ms_per_step'])"
