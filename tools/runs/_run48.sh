mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
for nrw in 1 0 1 0; do
CWSL_QUANT_NARROW=$nrw timeout 300 python bench.py --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench48_$nrw.json 2>/dev/null
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench48_$nrw.json') if l.startswith('{')][-1];print('narrow=$nrw bench', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['launch_ms'],4), d['kernel_ms']['isolated_per_receiver']['quantise_and_clear_ms'])"
done
