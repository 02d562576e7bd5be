set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "stft or golden or guard or stress or chan" 2>&1 | tail -6
PROBE_KINDS=bench,carrier80,two_strong,noise timeout 900 python tools/r2_probe.py > gpurun_out/r2_probe17.log 2>&1
grep -E "stft_raw|stft_guard" gpurun_out/r2_probe17.log | cut -c1-520
cp gpurun_out/r2_probe.json gpurun_out/r2_probe17.json
timeout 300 python bench.py --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench17.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r2_bench17.json'));print('bench', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['launch_ms'],4), d['roofline']['launch_ms_isolated'])"
