mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench64.json 2>/dev/null
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench64.json') if l.startswith('{')][-1];print('bench', round(d['value']), round(d['ms_per_step'],2), d['parity_in_run']['stft_vs_exact'] if 'stft_vs_exact' in d.get('parity_in_run',{}) else list(d.get('parity_in_run',{}).keys())[:4])"
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or packed or length" 2>&1 | tail -1
