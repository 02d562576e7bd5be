set -x
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest37.log
tail -4 gpurun_out/r2_pytest37.log
timeout 1200 python bench.py > gpurun_out/r2_bench37.json 2> gpurun_out/r2_bench37.err
tail -2 gpurun_out/r2_bench37.err
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench37.json') if l.startswith('{')][-1];print('bench', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['launch_ms'],4), d['roofline']['frac'], d['roofline']['step_hbm']['frac'], d['e2e']['value'], {k:round(v['value']) for k,v in d['other_modes'].items()}, d['device_footprint']['phase_state_bytes'])"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench37_ref.json 2>/dev/null
cut -c1-300 gpurun_out/r2_bench37_ref.json
