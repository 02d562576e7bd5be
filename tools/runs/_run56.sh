mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest56.log; cat gpurun_out/r2_pytest56.log
timeout 900 python bench.py > gpurun_out/r2_bench56.json 2> gpurun_out/r2_bench56.err; tail -c 300 gpurun_out/r2_bench56.err
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench56.json') if l.startswith('{')][-1];print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'], {k:v.get('value') for k,v in d['roofline_by_mode'].items()}, d['roofline']['frac'])"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench56_ref.json 2>/dev/null; tail -c 400 gpurun_out/r2_bench56_ref.json
