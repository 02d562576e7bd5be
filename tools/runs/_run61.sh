mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r2_bench61_n2.json 2> gpurun_out/r2_bench61_n2.err
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench61_n2.json') if l.startswith('{')][-1];print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['gpu_launches'], {k:v.get('value') for k,v in d['roofline_by_mode'].items()})"
wc -l gpurun_out/r2_bench61_n2.json
