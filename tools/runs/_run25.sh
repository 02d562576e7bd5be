mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/r2_bench25_n$N.json 2> gpurun_out/r2_bench25_n$N.err
tail -3 gpurun_out/r2_bench25_n$N.err
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench25_n$N.json') if l.startswith('{')][-1];print('bench N=$N', round(d['value']), round(d['ms_per_step'],2), d['e2e']['value'], d['e2e']['d2h_gbs_per_gpu'], {k:round(v['value']) for k,v in d['other_modes'].items()}, d.get('gathered_checksums'))"
