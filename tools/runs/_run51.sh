mkdir -p gpurun_out
for rep in 0 1; do
for v in base q192 r116; do
CWSL_B200_LIB=$PWD/build/libcwsl_$v.so timeout 300 python bench.py --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench51_${v}_$rep.json 2>/dev/null
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench51_${v}_$rep.json') if l.startswith('{')][-1];print('$v $rep bench', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['launch_ms'],4), d['kernel_ms']['isolated_per_receiver'])"
done
done
