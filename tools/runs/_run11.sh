set -x
mkdir -p gpurun_out
for v in 4 2; do
CWSL_QUANT_VEC=$v timeout 600 python bench.py --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench11_v$v.json 2> gpurun_out/r2_bench11_v$v.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench11_v$v.json'));print('vec$v', d['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['kernel_ms']['isolated_per_receiver'])"
done
