mkdir -p gpurun_out
for v in v4 pf148 pf296 pf592 pf1184 v4 pf148 pf296 pf592 pf1184; do
CWSL_B200_LIB=$PWD/build/libcwsl_$v.so timeout 300 python bench.py --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench57_${v}.json 2>/dev/null
python -c "
import json;d=[json.loads(l) for l in open('gpurun_out/r2_bench57_${v}.json') if l.startswith('{')][-1];k=d['kernel_ms']['isolated_per_receiver'];print('$v bench', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['launch_ms'],4), round(k['main_kernel_ms'],4), round(k['quantise_and_clear_ms'],4))" || true
done
