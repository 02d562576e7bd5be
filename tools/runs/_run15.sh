mkdir -p gpurun_out
for v in default 120_104_136 120_88_152 128_96_160 128_104_152; do
if [ $v = default ]; then unset CWSL_B200_LIB; else export CWSL_B200_LIB=$PWD/build/libcwsl_regs_$v.so; fi
PROBE_KINDS=bench timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe15_$v.log 2>&1
echo "$v probe: $(grep -E 'stft_raw' gpurun_out/r2_probe15_$v.log | cut -c1-140)"
timeout 300 python bench.py --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench15_$v.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r2_bench15_$v.json'));print('$v bench', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['launch_ms'],4))"
done
