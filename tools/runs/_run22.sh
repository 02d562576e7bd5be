set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest22.log
tail -5 gpurun_out/r2_pytest22.log
timeout 1200 python bench.py > gpurun_out/r2_bench22.json 2> gpurun_out/r2_bench22.err
tail -3 gpurun_out/r2_bench22.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench22.json'));print('bench', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['launch_ms'],4), d['e2e']['value'], d['kernel_ms']['isolated_per_receiver'], {k:round(v['value']) for k,v in d['other_modes'].items()})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"demod|quantise|guard|clear_u32|phase" -c 400 --csv --log-file gpurun_out/r2_launches22.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-configs --no-station --no-cpu-baseline --no-other-modes > gpurun_out/r2_bench22_ncu.log 2>&1
tail -3 gpurun_out/r2_launches22.csv | cut -c1-250
