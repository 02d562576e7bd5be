mkdir -p gpurun_out
for c in 32 48 64 96 128 256 512 1024 2048 4096; do
PROBE_KINDS=bench PROBE_CHANNELS=$c timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe49_$c.log 2>&1
python - $c <<'PY'
import json,sys,re
c=sys.argv[1]
out={}
for l in open(f'gpurun_out/r2_probe49_{c}.log'):
    m=re.match(r'bench (\w+) (\{.*\})',l)
    if m:
        d=json.loads(m.group(2)); out[m.group(1)]=(round(d['kernel_ms']['demod_ms'],4), d['resid_db']['worst'], d['max_lsb'])
print(c, out)
PY
done
