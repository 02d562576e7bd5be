mkdir -p gpurun_out
for v in $VARIANTS; do
export CWSL_B200_LIB=$PWD/build/libcwsl_$v.so
PROBE_KINDS=bench timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe27_$v.log 2>&1
echo "$v probe: $(grep -E 'stft_raw' gpurun_out/r2_probe27_$v.log | cut -c1-150)"
done
