mkdir -p gpurun_out
for v in v0 v1 v3 v4 v0 v1; do
export CWSL_B200_LIB=$PWD/build/libcwsl_$v.so
PROBE_KINDS=bench timeout 300 python tools/r2_probe.py > gpurun_out/r2_probe18_$v.log 2>&1
echo "$v probe: $(grep -E 'stft_raw' gpurun_out/r2_probe18_$v.log | cut -c1-150)"
done
export CWSL_B200_LIB=$PWD/build/libcwsl_v1.so
PROF_MODE=stft PROF_SLOTS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:demod_chan -c 1 -s 1 -o gpurun_out/r2_chan_18_v1 --force-overwrite python tools/profile_target.py > gpurun_out/ncu_chan18.log 2>&1
tail -2 gpurun_out/ncu_chan18.log
