set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/numa_probe.py > gpurun_out/r2_numa_n$n.log 2>&1
tail -1 gpurun_out/r2_numa_n$n.log | cut -c1-600
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench12_n8.json 2> gpurun_out/r2_bench12_n8.err
tail -c 600 gpurun_out/r2_bench12_n8.json; tail -3 gpurun_out/r2_bench12_n8.err
