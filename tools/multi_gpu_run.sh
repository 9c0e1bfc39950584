#!/bin/bash
# One multi-GPU gpurun call: the bench line and the three sharded BASELINE configs at N GPUs of one node.
# Usage (on the GPU box): bash tools/multi_gpu_run.sh N tag
N=${1:-8}; TAG=${2:-rX}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_g${N}_$TAG.json 2> gpurun_out/bench_g${N}_$TAG.err; tail -c 600 gpurun_out/bench_g${N}_$TAG.json
for cfg in cfg3 cfg4 cfg5; do
  $TR tools/configbench.py --config $cfg > gpurun_out/cfgbench_${cfg}_g${N}_$TAG.json 2> gpurun_out/cfgbench_${cfg}_g${N}_$TAG.err; tail -c 500 gpurun_out/cfgbench_${cfg}_g${N}_$TAG.json
done
