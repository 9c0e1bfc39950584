#!/bin/bash
# bench.py at N GPUs of one node (the driver's own launch line). Usage on the GPU box: bash tools/bench_n.sh N tag
N=${1:-2}; TAG=${2:-rX}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_g${N}_$TAG.json 2> gpurun_out/bench_g${N}_$TAG.err
tail -c 700 gpurun_out/bench_g${N}_$TAG.json
