#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for cfg in cfg2 cfg1 cfg5 cfg3; do echo "--- $cfg default"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
echo "--- cfg2 analysis tps 384"; FLAN_B200_TPS_ANALYSIS=384 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
echo "--- cfg3 analysis tps 384"; FLAN_B200_TPS_ANALYSIS=384 python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-130
echo "--- cfg3 synthesis tps 1024 onebuf"; FLAN_B200_TPS_SYNTHESIS=1024 python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-260
echo "--- cfg3 synthesis tps 512"; FLAN_B200_TPS_SYNTHESIS=512 python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-260
