#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for cfg in cfg2 cfg1 cfg5 cfg3; do echo "--- $cfg default (newest-pair register prefetch, seg 128 for N>=4096)"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
echo "--- cfg5 pt8"; FLAN_B200_PT_ANALYSIS=8 python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-130
echo "--- cfg2 pt8"; FLAN_B200_PT_ANALYSIS=8 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
