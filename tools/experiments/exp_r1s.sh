#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for cfg in cfg2 cfg3 cfg1; do echo "--- $cfg default (newest pair staged by cp.async)"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
echo "--- chain default"; python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-200
echo "--- cfg5 pt8"; FLAN_B200_PT_ANALYSIS=8 python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-130
echo "--- cfg2 onebuf=0"; FLAN_B200_ONEBUF=0 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
