#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for cfg in cfg2 cfg1 cfg5 cfg3; do echo "--- $cfg default"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
echo "--- cfg2 onebuf=0"; FLAN_B200_ONEBUF=0 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
echo "--- cfg2 onebuf tps 640"; FLAN_B200_TPS_ANALYSIS=640 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
echo "--- cfg5 analysis pt8"; FLAN_B200_PT_ANALYSIS=8 python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-130
echo "--- cfg5 analysis mirror 384 nobuf"; FLAN_B200_TPS_ANALYSIS=384 FLAN_B200_ONEBUF=0 python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-130
