#!/bin/bash
mkdir -p gpurun_out
echo "--- cfg2 default"; python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260
echo "--- cfg2 synth 4 CTAs/SM: 128 regs + one buffer"; FLAN_B200_SYNTH_ONEBUF=1 FLAN_B200_TPS_SYNTHESIS=512 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260
echo "--- cfg5 synth 128 regs + one buffer"; FLAN_B200_SYNTH_ONEBUF=1 FLAN_B200_TPS_SYNTHESIS=512 python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-260
echo "--- chain synth 128 regs + one buffer"; FLAN_B200_SYNTH_ONEBUF=1 FLAN_B200_TPS_SYNTHESIS=512 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-330
