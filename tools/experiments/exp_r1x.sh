#!/bin/bash
mkdir -p gpurun_out
echo "--- cfg5 default"; python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-260
echo "--- cfg5 synth tps 512"; FLAN_B200_TPS_SYNTHESIS=512 python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-260
echo "--- chain default"; python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-330
echo "--- chain synth tps 512"; FLAN_B200_TPS_SYNTHESIS=512 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-330
