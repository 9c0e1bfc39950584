#!/bin/bash
mkdir -p gpurun_out
echo "--- default"; python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
for v in stna ldel both; do echo "--- $v"; FLAN_B200_LIB=flan_b200/lib/abl/$v/libflan_b200.so python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130; done
for v in stna both; do echo "--- $v cfg3"; FLAN_B200_LIB=flan_b200/lib/abl/$v/libflan_b200.so python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-130; done
echo "--- both chain"; FLAN_B200_LIB=flan_b200/lib/abl/both/libflan_b200.so python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-160
echo "--- both cfg5"; FLAN_B200_LIB=flan_b200/lib/abl/both/libflan_b200.so python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-130
