#!/bin/bash
mkdir -p gpurun_out
echo "--- default"; python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
export FLAN_B200_LIB=flan_b200/lib/abl/winld/libflan_b200.so
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "convert_to_pv or odd_shapes or ragged or golden" 2>&1 | tail -2
for tps in 512 640 768; do echo "--- winld tps=$tps"; FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130; done
for tps in 640 768; do echo "--- winld tps=$tps onebuf=0"; FLAN_B200_ONEBUF=0 FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130; done
for tps in 512 640 768; do echo "--- winld cfg3 tps=$tps"; FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-130; done
for tps in 768 1024; do echo "--- winld cfg5 pt8 tps=$tps"; FLAN_B200_PT_ANALYSIS=8 FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-130; done
