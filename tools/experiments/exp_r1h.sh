#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
echo "--- default (PT16 analysis, mirror synthesis v3)"; python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260
for tps in 512 640 768; do echo "--- onebuf PT16 tps=$tps"; FLAN_B200_LIB=flan_b200/lib/abl/onebuf/libflan_b200.so FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130; done
for co in 30 44 58; do echo "--- onebuf PT16 tps=512 carveout $co"; FLAN_B200_CARVEOUT=$co FLAN_B200_LIB=flan_b200/lib/abl/onebuf/libflan_b200.so FLAN_B200_TPS_ANALYSIS=512 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130; done
echo "--- onebuf PT8 tps=768"; FLAN_B200_PT_ANALYSIS=8 FLAN_B200_LIB=flan_b200/lib/abl/onebuf/libflan_b200.so FLAN_B200_TPS_ANALYSIS=768 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
echo "--- onebuf PT8 tps=1024"; FLAN_B200_PT_ANALYSIS=8 FLAN_B200_LIB=flan_b200/lib/abl/onebuf/libflan_b200.so FLAN_B200_TPS_ANALYSIS=1024 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
for cfg in cfg1 cfg5 cfg3; do echo "--- $cfg default"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
