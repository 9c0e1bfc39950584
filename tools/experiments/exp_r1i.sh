#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
echo "--- default"; python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260
echo "--- mirror analysis onebuf 512"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_ONEBUF=1 FLAN_B200_TPS_ANALYSIS=512 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
echo "--- mirror analysis onebuf 384"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_ONEBUF=1 FLAN_B200_TPS_ANALYSIS=384 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
echo "--- mirror analysis onebuf 640"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_ONEBUF=1 FLAN_B200_TPS_ANALYSIS=640 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
for cfg in cfg1 cfg5 cfg3; do for ob in 0 1; do echo "--- $cfg onebuf=$ob"; FLAN_B200_ONEBUF=$ob python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-130; done; done
for cfg in cfg1 cfg5; do echo "--- $cfg PT16 onebuf=1"; FLAN_B200_PT_ANALYSIS=16 FLAN_B200_ONEBUF=1 python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-130; done
for cfg in cfg1 cfg5; do echo "--- $cfg mirror onebuf=1 tps 512"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_TPS_ANALYSIS=512 FLAN_B200_ONEBUF=1 python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-130; done
