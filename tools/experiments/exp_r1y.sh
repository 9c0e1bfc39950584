#!/bin/bash
mkdir -p gpurun_out
echo "--- cfg2 default"; python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
export FLAN_B200_PT_ANALYSIS=17 FLAN_B200_RING=0
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "convert_to_pv or golden or full_size_cfg2 or shards" 2>&1 | tail -2
for ob in 1 0; do for tps in 512 384; do echo "--- cfg2 mirror no ring onebuf=$ob tps=$tps"; FLAN_B200_ONEBUF=$ob FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130; done; done
echo "--- cfg5 mirror no ring onebuf 512"; FLAN_B200_ONEBUF=1 FLAN_B200_TPS_ANALYSIS=512 python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-130
echo "--- chain mirror no ring onebuf 512"; FLAN_B200_ONEBUF=1 FLAN_B200_TPS_ANALYSIS=512 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-200
