#!/bin/bash
# experiment: mirrored analysis kernel vs the PT16 one (parity first, then timing)
mkdir -p gpurun_out
FLAN_B200_PT_ANALYSIS=17 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "convert_to_pv or odd_shapes or ragged or golden or full_size_cfg2 or shards" 2>&1 | tail -3
echo "--- default"; python tools/kbench.py cfg2 2>&1 | tail -1
for tps in 384 512 640; do echo "--- mirror tps=$tps"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1; done
for cfg in cfg1 cfg5; do
 echo "--- $cfg default"; python tools/kbench.py $cfg 2>&1 | tail -1
 echo "--- $cfg mirror 512"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_TPS_ANALYSIS=512 python tools/kbench.py $cfg 2>&1 | tail -1
 echo "--- $cfg mirror 384"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_TPS_ANALYSIS=384 python tools/kbench.py $cfg 2>&1 | tail -1
done
