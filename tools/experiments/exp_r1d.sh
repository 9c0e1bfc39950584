#!/bin/bash
# experiment: mirrored synthesis kernel vs the 8-point one (parity first, then timing)
mkdir -p gpurun_out
FLAN_B200_SYNTH_VARIANT=17 FLAN_B200_TPS_SYNTHESIS=384 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for tps in 384 512; do echo "--- mirror synth tps=$tps"; FLAN_B200_SYNTH_VARIANT=17 FLAN_B200_TPS_SYNTHESIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1; done
for cfg in cfg1 cfg5; do
 echo "--- $cfg mirror synth 384"; FLAN_B200_SYNTH_VARIANT=17 FLAN_B200_TPS_SYNTHESIS=384 python tools/kbench.py $cfg 2>&1 | tail -1
done
