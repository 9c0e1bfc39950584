#!/bin/bash
mkdir -p gpurun_out
echo "--- cfg2 default"; python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260
export FLAN_B200_SYNTH_ONEBUF=1
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for cfg in cfg2 cfg5 cfg1; do echo "--- $cfg synth onebuf"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
for co in 44 58 72 86; do echo "--- cfg2 synth onebuf carveout $co"; FLAN_B200_CARVEOUT=$co python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260; done
echo "--- chain synth onebuf"; python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-330
