#!/bin/bash
mkdir -p gpurun_out
echo "--- cfg3 default (8-point one-buffer resynthesis)"; FLAN_B200_SYNTH_VARIANT=8 python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-260
export FLAN_B200_SYNTH_VARIANT=17
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg3 or 8192 or shards" 2>&1 | tail -2
echo "--- cfg3 mirrored, 1 CTA/SM (212 regs, two buffers)"; FLAN_B200_SYNTH_ONEBUF=0 python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-260
echo "--- cfg3 mirrored, 2 CTAs/SM (128 regs, one buffer)"; FLAN_B200_SYNTH_ONEBUF=1 python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-260
