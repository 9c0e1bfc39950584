#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "--- cfg3 default (mirrored, 2 CTAs/SM)"; python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-260
echo "--- cfg3 mirrored without twiddle prefetch"; FLAN_B200_LIB=flan_b200/lib/abl/nopre/libflan_b200.so python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-260
python tools/configbench.py --config cfg3 2>/dev/null | tail -1 | cut -c1-400
