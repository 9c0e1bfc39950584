#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for cfg in cfg1 cfg2; do echo "--- $cfg"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
python tools/kbench.py cfg5 1 2>&1 | tail -1 | cut -c1-260
