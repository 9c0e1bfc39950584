#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "--- chain default"; python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-330
echo "--- chain seg 128"; FLAN_B200_SEG_LEN=128 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-330
echo "--- chain seg 256"; FLAN_B200_SEG_LEN=256 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-330
echo "--- cfg1"; python tools/kbench.py cfg1 2>&1 | tail -1 | cut -c1-260
