#!/bin/bash
mkdir -p gpurun_out
for sl in 32 48 64 96 128 192; do echo "--- cfg2 seg_len $sl"; FLAN_B200_SEG_LEN=$sl python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260; done
for sl in 64 128; do echo "--- cfg5 seg_len $sl"; FLAN_B200_SEG_LEN=$sl python tools/kbench.py cfg5 2>&1 | tail -1 | cut -c1-260; done
for sl in 64 128; do echo "--- cfg3 seg_len $sl"; FLAN_B200_SEG_LEN=$sl python tools/kbench.py cfg3 2>&1 | tail -1 | cut -c1-260; done
