#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for cfg in cfg2 cfg1 cfg5 cfg3; do echo "--- $cfg default"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
echo "--- cfg2 mirror analysis 512 onebuf"; FLAN_B200_PT_ANALYSIS=17 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130
ncu --set full --clock-control none --import-source on -k regex:"pv_analysis_kernel|pv_synthesis_mirror_kernel" -s 8 -c 2 \
    -f -o gpurun_out/prof_r1k python tools/kbench.py cfg2 > gpurun_out/ncu_f_r1k.log 2>&1
ls -la gpurun_out/prof_r1k*
