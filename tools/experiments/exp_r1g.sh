#!/bin/bash
mkdir -p gpurun_out
export FLAN_B200_PT_ANALYSIS=17
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for tps in 384 512; do echo "--- mirror analysis v2 tps=$tps"; FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260; done
for cfg in cfg1 cfg5; do
 for tps in 384 512; do echo "--- $cfg mirror analysis $tps"; FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
done
FLAN_B200_TPS_ANALYSIS=512 ncu --set full --clock-control none --import-source on -k regex:"pv_analysis_mirror_kernel" -s 4 -c 1 \
    -f -o gpurun_out/prof_r1g python tools/kbench.py cfg2 > gpurun_out/ncu_f_r1g.log 2>&1
ls -la gpurun_out/prof_r1g*
