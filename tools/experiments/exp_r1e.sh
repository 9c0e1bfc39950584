#!/bin/bash
mkdir -p gpurun_out
./tools/micro/pipes.bin; ./tools/micro/lat.bin
FLAN_B200_PT_ANALYSIS=17 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "convert_to_pv or odd_shapes or ragged or golden or full_size_cfg2 or shards" 2>&1 | tail -2
export FLAN_B200_SYNTH_VARIANT=17 FLAN_B200_TPS_SYNTHESIS=384 FLAN_B200_PT_ANALYSIS=17
for tps in 384 512; do echo "--- mirror(conj tw) tps=$tps"; FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-230; done
for tps in 384 512; do echo "--- mirror + prefetch L1 tps=$tps"; FLAN_B200_LIB=flan_b200/lib/abl/pfl1/libflan_b200.so FLAN_B200_TPS_ANALYSIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1| cut -c1-230; done
for co in 50 60 72 86 100; do echo "--- carveout $co tps 512"; FLAN_B200_CARVEOUT=$co FLAN_B200_TPS_ANALYSIS=512 python tools/kbench.py cfg2 2>&1 | tail -1| cut -c1-230; done
for co in 50 72 100; do echo "--- carveout $co tps 384"; FLAN_B200_CARVEOUT=$co FLAN_B200_TPS_ANALYSIS=384 python tools/kbench.py cfg2 2>&1 | tail -1| cut -c1-230; done
FLAN_B200_TPS_ANALYSIS=512 ncu --set full --clock-control none --import-source on -k regex:"pv_analysis_mirror_kernel|pv_synthesis_mirror_kernel" -s 8 -c 2 \
    -f -o gpurun_out/prof_r1e python tools/kbench.py cfg2 > gpurun_out/ncu_f_r1e.log 2>&1
ls -la gpurun_out/prof_r1e*
