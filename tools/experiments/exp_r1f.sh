#!/bin/bash
mkdir -p gpurun_out
export FLAN_B200_SYNTH_VARIANT=17 FLAN_B200_TPS_SYNTHESIS=384
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for tps in 384 512; do echo "--- mirror synth v2 tps=$tps"; FLAN_B200_TPS_SYNTHESIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-260; done
for cfg in cfg1 cfg5; do
 echo "--- $cfg mirror synth 384"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260
done
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "convert_to_audio_matches_oracle or shards" 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:"pv_synthesis_mirror_kernel" -s 4 -c 1 \
    -f -o gpurun_out/prof_r1f python tools/kbench.py cfg2 > gpurun_out/ncu_f_r1f.log 2>&1
ls -la gpurun_out/prof_r1f*
