#!/bin/bash
mkdir -p gpurun_out
for cfg in cfg3 cfg4 cfg5; do python tools/configbench.py --config $cfg > gpurun_out/cfgbench_${cfg}_g1_v4.json 2> gpurun_out/cfgbench_${cfg}_g1_v4.err; tail -c 400 gpurun_out/cfgbench_${cfg}_g1_v4.json; done
echo "--- chain default"; python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-330
echo "--- chain analysis pt16 onebuf"; FLAN_B200_PT_ANALYSIS=16 FLAN_B200_ONEBUF=1 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-200
echo "--- chain analysis pt16"; FLAN_B200_PT_ANALYSIS=16 FLAN_B200_ONEBUF=0 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-200
echo "--- chain analysis mirror 512 onebuf"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_ONEBUF=1 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-200
echo "--- chain analysis mirror 384"; FLAN_B200_PT_ANALYSIS=17 FLAN_B200_ONEBUF=0 FLAN_B200_TPS_ANALYSIS=384 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-200
echo "--- chain analysis pt8 tps 1024"; FLAN_B200_PT_ANALYSIS=8 FLAN_B200_TPS_ANALYSIS=1024 python tools/chainbench.py 1800 1 2>&1 | tail -1 | cut -c1-200
