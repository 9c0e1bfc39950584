for tps in 512 640; do echo "PT16 tps=$tps"; FLAN_B200_TPS_ANALYSIS=$tps FLAN_B200_PT_ANALYSIS=16 python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c1-130; done
