python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in cfg2 cfg5 cfg3 cfg1; do
 echo "$cfg"; python tools/kbench.py $cfg 2>&1 | tail -1
done
for tps in 512 1024; do echo "tps_s=$tps"; FLAN_B200_TPS_SYNTHESIS=$tps python tools/kbench.py cfg2 2>&1 | tail -1 | cut -c100-260; done
