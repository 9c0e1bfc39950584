python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
for pt in 8 16; do for cfg in cfg2 cfg5 cfg3 cfg1; do
 echo "PT=$pt $cfg"; FLAN_B200_PT_ANALYSIS=$pt python tools/kbench.py $cfg 2>&1 | tail -1
done; done
echo TPS768; FLAN_B200_PT_ANALYSIS=16 FLAN_B200_TPS_ANALYSIS=768 python tools/kbench.py cfg2 2>&1 | tail -1
cat gpurun_out/pytest_gpu.log
