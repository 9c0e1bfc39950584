#!/bin/bash
# One gpurun call: GPU tests, smoke, bench line, ncu launch list and ncu --set full captures of the hot kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_profile.sh <tag>
TAG=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || tail -5 gpurun_out/bench_$TAG.err
cut -c1-400 gpurun_out/bench_$TAG.json
python tools/chainbench.py 1800 1 > gpurun_out/chain_$TAG.json 2>&1; tail -1 gpurun_out/chain_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pv_analysis_kernel|pv_synthesis_mirror_kernel|pv_phase_seg" -s 9 -c 3 \
    -f -o gpurun_out/prof_$TAG python tools/kbench.py cfg2 > gpurun_out/ncu_f_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pv_repitch_shared_kernel|pv_stretch_planned_kernel" -s 4 -c 2 \
    -f -o gpurun_out/prof_chain_$TAG python tools/chainbench.py 600 1 > gpurun_out/ncu_c_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
