#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1u.json 2> gpurun_out/bench_r1u.err || tail -5 gpurun_out/bench_r1u.err
cut -c1-330 gpurun_out/bench_r1u.json
for cfg in cfg2 cfg5 cfg3; do echo "--- $cfg"; python tools/kbench.py $cfg 2>&1 | tail -1 | cut -c1-260; done
