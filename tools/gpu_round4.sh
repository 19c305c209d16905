#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu4.log
timeout 600 python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc $?"; tail -c 400 gpurun_out/bench4.json
timeout 300 python tools/bench_configs.py > gpurun_out/configs4.json 2>gpurun_out/configs4.err; cat gpurun_out/configs4.json
timeout 600 python tools/config4_n2.py > gpurun_out/config4_1gpu_k4d.json 2>gpurun_out/config4.err; cat gpurun_out/config4_1gpu_k4d.json; tail -3 gpurun_out/config4.err
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 200 python tools/prof_n2.py 148 400 1 0.1 > gpurun_out/prof_n2_z400.log 2>&1 && \
  timeout 600 $NCU -k regex:pbr_k4d -o gpurun_out/k4d_z400 -f python tools/prof_n2.py 148 400 1 0.1 > gpurun_out/ncu_k4d.log 2>&1; echo "k4d rc $?"
timeout 200 python tools/prof_n2.py 592 100 1 0.1 > gpurun_out/prof_n2_z100.log 2>&1 && \
  timeout 600 $NCU -k regex:pbr_k4b -o gpurun_out/k4b_v2_z100 -f python tools/prof_n2.py 592 100 1 0.1 > gpurun_out/ncu_k4b.log 2>&1; echo "k4b rc $?"
for f in k4d_z400 k4b_v2_z100; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.txt 2>&1; done
