#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/ab_variants.py 4096 100 > gpurun_out/ab2_z100.jsonl 2>&1; cat gpurun_out/ab2_z100.jsonl
timeout 600 python tools/ab_variants.py 1184 400 > gpurun_out/ab2_z400.jsonl 2>&1; cat gpurun_out/ab2_z400.jsonl
for z in 160 224; do for from in 128 512; do
  echo "zNo=$z scan_from=$from"; PBR_K4B_SCAN_FROM=$from timeout 300 python tools/bench_n2.py 1184 $z 2>&1 | tail -1 | cut -c1-260
done; done | tee gpurun_out/ab_scan_crossover.txt
timeout 600 python tools/bench_configs.py > gpurun_out/configs.json 2>gpurun_out/configs.err; cat gpurun_out/configs.json
timeout 600 python tools/config4_n2.py > gpurun_out/config4_1gpu.json 2>gpurun_out/config4.err; cat gpurun_out/config4_1gpu.json; tail -3 gpurun_out/config4.err
bash tools/gpu_final.sh
