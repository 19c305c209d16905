#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu5.log
timeout 300 python tools/ab_variants.py 4096 100 > gpurun_out/ab5_z100.jsonl 2>&1; cat gpurun_out/ab5_z100.jsonl
timeout 300 python tools/ab_variants.py 1184 400 > gpurun_out/ab5_z400.jsonl 2>&1; cat gpurun_out/ab5_z400.jsonl
timeout 300 python tools/config4_n2.py > gpurun_out/config4_1gpu_cont.json 2>gpurun_out/config4.err; cat gpurun_out/config4_1gpu_cont.json; tail -3 gpurun_out/config4.err
timeout 200 python tools/bench_configs.py > gpurun_out/configs5.json 2>gpurun_out/configs5.err; cat gpurun_out/configs5.json
