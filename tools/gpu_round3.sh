#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu3.log
timeout 600 python tools/ab_variants.py 4096 100 > gpurun_out/ab3_z100.jsonl 2>&1; cat gpurun_out/ab3_z100.jsonl
timeout 600 python tools/ab_variants.py 1184 400 > gpurun_out/ab3_z400.jsonl 2>&1; cat gpurun_out/ab3_z400.jsonl
for oc in 0 1; do echo "zNo=400 PBR_K4C_ONCHIP=$oc"; PBR_K4C_ONCHIP=$oc timeout 300 python tools/bench_n2.py 1184 400 2>&1 | tail -1 | cut -c1-230; done | tee gpurun_out/ab3_onchip.txt
for oc in 0 1; do echo "zNo=300 PBR_K4C_ONCHIP=$oc"; PBR_K4C_ONCHIP=$oc timeout 300 python tools/bench_n2.py 1184 300 2>&1 | tail -1 | cut -c1-230; done | tee -a gpurun_out/ab3_onchip.txt
timeout 300 python tools/bench_configs.py > gpurun_out/configs3.json 2>gpurun_out/configs3.err; cat gpurun_out/configs3.json
