#!/bin/bash
# One GPU-box session: parity suite, bench, launch list, variant A/B.  Outputs under gpurun_out/.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" 
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"; tail -c 1500 gpurun_out/bench.json
timeout 600 python tools/ab_variants.py 4096 100 > gpurun_out/ab_z100.jsonl 2>&1; cat gpurun_out/ab_z100.jsonl
timeout 600 python tools/ab_variants.py 1184 400 > gpurun_out/ab_z400.jsonl 2>&1; cat gpurun_out/ab_z400.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc $?"
tail -3 gpurun_out/pytest_gpu.log
