#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/parity_report.py > gpurun_out/parity_final.txt 2>gpurun_out/parity_final.err; echo "parity rc $?"; cat gpurun_out/parity_final.txt
timeout 400 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_last.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_last.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_last.log 2>&1; echo "ncu list rc $?"
