#!/bin/bash
# Final evidence of round 2: GPU suite, bench line, launch list of the same bench command, ncu --set full capture of the block-per-reactor
# kernel running model M9 (each ncu pass only after the same program exited 0 without ncu).  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r2_gputests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_gputests.log
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc $?"; tail -c 400 gpurun_out/r2_bench_1gpu.json
timeout 300 python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_line_before_launch_list.json 2> /dev/null && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc $?"
NCU="ncu --set full --clock-control none --import-source on -c 1"
PBR_PROF_MODEL=M9 timeout 200 python tools/prof_n2.py 1184 100 1 0.2 > gpurun_out/prof_m9.log 2>&1 && \
  PBR_PROF_MODEL=M9 timeout 900 $NCU -k regex:pbr_k4b_solve -o gpurun_out/k4b_m9 -f python tools/prof_n2.py 1184 100 1 0.2 > gpurun_out/ncu_k4b_m9.log 2>&1; echo "k4b m9 rc $?"
python tools/ncu_summary.py gpurun_out/k4b_m9.ncu-rep > gpurun_out/r2_k4b_m9_N1184_z100.txt 2>&1
timeout 200 python tools/prof_n2.py 4096 100 5 0.5 > gpurun_out/prof_c3.log 2>&1 && \
  timeout 900 $NCU -k regex:pbr_k4bs128_solve -o gpurun_out/k4bs128_c3 -f python tools/prof_n2.py 4096 100 5 0.5 > gpurun_out/ncu_k4bs128.log 2>&1; echo "k4bs128 rc $?"
python tools/ncu_summary.py gpurun_out/k4bs128_c3.ncu-rep > gpurun_out/r2_k4bs128_N4096_z100.txt 2>&1
cat gpurun_out/prof_m9.log; head -34 gpurun_out/r2_k4bs128_N4096_z100.txt
