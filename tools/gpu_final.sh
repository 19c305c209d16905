#!/bin/bash
# Final evidence of the round: bench line, launch list, ncu --set full captures of the solver kernels (each only after its
# program exited 0 without ncu).  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc $?"; tail -c 600 gpurun_out/bench_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc $?"
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 200 python tools/prof_run.py 65536 0 > gpurun_out/prof_run.log 2>&1 && \
  timeout 600 $NCU -k pbr_k2_solve_n1 -o gpurun_out/k2_final -f python tools/prof_run.py 65536 0 > gpurun_out/ncu_k2.log 2>&1; echo "k2 rc $?"
timeout 200 python tools/prof_n2.py 592 100 1 0.1 > gpurun_out/prof_n2_z100.log 2>&1 && \
  timeout 600 $NCU -k regex:pbr_k4b -o gpurun_out/k4b_z100 -f python tools/prof_n2.py 592 100 1 0.1 > gpurun_out/ncu_k4b.log 2>&1; echo "k4b rc $?"
timeout 200 python tools/prof_n2.py 148 400 1 0.1 > gpurun_out/prof_n2_z400.log 2>&1 && \
  timeout 600 $NCU -k regex:pbr_k4c -o gpurun_out/k4c_z400 -f python tools/prof_n2.py 148 400 1 0.1 > gpurun_out/ncu_k4c.log 2>&1; echo "k4c rc $?"
for f in k2_final k4b_z100 k4c_z400; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.txt 2>&1; done
ls -la gpurun_out
