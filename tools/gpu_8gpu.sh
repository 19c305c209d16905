#!/bin/bash
# 8-GPU records of the round: BASELINE config 4 at full size, the N1 north-star ensemble, bench.py --gpus 8.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
ORACLE=0 timeout 130 $TR tools/config4_n2.py > gpurun_out/config4_8gpu.json 2> gpurun_out/config4_8gpu.err; echo "config4 rc $?"; tail -1 gpurun_out/config4_8gpu.json
timeout 70 $TR tools/northstar_1m.py > gpurun_out/northstar_8gpu.json 2> gpurun_out/northstar_8gpu.err; echo "northstar rc $?"; tail -1 gpurun_out/northstar_8gpu.json
timeout 60 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc $?"; tail -c 300 gpurun_out/bench_8gpu.json
