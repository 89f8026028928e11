#!/bin/bash
# 2-GPU validation of the multi-rank paths (replicas + C5 shards with the NCCL gather) + C5 breakdown on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_BATCH_TRACE=1 timeout 600 python tools/c5_probe.py > gpurun_out/r2h_c5_probe.log 2>&1; echo "c5 probe rc=$?"; grep "^call\|^\[batch\]" gpurun_out/r2h_c5_probe.log | tail -4 | cut -c1-400
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/r2h_bench_n2.json; tail -3 gpurun_out/r2h_bench_n2.err | cut -c1-300
