#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NGPU:-2} --steps 5 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"
tail -c 1200 gpurun_out/n2_bench.json; echo; tail -3 gpurun_out/n2_bench.err | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus ${NGPU:-2} --steps 1 --warmup 1 > gpurun_out/n2_reference.json 2> gpurun_out/n2_reference.err; echo "reference n2 rc=$?"
cut -c1-400 gpurun_out/n2_reference.json
