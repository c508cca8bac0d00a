#!/bin/bash
# N=8 through the driver's launch line: headline (config 2, weak) + extra_configs 3/4/5 at 8 GPUs
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err
tail -c 3000 gpurun_out/r2k_bench_n8.json; tail -4 gpurun_out/r2k_bench_n8.err
