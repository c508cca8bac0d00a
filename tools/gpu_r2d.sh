#!/bin/bash
# N=2: async gather + extra configs through the driver's launch line; N=1 default line with extras
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
tail -c 6000 gpurun_out/r2d_bench_n2.json; tail -5 gpurun_out/r2d_bench_n2.err
timeout 900 python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
tail -c 3000 gpurun_out/r2d_bench_n1.json; tail -3 gpurun_out/r2d_bench_n1.err
timeout 600 python -m pytest tests/test_pack_c.py -x -q 2>&1 | tail -3
