#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_gpu_tests.log 2>&1; tail -4 gpurun_out/r2j_gpu_tests.log
timeout 600 python bench.py --config 5 --steps 3 > gpurun_out/r2j_bench_c5.json 2> gpurun_out/r2j_bench_c5.err; tail -c 1800 gpurun_out/r2j_bench_c5.json; tail -3 gpurun_out/r2j_bench_c5.err
timeout 600 python bench.py --config 5 --precision bf16 --steps 3 > gpurun_out/r2j_bench_c5_bf16.json 2>/dev/null; tail -c 900 gpurun_out/r2j_bench_c5_bf16.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
