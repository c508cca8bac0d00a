#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_pack_c.py -x -q > gpurun_out/r2c_parity.log 2>&1
tail -15 gpurun_out/r2c_parity.log
timeout 600 python tools/ps_timing.py > gpurun_out/r2c_ps_timing.log 2>&1; tail -8 gpurun_out/r2c_ps_timing.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 2500 gpurun_out/r2c_bench.json
