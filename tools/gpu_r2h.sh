#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -x -q > gpurun_out/r2h_stages.log 2>&1; rc=$?
tail -5 gpurun_out/r2h_stages.log
if [ $rc -ne 0 ]; then echo "stage tests failed/hung rc=$rc"; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_gpu_tests.log 2>&1; tail -6 gpurun_out/r2h_gpu_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 1500 gpurun_out/r2h_bench.json
timeout 900 python tools/latency.py --save > gpurun_out/r2h_latency.log 2>&1; grep '"t_mel": 86\|"t_mel": 503' gpurun_out/r2h_latency.log | cut -c1-260
