#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flow_decoder.py -x -q > gpurun_out/r2e_fd.log 2>&1; tail -12 gpurun_out/r2e_fd.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_gpu_tests.log 2>&1; tail -6 gpurun_out/r2e_gpu_tests.log
