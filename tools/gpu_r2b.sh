#!/bin/bash
# C-side packing + torch.ops boundary: pack parity, stage tests, full GPU suite; then the round-2 profile pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pack_c.py tests/test_gpu_stages.py -x -q > gpurun_out/r2b_pack.log 2>&1; rc=$?
tail -15 gpurun_out/r2b_pack.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_gpu_tests.log 2>&1
tail -12 gpurun_out/r2b_gpu_tests.log
bash tools/profile_r2a.sh
