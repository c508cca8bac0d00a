#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python tools/memcheck_small.py > gpurun_out/r2l_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|Out-of-range|mel20|speaker256|fd_untts|Error" gpurun_out/r2l_memcheck.log | head -40
