#!/bin/bash
# ncu captures of the "next"-row kernels: Denoiser GEMMs, int16 PCM conversion, front-end conv.  Under gpurun.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
python tools/bench_post.py > gpurun_out/bench_post.log 2>&1
ncu --set full --clock-control none -k regex:k_dn_gemm -s 6 -c 2 -o gpurun_out/prof_denoise_gemm -f python tools/bench_post.py > gpurun_out/ncu_dn.log 2>&1
ncu --set full --clock-control none -k regex:k_pcm16 -s 16 -c 1 -o gpurun_out/prof_pcm16 -f python tools/bench_post.py > gpurun_out/ncu_pcm.log 2>&1
ncu --set full --clock-control none -k regex:k_cn_conv -s 14 -c 1 -o gpurun_out/prof_cn_conv -f python tools/bench_post.py > gpurun_out/ncu_cn.log 2>&1
cat gpurun_out/bench_post.log
