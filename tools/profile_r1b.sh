#!/bin/bash
# ncu captures of the 512-channel and WaveFlow kernels (round 1, second batch).  Runs under gpurun.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
ncu --set full --clock-control none --import-source on -k regex:k_gate512_tc -s 10 -c 1 -o gpurun_out/prof_gate512_bf16 -f \
    python bench.py --channels 512 --batch 4 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_g512.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_res512_tc -s 10 -c 1 -o gpurun_out/prof_res512_bf16 -f \
    python bench.py --channels 512 --batch 4 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r512.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_layer_tc -s 40 -c 1 -o gpurun_out/prof_wf_layer_bf16 -f \
    python bench.py --workload waveflow --batch 16 --precision bf16 --steps 1 --warmup 3 > gpurun_out/ncu_wf.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3246 -c 1082 --csv --log-file gpurun_out/launches_waveflow_bf16.csv \
    python bench.py --workload waveflow --batch 16 --precision bf16 --steps 1 --warmup 3 > gpurun_out/bench_under_ncu_wf.log 2>&1
ls -la gpurun_out
