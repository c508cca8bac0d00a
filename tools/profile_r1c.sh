#!/bin/bash
# Round-1 final captures (f16f8 = the bench default) (after the TMEM-A GEMM2, vectorised boundary kernel and two-CTA cond GEMM).  Runs under gpurun:
# ncu launch list of the bench command + one full capture of each kernel of the 256-channel path.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
for p in f16f8 bf16x3 bf16; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 366 -c 122 --csv \
      --log-file gpurun_out/launches_$p.csv python bench.py --precision $p --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$p.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_layer_tc -s 20 -c 1 \
      -o gpurun_out/prof_layer_$p -f python bench.py --precision $p --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$p.log 2>&1
done
ncu --set full --clock-control none -k regex:k_cond_tc -s 2 -c 1 -o gpurun_out/prof_cond_bf16x3 -f \
    python bench.py --precision bf16x3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cond.log 2>&1
ncu --set full --clock-control none -k regex:k_flow_boundary -s 2 -c 1 -o gpurun_out/prof_boundary -f \
    python bench.py --precision bf16x3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_boundary.log 2>&1
ls -la gpurun_out
