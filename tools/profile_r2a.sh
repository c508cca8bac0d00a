#!/bin/bash
# Round-2 captures of the persistent layer kernel (k_layer_ps): per-CTA timing, ncu launch list of the bench command,
# one full capture of k_layer_ps per mode.  Runs under gpurun (one GPU).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
python -c "import bench; print(bench.kernel_source_hash())" > gpurun_out/ncu_source_sha256.txt
timeout 600 python tools/ps_timing.py > gpurun_out/r2_ps_timing.log 2>&1; tail -8 gpurun_out/r2_ps_timing.log
for p in f16f8 bf16x3 bf16; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 366 -c 122 --csv \
      --log-file gpurun_out/launches_$p.csv python bench.py --precision $p --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$p.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_layer_ps -s 20 -c 1 \
      -o gpurun_out/prof_layer_ps_$p -f python bench.py --precision $p --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$p.log 2>&1
done
for p in bf16x3 bf16; do
  timeout 600 python bench.py --precision $p --no-cpu-baseline > gpurun_out/r2a_bench_$p.json 2> gpurun_out/r2a_bench_$p.err
  tail -c 1500 gpurun_out/r2a_bench_$p.json
done
ls -la gpurun_out
