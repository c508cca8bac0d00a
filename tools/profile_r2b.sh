#!/bin/bash
# Round-2 final captures: ncu launch list of the bench command per mode, one full capture of k_layer_ps per mode, of the
# cond GEMM, the boundary kernel and the WaveFlow layer kernel.  Records the sha256 of the kernel sources it ran from.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
python -c "import bench; print(bench.kernel_source_hash())" > gpurun_out/ncu_source_sha256.txt
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
for p in f16f8 bf16x3 bf16; do
  # the kernels of a bench step only (the one-time cwg_pack_weights kernels k_effective / k_w1 / ... are filtered out)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k regex:"k_layer_ps|k_cond_tc|k_flow_boundary|k_im2col_mel|k_nonfinite|k_cond_bias" -s 372 -c 124 --csv \
      --log-file gpurun_out/launches_$p.csv $B --precision $p > gpurun_out/bench_under_ncu_$p.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_layer_ps -s 20 -c 1 \
      -o gpurun_out/prof_layer_ps_$p -f $B --precision $p > gpurun_out/ncu_$p.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_layer_ps -s 24 -c 1 \
    -o gpurun_out/prof_layer0_fold_f16f8 -f $B --precision f16f8 > gpurun_out/ncu_l0.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_cond_tc -s 2 -c 1 -o gpurun_out/prof_cond_f16f8 -f $B --precision f16f8 > gpurun_out/ncu_cond.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_flow_boundary -s 2 -c 1 -o gpurun_out/prof_boundary_f16f8 -f $B --precision f16f8 > gpurun_out/ncu_boundary.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf_layer_tc -s 300 -c 1 -o gpurun_out/prof_wf_layer_bf16 -f \
    python bench.py --config 5 --precision bf16 --steps 1 --warmup 3 > gpurun_out/ncu_wf.log 2>&1
ls -la gpurun_out | tail -20
