#!/bin/bash
# Round-2 addendum: one full ncu capture of the general fp32 mode's gated conv kernel (k_fd_conv<2>, cwg_axg_flow) on the
# 12-flow 8 x 256 ax model, 1 x 861 frames (tools/general_mode_timing.py), and of the WaveFlow fp32 GEMM (k_wff_gemm<0>).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fd_conv -s 40 -c 1 -o gpurun_out/prof_fd_conv_gate -f \
    python tools/general_mode_timing.py 861 > gpurun_out/ncu_fd_conv.log 2>&1
ncu -i gpurun_out/prof_fd_conv_gate.ncu-rep --page raw --csv > gpurun_out/prof_fd_conv_gate_raw.csv 2>/dev/null
ls -la gpurun_out | grep prof_fd
