"""Latency of one infer call for small inputs (serving shapes): device time by CUDA events and host wall time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookietts_b200 import WaveGlow
from oracle.waveglow_oracle import OracleConfig, synthetic_state_dict
import bench

sd = synthetic_state_dict(OracleConfig(), 1234)
for prec in ("bf16x3", "bf16"):
    m = WaveGlow(precision=prec, **bench.MODEL_KW)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
    for B, Tm in ((1, 86), (1, 503), (1, 861), (4, 861)):
        mel = torch.randn(B, 80, Tm, device="cuda"); z = torch.randn(B, Tm * 256, device="cuda")
        for _ in range(3): m.infer(mel, sigma=0.666, z=z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        t = time.perf_counter(); e0.record()
        for _ in range(n): m.infer(mel, sigma=0.666, z=z)
        e1.record(); t_host = (time.perf_counter() - t) / n
        torch.cuda.synchronize()
        t_wall = (time.perf_counter() - t) / n
        dev = e0.elapsed_time(e1) / n
        print(f"{prec} B={B} T_mel={Tm} ({Tm*256/22050:.2f} s): device {dev:.3f} ms, host enqueue {t_host*1e3:.3f} ms, wall {t_wall*1e3:.3f} ms -> {B*Tm*256/22050/(t_wall):.0f}x RT")
