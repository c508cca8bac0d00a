"""Debug: per-CTA phase timing of the WN layer kernel (clock64 stamps) at BASELINE config 2 size."""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookietts_b200 import WaveGlow, _cabi
from oracle.waveglow_oracle import OracleConfig, synthetic_state_dict
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, TM = 16, 861
sd = synthetic_state_dict(OracleConfig(), 1234)
m = WaveGlow(precision=prec, **bench.MODEL_KW)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval(); m._ensure_packed()
lib = _cabi.load()
tp = TM * 32
if prec == "f16f8":           # fp16 hi [, fp16 lo], e5m2(lo * 2^6), e5m2(hi * 2^-8) planes (include/cwg.h)
    from tests.test_gpu_stages import f8_planes
    x = f8_planes(torch.randn(B, tp, 256, device="cuda"), True)
    h2 = f8_planes(torch.randn(B, tp, 256, device="cuda"), False)
else:
    x = torch.randn(2, B, tp, 256, device="cuda").to(torch.bfloat16)
    h2 = torch.randn(2, B, tp, 256, device="cuda").to(torch.bfloat16)
xo = torch.zeros_like(x); eo = torch.zeros(B, tp, 16, device="cuda")
ntile = (tp + 127) // 128
dbg = torch.zeros(B * ntile * 16, dtype=torch.int64, device="cuda")
mode = _cabi.MODES[prec]
st = torch.cuda.current_stream().cuda_stream
def run():
    _cabi.check(lib.cwg_wn_layer(m._ccfg, m._cw, mode, 5, layer, x.data_ptr(), xo.data_ptr(), h2.data_ptr(), eo.data_ptr(), 0, 0, B, TM, st))
for _ in range(3): run()
torch.cuda.synchronize()
lib.cwg_debug_set_timing(C.c_void_p(dbg.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.cwg_debug_set_timing(None)
d = dbg.cpu().numpy().reshape(-1, 16).astype(np.float64)
names = {"first TMA landed (5-0)": (0, 5), "GEMM1 issue done (6-5)": (5, 6), "GEMM1 complete seen by epi (1-0)": (0, 1),
         "gate (2-1)": (1, 2), "acts_ready->GEMM2 issued (8-7)": (7, 8), "GEMM2 wait in epi (3-2)": (2, 3), "epi2 skip/eo (9-3)": (3, 9), "epi2 res loop (10-9)": (9, 10), "store wait (4-10)": (10, 4), "epi2+store (4-3)": (3, 4), "total (4-0)": (0, 4)}
print(f"{prec} layer {layer}: kernel {e0.elapsed_time(e1):.3f} ms, {d.shape[0]} CTAs")
for k, (a, b) in names.items():
    v = d[:, b] - d[:, a]
    print(f"  {k:38s} median {np.median(v):9.0f}  p10 {np.percentile(v,10):9.0f}  p90 {np.percentile(v,90):9.0f} clk")
