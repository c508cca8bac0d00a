"""Latency of one infer call for small inputs (serving shapes): device time by CUDA events and host wall time, with the
module's CUDA-graph replay (graphs="auto", the default) and without it.  Writes profiles/r2_latency.json when run with --save.
Shapes: BASELINE config 1 (1 x 1 s), the notebook's 5.84-s clip (scripts/WaveGlowFlow Inference Speed Testing.ipynb), 1 x 10 s, 4 x 10 s."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookietts_b200 import WaveGlow
from cookietts_b200.synthetic import ModelConfig, synthetic_state_dict
import bench

sd = synthetic_state_dict(ModelConfig(), 1234)
recs = []
for prec in ("f16f8", "bf16x3", "bf16"):
    for graphs in ("auto", False):
        m = WaveGlow(precision=prec, graphs=graphs, **bench.MODEL_KW)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
        for B, Tm in ((1, 86), (1, 503), (1, 861), (4, 861)):
            mel = torch.randn(B, 80, Tm, device="cuda"); z = torch.randn(B, Tm * 256, device="cuda")
            for _ in range(3): m.infer(mel, sigma=0.666, z=z)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            t = time.perf_counter(); e0.record()
            for _ in range(n): m.infer(mel, sigma=0.666, z=z)
            e1.record()
            torch.cuda.synchronize()
            t_wall = (time.perf_counter() - t) / n
            dev = e0.elapsed_time(e1) / n
            used = bool(m._graphs)
            rec = dict(precision=prec, graphs=str(graphs), graph_used=used, batch=B, t_mel=Tm, audio_s=round(Tm * 256 / 22050, 2),
                       device_ms=round(dev, 3), wall_ms=round(t_wall * 1e3, 3), xrt=round(B * Tm * 256 / 22050 / t_wall, 1))
            print(json.dumps(rec), flush=True)
            recs.append(rec)
        del m
if "--save" in sys.argv:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(recs, open("gpurun_out/r2_latency.json", "w"), indent=1)
