"""Per-call time of the fp32 CUDA-core WaveFlow mode on the layout of the reference author's trained checkpoints (squeeze
height 20, 8 flows, 8 x 128, depthwise-separable 7x7 in_layers; SURVEY 8d config-5 note), one utterance.
Appends to gpurun_out/waveflow_sep_timing.jsonl."""
import json, os, sys, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cookietts_b200 import WaveFlow
from cookietts_b200.synthetic import WaveFlowConfig, waveflow_state_dict, waveflow_reference_kwargs

warnings.simplefilter("ignore")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 160
cfg = WaveFlowConfig(n_group=20, kernel_size_h=7, kernel_size_w=7, seperable_conv=True, win_length=1200, hop_length=300)
graphs = len(sys.argv) > 2 and sys.argv[2] == "graph"          # CUDA-graph replay of the row-by-row launch sequence
m = WaveFlow(precision="ffma", graphs=True if graphs else False, **waveflow_reference_kwargs(cfg))
m.load_state_dict({k: torch.from_numpy(v) for k, v in waveflow_state_dict(cfg, 3).items()})
m = m.cuda().eval()
rs = np.random.RandomState(0)
mel = torch.from_numpy(np.clip(rs.standard_normal((1, 80, frames)) * 2 - 5, -11.5, 2).astype(np.float32)).cuda()
z = torch.randn(1, frames * cfg.hop_length, device="cuda")
for _ in range(2):
    a = m.infer(mel, sigma=0.666, z=z, return_CPU=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    a = m.infer(mel, sigma=0.666, z=z, return_CPU=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
out = dict(frames=frames, cuda_graph=graphs, samples=int(a.shape[1]), ms_per_call=ms, samples_per_s=a.shape[1] / ms * 1e3, finite=bool(torch.isfinite(a).all()),
           model="WaveFlow h=20, 8 flows, 8 x 128, separable 7x7, precision ffma")
print(out)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "waveflow_sep_timing.jsonl"), "a") as f:
    f.write(json.dumps(out) + "\n")
