"""Per-call time of the general fp32 mode (cwg_axg_flow) next to the packed kernels on the same shape: the 12-flow 8 x 256 ax
WaveGlow on one 10-s utterance (861 frames).  Writes gpurun_out/general_mode_timing.json."""
import json, os, sys, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cookietts_b200 import WaveGlowAx
from oracle.make_golden_waveflow import reference_kwargs_ax1d          # constructor kwargs only (no oracle arithmetic)
from oracle.waveglow_ax_oracle import AxConfig, synthetic_state_dict

warnings.simplefilter("ignore")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 861
out = {}
rs = np.random.RandomState(0)
for tag, kw, prec in (("packed_f16f8", dict(), "f16f8"), ("packed_ffma", dict(), "ffma"),
                      ("general_gtu_listed_dilations", dict(dilations_w=[1, 2, 4, 8, 16, 32, 64, 1]), "ffma"),
                      ("general_gsiu_merged", dict(gated_unit="GSIU", merge_res_skip=True), "ffma")):
    cfg = AxConfig(**kw)
    m = WaveGlowAx(precision=prec, graphs=False, **reference_kwargs_ax1d(cfg))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synthetic_state_dict(cfg, 7).items()})
    m = m.cuda().eval()
    mel = torch.from_numpy(np.clip(rs.standard_normal((1, 80, frames)) * 2 - 5, -11.5, 2).astype(np.float32)).cuda()
    z = torch.randn(1, frames * cfg.hop_length, device="cuda")
    for _ in range(2):
        a = m.infer(mel, sigma=0.666, z=z, return_CPU=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        a = m.infer(mel, sigma=0.666, z=z, return_CPU=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out[tag] = dict(ms_per_call=ms, samples_per_s=a.shape[1] / ms * 1e3, xrt_22k=a.shape[1] / ms * 1e3 / 22050,
                    finite=bool(torch.isfinite(a).all()), general=bool(m.general), precision=m.precision)
    print(tag, out[tag], flush=True)
    del m
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(frames=frames, model="ax WaveGlow 12 flows, 8 x 256, n_group 8, 1 utterance", runs=out),
          open(os.path.join(ROOT, "gpurun_out", "general_mode_timing.json"), "w"), indent=1)
