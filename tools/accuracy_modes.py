import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from cookietts_b200 import WaveGlow
from tests.helpers import load_golden
from tests.test_cabi_cpu import module_kwargs
from oracle.waveglow_oracle import snr_db
for name in ("config1", "mel20_256"):
    cfg, sd, g = load_golden(name)
    ref = g["audio_ref_fp64"]
    for prec in ("ffma", "bf16x3", "f16f8", "bf16"):
        m = WaveGlow(precision=prec, **module_kwargs(cfg)); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
        out = m.infer(torch.from_numpy(g["mel"]).cuda(), sigma=float(g["sigma"]), z=torch.from_numpy(g["z"]).cuda()).cpu().numpy()
        print(name, prec, "max-abs %.3e" % np.abs(out - ref).max(), "snr %.1f dB" % snr_db(ref, out))
# full length (2 x 10 s) vs the exact-fp32 CUDA-core mode
import bench
from oracle.waveglow_oracle import OracleConfig, synthetic_state_dict
sd = synthetic_state_dict(OracleConfig(), 1234)
torch.manual_seed(0)
mel = (torch.randn(2, 80, 861, device="cuda") * 2 - 5).clamp(-11.5129, 2.0); z = torch.randn(2, 861 * 256, device="cuda")
outs = {}
for prec in ("ffma", "bf16x3", "f16f8"):
    m = WaveGlow(precision=prec, **bench.MODEL_KW); m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
    outs[prec] = m.infer(mel, sigma=0.666, z=z).double()
for prec in ("bf16x3", "f16f8"):
    e = (outs[prec] - outs["ffma"]).abs().max().item()
    snr = 10 * torch.log10(outs["ffma"].pow(2).sum() / (outs[prec] - outs["ffma"]).pow(2).sum()).item()
    print("2 x 10 s vs ffma:", prec, "max-abs %.3e" % e, "snr %.1f dB" % snr, "peak %.2f" % outs["ffma"].abs().max().item())
