"""Small end-to-end calls of every product path for `compute-sanitizer --tool memcheck python tools/memcheck_small.py`."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cookietts_b200 import WaveGlow, FlowDecoder
from tests.helpers import load_golden
from tests.test_cabi_cpu import module_kwargs
from oracle.flow_decoder_oracle import FlowDecoderConfig, synthetic_state_dict as fd_sd
from oracle.make_golden_flow_decoder import hparams_for

for name in ("mel20_256", "speaker256"):
    cfg, sd, g = load_golden(name)
    for prec in ("f16f8", "bf16x3", "bf16", "ffma"):
        m = WaveGlow(precision=prec, graphs=False, **module_kwargs(cfg))
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
        spk = torch.from_numpy(g["speaker_id"]).cuda() if g["speaker_id"].size else None
        out = m.infer(torch.from_numpy(g["mel"]).cuda(), spk, sigma=float(g["sigma"]), z=torch.from_numpy(g["z"]).cuda())
        torch.cuda.synchronize()
        print(name, prec, float(np.abs(out.cpu().numpy() - g["audio_ref_fp64"]).max()), flush=True)
gf = np.load("tests/golden/fd_untts.npz")
cfg = FlowDecoderConfig(**json.loads(str(gf["config"])))
fd = FlowDecoder(hparams_for(cfg, "untts"))
fd.load_state_dict({k: torch.from_numpy(v) for k, v in fd_sd(cfg, int(gf["weight_seed"])).items()}); fd = fd.cuda().eval()
o, _ = fd.inverse(torch.from_numpy(gf["z"]).cuda() * float(gf["sigma"]), torch.from_numpy(gf["cond"]).cuda())
torch.cuda.synchronize()
print("fd_untts", float(np.abs(o.cpu().numpy() - gf["out_ref_fp64"]).max()), flush=True)
# wide group layout (n_group 24), WN-level speaker embeddings (per-utterance gate bias), upsample_first=False; classic wide
from cookietts_b200 import WaveGlowAx
from tests.ax_frontend_helpers import load_case, module_kwargs as ax_kwargs
kind, cfg, fe, sd, g = load_case("axfe_nb_256")
for prec in ("f16f8", "bf16x3", "bf16", "ffma"):
    m = WaveGlowAx(precision=prec, graphs=False, **ax_kwargs(kind, cfg, fe))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
    out = m.infer(torch.from_numpy(g["mel"]).cuda(), speaker_ids=torch.from_numpy(g["speaker_ids"]).cuda(), sigma=float(g["sigma"]),
                  z=torch.from_numpy(g["z"]).cuda())
    torch.cuda.synchronize()
    print("axfe_nb_256", prec, float(np.abs(out.numpy() - g["infer_ref_fp64"]).max()), flush=True)
cfg, sd, g = load_golden("group24_256")
for prec in ("f16f8", "bf16x3", "ffma"):
    m = WaveGlow(precision=prec, graphs=False, **module_kwargs(cfg))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m = m.cuda().eval()
    out = m.infer(torch.from_numpy(g["mel"]).cuda(), None, sigma=float(g["sigma"]), z=torch.from_numpy(g["z"]).cuda())
    torch.cuda.synchronize()
    print("group24_256", prec, float(np.abs(out.cpu().numpy() - g["audio_ref_fp64"]).max()), flush=True)
# round 2, general fp32 modes: WN_config variants of the ax 1-D WN (cwg_axg_flow) and of WN_2d (cwg_wf_ffma.cu)
if True:
    import warnings
    from cookietts_b200 import WaveFlow
    from oracle.make_golden_waveflow import reference_kwargs_ax1d, reference_kwargs
    from oracle.waveglow_ax_oracle import AxConfig, synthetic_state_dict as ax_sd
    from oracle.waveflow_oracle import WaveFlowConfig, synthetic_state_dict as wf_sd
    warnings.simplefilter("ignore")
    for name in ("waveglow_axv_gsirru", "waveglow_axv_merge", "waveglow_axv_noskip", "waveglow_axv_cond", "waveglow_axv_tconv_crop",
                 "waveglow_axv_tconv_interp"):
        g = np.load(f"tests/golden/{name}.npz")
        cfg = AxConfig(**json.loads(str(g["config"])))
        m = WaveGlowAx(precision="ffma", **reference_kwargs_ax1d(cfg))
        m.load_state_dict({k: torch.from_numpy(v) for k, v in ax_sd(cfg, int(g["weight_seed"])).items()}); m = m.cuda().eval()
        ids = torch.from_numpy(g["speaker_ids"]).cuda() if "speaker_ids" in g.files and g["speaker_ids"].size else None
        out = m.infer(torch.from_numpy(g["mel"]).cuda(), speaker_ids=ids, sigma=float(g["sigma"]), z=torch.from_numpy(g["z"]).cuda())
        torch.cuda.synchronize()
        print(name, float(np.abs(out.numpy() - g["infer_ref_fp64"]).max()), flush=True)
    for name in ("waveflow_sep7", "waveflow_v_gate", "waveflow_v_merge", "waveflow_v_noskip", "waveflow_v_speaker", "waveflow_v_early", "waveflow_v_mixlast",
                 "waveflow_v_conv", "waveflow_v_conv_mixlast", "waveflow_v_cond", "waveflow_v_tconv"):
        g = np.load(f"tests/golden/{name}.npz")
        cfg = WaveFlowConfig(**json.loads(str(g["config"])))
        m = WaveFlow(precision="ffma", graphs=False, **reference_kwargs(cfg))
        m.load_state_dict({k: torch.from_numpy(v) for k, v in wf_sd(cfg, int(g["weight_seed"])).items()}); m = m.cuda().eval()
        ids = torch.from_numpy(g["speaker_ids"]).cuda() if "speaker_ids" in g.files and g["speaker_ids"].size else None
        out = m.infer(torch.from_numpy(g["mel"]).cuda(), speaker_ids=ids, sigma=float(g["sigma"]), z=torch.from_numpy(g["z"]).cuda())
        torch.cuda.synchronize()
        print(name, float(np.abs(out.numpy() - g["infer_ref_fp64"]).max()), flush=True)
