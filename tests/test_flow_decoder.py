"""Mel-domain flow decoders: oracle vs the reference's golden vectors (CPU), module layout vs the reference
state_dict (CPU), and cwg_fd_inverse vs the goldens (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

from cookietts_b200.flow_decoder import FlowDecoder
from oracle.flow_decoder_oracle import FlowDecoderConfig, synthetic_state_dict, inverse, snr_db
from oracle.make_golden_flow_decoder import hparams_for

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["fd_small", "fd_flowtts", "fd_layers", "fd_untts", "fd_separable", "fd_condstack"]


def load(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = FlowDecoderConfig(**json.loads(str(g["config"])))
    return cfg, synthetic_state_dict(cfg, int(g["weight_seed"])), g


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    cfg, sd, g = load(name)
    out = inverse(sd, cfg, g["z"].astype(np.float64) * float(g["sigma"]), g["cond"])
    assert np.abs(out - g["out_ref_fp64"]).max() < 1e-9
    assert np.abs(out - g["out_ref_fp32"]).max() < 1e-3


@pytest.mark.parametrize("name", NAMES)
def test_module_has_the_reference_state_dict_layout(name):
    cfg, sd, g = load(name)
    m = FlowDecoder(hparams_for(cfg, str(g["variant"])))
    own = m.state_dict()
    assert set(own.keys()) == set(sd.keys())
    for k, v in sd.items():
        assert tuple(own[k].shape) == tuple(v.shape), k
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    with pytest.raises(NotImplementedError):
        m.forward(None, None)
    with pytest.raises(RuntimeError):                      # no CPU fallback
        m.inverse(torch.zeros(1, cfg.n_mel_channels, 4), torch.zeros(1, cfg.cond_channels, cfg.n_mel_channels * 4 // cfg.n_group))


def test_unsupported_variants_raise():
    cfg = FlowDecoderConfig(n_mel_channels=16, n_group=16, n_flows=2, cond_channels=8, wn_n_channels=16)
    hp = hparams_for(cfg, "flowtts")
    hp.cond_layers = 2
    with pytest.raises(NotImplementedError):
        FlowDecoder(hp)
    hp = hparams_for(cfg, "flowtts")
    hp.wn_cond_act_func = "relu"                   # NameError in the reference's own constructor (glow.py:96)
    with pytest.raises(NotImplementedError):
        FlowDecoder(hp)
    hp = hparams_for(cfg, "flowtts")
    hp.wn_cond_layers, hp.wn_cond_act_func = 2, "tanh"      # WN cond stacks are served (fd_condstack golden)
    assert FlowDecoder(hp).cond_external


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_inverse_matches_reference(name):
    cfg, sd, g = load(name)
    m = FlowDecoder(hparams_for(cfg, str(g["variant"])))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    m = m.cuda().eval()
    z, cond, sigma = torch.from_numpy(g["z"]).cuda(), torch.from_numpy(g["cond"]).cuda(), float(g["sigma"])
    out, logdet = m.inverse(z * sigma, cond)
    assert logdet is None
    out = out.cpu().numpy()
    ref = g["out_ref_fp64"]
    assert out.shape == ref.shape and np.isfinite(out).all()
    assert np.abs(out - ref).max() <= 2e-4 and snr_db(ref, out) >= 90.0          # fp32 arithmetic, other summation order
    assert np.abs(out - g["out_ref_fp32"]).max() <= 2e-4
    out2 = m.infer(cond, sigma=sigma, z=z).cpu().numpy()                          # infer = draw + inverse (glow.py:345-352)
    assert np.array_equal(out, out2)
    assert m.infer(cond, sigma=sigma).shape == z.shape                            # internal draw
    assert m.launch_count() > 0


@pytest.mark.gpu
def test_gpu_inverse_inverts_the_reference_forward_direction():
    """Size-independent property at a realistic size (flowtts defaults, 16 x 800 frames): pushing the decoder's output
    back through the forward direction (evaluated with torch ops on the GPU from the same weights) returns the latent."""
    cfg = FlowDecoderConfig(end_std=0.01)
    sd = synthetic_state_dict(cfg, 5)
    m = FlowDecoder(hparams_for(cfg, "flowtts"))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(0)
    B, frames = 16, 800
    z = torch.randn(B, cfg.n_mel_channels, frames, device="cuda", generator=g)
    cond = torch.randn(B, cfg.cond_channels, frames, device="cuda", generator=g)
    mel, _ = m.inverse(z, cond)
    assert torch.isfinite(mel).all()
    # forward direction (glow.py:258-295) in fp64 torch ops
    from cookietts_b200.flow_decoder import _eff
    sdt = {k: torch.from_numpy(v).cuda() for k, v in sd.items()}
    x = mel.double().view(B, cfg.n_group, -1)
    cd = cond.double()
    outs, C = [], cfg.wn_n_channels
    for k, n_rem in enumerate(cfg.flow_channels()):
        if k % cfg.n_early_every == 0 and k > 0:
            outs.append(x[:, :cfg.n_early_size]); x = x[:, cfg.n_early_size:]
        W = sdt[f"convinv.{k}.weight"].double()
        x = torch.nn.functional.conv1d(x, W)                                       # mix_first
        n_half = n_rem // 2
        x0, x1 = x[:, :n_half], x[:, n_half:]
        p = f"WN.{k}.WN."
        h = torch.nn.functional.conv1d(x0, _eff(sdt, p + "start"), sdt[p + "start.bias"].double())
        c = torch.nn.functional.conv1d(cd, _eff(sdt, p + "cond_layers.0"), sdt[p + "cond_layers.0.bias"].double())
        pre = torch.nn.functional.conv1d(h, _eff(sdt, p + "in_layers.0"), sdt[p + "in_layers.0.bias"].double(), padding=1) + c
        h = h + torch.tanh(pre[:, :C]) * torch.sigmoid(pre[:, C:])
        e = torch.nn.functional.conv1d(h, sdt[p + "end.weight"].double(), sdt[p + "end.bias"].double())
        x = torch.cat([x0, x1 * e[:, :n_half].exp() + e[:, n_half:]], 1)
    outs.append(x)
    z_back = torch.cat(outs, 1).view(B, cfg.n_mel_channels, -1)
    assert float((z_back - z.double()).abs().max()) < 2e-3
