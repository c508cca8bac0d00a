"""Golden vectors for the ax models' conditioning front-end / output filters from the UNMODIFIED
reference (`CookieTTS/_4_mtw/waveglow/efficient_model_ax.py::WaveGlow` with speaker_embed, cond_layers,
transposed_conv_scales, group_conv_output_dim, preempthasis, preceived_vol_scaling switched on).

    python oracle/make_golden_ax_frontend.py        (build container only; needs /root/reference)

Same harness as make_golden_waveflow.py (iso226 stub, `Tensor.normal_` patched to return the pre-drawn z) plus
`np.product = np.prod`: the reference calls `np.product` (efficient_model_ax.py:124, glow_ax.py:230), which numpy 2
removed.  Each case loads the seeded synthetic checkpoint with strict=True (pins the key layout) and stores the
reference's own `infer` output in fp32 and fp64 in tests/golden/axfe_*.npz.
"""
from __future__ import annotations

import dataclasses
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ax_frontend_oracle import FrontEndConfig, synthetic_frontend_state_dict  # noqa: E402
from oracle.make_golden import InjectedNormal  # noqa: E402
from oracle.make_golden_waveflow import load_reference_ax, reference_kwargs, reference_kwargs_ax1d  # noqa: E402
from oracle.waveflow_oracle import WaveFlowConfig, synthetic_state_dict as wf_sd  # noqa: E402
from oracle.waveglow_ax_oracle import AxConfig, synthetic_state_dict as ax_sd  # noqa: E402

SMALL = dict(n_mel_channels=8, n_flows=4, n_group=8, n_early_every=2, n_early_size=2, n_layers=3, n_channels=16,
             win_length=64, hop_length=16)

# notebook cell 2: waveglow_config / WN_config / data_config (hop_length 600, win_length 2400), see the cases below
NB = dict(n_mel_channels=160, n_flows=48, n_group=24, n_early_every=16, n_early_size=2, n_layers=8, n_channels=256,
          win_length=2400, hop_length=600, channel_mixing="permuteheight", mix_first=False, wn_speaker_embed_dim=96,
          upsample_first=False)
NB_FE = dict(speaker_embed=96, cond_layers=3, cond_activation_func="lrelu", negative_slope=0.5, cond_hidden_channels=256,
             cond_output_channels=256, cond_residual=True, cond_res_rezero=True, cond_padding_mode="replicate",
             cond_kernel_size=2)

# name: (kind, model kwargs, front-end kwargs, batch, frames, sigma, weight seed, input seed)
CASES = {
    "axfe_speaker_cond": ("ax", SMALL, dict(speaker_embed=4, cond_layers=2, cond_hidden_channels=10, cond_output_channels=12,
                                            cond_kernel_size=2, cond_residual="1x1conv", cond_padding_mode="replicate",
                                            cond_res_rezero=True, cond_activation_func="relu", negative_slope=0.2),
                          3, 7, 0.8, 41, 1),
    "axfe_tconv_crop": ("ax", dict(SMALL, n_group=4), dict(transposed_conv_hidden_dim=10, transposed_conv_kernel_size=[4, 4],
                                                           transposed_conv_scales=[2, 2], transposed_conv_output_dim=8,
                                                           transposed_conv_residual=True, transposed_conv_residual_linear=True,
                                                           transposed_conv_res_rezero=True),
                        2, 6, 0.9, 42, 2),
    "axfe_tconv_interp_group": ("ax", dict(SMALL, n_group=4, upsample_mode="nearest"),
                                dict(cond_layers=1, cond_output_channels=12, cond_activation_func="lrelu",
                                     transposed_conv_hidden_dim=10, transposed_conv_kernel_size=5, transposed_conv_scales=[3],
                                     transposed_conv_output_dim=12, group_conv_output_dim=6, group_conv_groupped=True),
                                2, 5, 1.0, 43, 3),
    "axfe_post": ("ax", SMALL, dict(cond_layers=1, cond_kernel_size=3, cond_residual=True, cond_padding_mode="reflect",
                                    cond_activation_func="tanh", group_conv_output_dim=5, group_conv_groupped=False,
                                    shift_spect=5.0, scale_spect=0.25, preempthasis=0.97, preceived_vol_scaling=True),
                  2, 6, 0.7, 44, 4),
    # tensor-core sized 1-D model (12 flows, 8 x 256) with the speaker + cond-net + upsample-net front-end
    "axfe_256": ("ax", dict(), dict(speaker_embed=16, cond_layers=2, cond_hidden_channels=64, cond_output_channels=48,
                                    cond_residual="1x1conv", cond_res_rezero=True, cond_activation_func="relu", negative_slope=0.1,
                                    transposed_conv_hidden_dim=64, transposed_conv_kernel_size=[4, 8, 4],
                                    transposed_conv_scales=[2, 4, 4], transposed_conv_output_dim=96),
                 2, 5, 0.666, 45, 5),
    # WaveFlow (8 flows, h=16, 8 x 128, 3x3) with the same kind of front-end
    "axfe_waveflow": ("wf", dict(), dict(speaker_embed=8, cond_layers=2, cond_hidden_channels=32, cond_output_channels=40,
                                         cond_residual="1x1conv", cond_activation_func="sigmoid",
                                         transposed_conv_hidden_dim=32, transposed_conv_kernel_size=[4, 8],
                                         transposed_conv_scales=[2, 8], transposed_conv_output_dim=64, preempthasis=0.9),
                      2, 4, 0.666, 46, 6),
    # seperable_conv=True: in_layers are Sequential(depthwise, pointwise) (what the author's checkpoints use)
    "axfe_separable": ("ax", dict(SMALL, seperable_conv=True), dict(speaker_embed=3), 2, 6, 0.8, 47, 7),
    "axfe_separable_256": ("ax", dict(seperable_conv=True), dict(), 1, 5, 0.666, 48, 8),
    "axfe_waveflow_separable": ("wf", dict(seperable_conv=True), dict(speaker_embed=8), 1, 4, 0.666, 49, 9),
    # ---- the layout of `scripts/WaveGlowFlow Inference Speed Testing.ipynb` cell 2 (the one config the reference records a
    # speed for): n_group 24 with early outputs, 'permute' mixing after the coupling, WN-level speaker embeddings,
    # upsample_first=False, residual ReZero cond net with 'lrelu' and replicate padding.  Small widths (fp32 path) ...
    "axfe_nb_small": ("ax", dict(NB, n_mel_channels=8, n_flows=6, n_early_every=2, n_layers=3, n_channels=16, win_length=192,
                                 hop_length=48, wn_speaker_embed_dim=5),
                      dict(NB_FE, speaker_embed=4, cond_hidden_channels=10), 3, 7, 0.8, 50, 10),
    # ... the notebook's WN (8 x 256, 160 mels, 96 + 96 speaker dims, hop 600) with 6 of its 48 flows ...
    "axfe_nb_256": ("ax", dict(NB, n_flows=6, n_early_every=2), dict(NB_FE), 2, 4, 0.9, 51, 11),
    # ... and the notebook's model itself (48 flows, early outputs every 16) on a short clip
    "axfe_notebook": ("ax", dict(NB), dict(NB_FE), 1, 5, 1.0, 52, 12),
}


def build_case(kind, mkw, fkw):
    if kind == "ax":
        cfg = AxConfig(**mkw)
        base = reference_kwargs_ax1d(cfg)
    else:
        cfg = WaveFlowConfig(**mkw)
        base = reference_kwargs(cfg)
    fe = FrontEndConfig(n_mel_channels=cfg.n_mel_channels, n_flows=cfg.n_flows, n_group=cfg.n_group,
                        hop_length=cfg.hop_length, upsample_mode=cfg.upsample_mode,
                        upsample_first=getattr(cfg, "upsample_first", True), **fkw)
    kw = dict(base)
    for f in dataclasses.fields(fe):
        if f.name in ("n_mel_channels", "n_flows", "n_group", "hop_length", "upsample_mode", "upsample_first"):
            continue
        kw[f.name] = getattr(fe, f.name)
    return cfg, fe, kw


def state_dict_for(kind, cfg, fe, seed):
    sd = (ax_sd if kind == "ax" else wf_sd)(cfg, seed, cond_in_channels=fe.wn_cond_in())
    sd.update(synthetic_frontend_state_dict(fe, seed + 1000))
    return sd


def main():
    np.product = np.prod                                   # removed in numpy 2; the reference still calls it
    Model = load_reference_ax()
    outdir = os.path.join(ROOT, "tests", "golden")
    only = set(sys.argv[1:])
    for name, (kind, mkw, fkw, batch, frames, sigma, wseed, iseed) in CASES.items():
        if only and name not in only:
            continue
        cfg, fe, kw = build_case(kind, mkw, fkw)
        sd = state_dict_for(kind, cfg, fe, wseed)
        rs = np.random.RandomState(iseed)
        mel = np.clip(rs.standard_normal((batch, cfg.n_mel_channels, frames)) * 2.0 - 5.0, -11.5129, 2.0).astype(np.float32)
        z = rs.standard_normal((batch, frames * cfg.hop_length)).astype(np.float32)
        spk = rs.randint(0, 512, size=(batch,)).astype(np.int64)
        outs = {}
        for dt, tag in ((torch.float32, "fp32"), (torch.float64, "fp64")):
            model = Model(**kw)
            model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()}, strict=True)
            model = model.eval().to(dt)
            if dt == torch.float64 and kind == "ax" and cfg.channel_mixing == "1x1conv":
                for conv in model.convinv:                 # W_inverse is always created fp32 (efficient_modules.py:271-275)
                    conv.W_inverse = conv.weight.squeeze().double().inverse().unsqueeze(-1)
            with torch.no_grad(), InjectedNormal([torch.from_numpy(z)]):
                aud = model.infer(torch.from_numpy(mel).to(dt), speaker_ids=torch.from_numpy(spk), sigma=sigma)
            outs[tag] = aud.numpy()
        err = np.abs(outs["fp32"] - outs["fp64"]).max()
        print(f"{name}: infer {outs['fp64'].shape} rms {np.sqrt((outs['fp64'] ** 2).mean()):.3f} max {np.abs(outs['fp64']).max():.2f} "
              f"fp32-vs-fp64 {err:.2e}")
        np.savez_compressed(os.path.join(outdir, f"{name}.npz"), kind=kind, model=json.dumps(mkw), frontend=json.dumps(fkw),
                            batch=batch, frames=frames, sigma=sigma, weight_seed=wseed, input_seed=iseed, mel=mel, z=z,
                            speaker_ids=spk, infer_ref_fp32=outs["fp32"], infer_ref_fp64=outs["fp64"])


if __name__ == "__main__":
    main()
