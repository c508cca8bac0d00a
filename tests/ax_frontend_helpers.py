"""Shared by the ax front-end tests: rebuild a golden case's config / checkpoint and run the oracle end to end."""
import json
import os

import numpy as np

from oracle.ax_frontend_oracle import FrontEndConfig, frontend, post_filter, synthetic_frontend_state_dict
from oracle.waveflow_oracle import WaveFlowConfig, synthetic_state_dict as wf_sd, inverse as wf_inverse
from oracle.waveglow_ax_oracle import AxConfig, synthetic_state_dict as ax_sd, inverse as ax_inverse
from tests.helpers import GOLDEN_DIR

CASES = ["axfe_speaker_cond", "axfe_tconv_crop", "axfe_tconv_interp_group", "axfe_post", "axfe_256", "axfe_waveflow",
         "axfe_separable", "axfe_separable_256", "axfe_waveflow_separable", "axfe_nb_small", "axfe_nb_256"]
# the notebook's own 48-flow model: ~1 GB of synthetic weights, GPU parity only (tests/test_gpu_ax_frontend.py)
BIG_CASES = ["axfe_notebook"]


def load_case(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    kind = str(g["kind"])
    mkw, fkw = json.loads(str(g["model"])), json.loads(str(g["frontend"]))
    cfg = (AxConfig if kind == "ax" else WaveFlowConfig)(**mkw)
    fe = FrontEndConfig(n_mel_channels=cfg.n_mel_channels, n_flows=cfg.n_flows, n_group=cfg.n_group,
                        hop_length=cfg.hop_length, upsample_mode=cfg.upsample_mode,
                        upsample_first=getattr(cfg, "upsample_first", True), **fkw)
    seed = int(g["weight_seed"])
    sd = (ax_sd if kind == "ax" else wf_sd)(cfg, seed, cond_in_channels=fe.wn_cond_in())
    sd.update(synthetic_frontend_state_dict(fe, seed + 1000))
    return kind, cfg, fe, sd, g


def oracle_infer(kind, cfg, fe, sd, mel, z, speaker_ids, sigma, dtype=np.float64, artifact_trimming=1):
    """efficient_model_ax.py:359-388 with the latent passed in."""
    spect = np.asarray(mel, dtype)
    spect = np.concatenate([spect, np.zeros(spect.shape[:2] + (artifact_trimming,), dtype)], axis=2)
    samples = (spect.shape[2] - 1) * cfg.hop_length
    samples -= samples % cfg.n_group
    zz = np.asarray(z, dtype)[:, :samples] * dtype(sigma)
    cond_up = frontend(sd, fe, spect, speaker_ids, samples // cfg.n_group, dtype)
    if kind == "ax":
        audio = ax_inverse(sd, cfg, zz, None, dtype, cond_up=cond_up, speaker_ids=speaker_ids)
    else:
        audio = wf_inverse(sd, cfg, zz, None, dtype, cond_up=cond_up)
    audio = post_filter(fe, audio, dtype)
    return audio[:, :-artifact_trimming * cfg.hop_length]


def module_kwargs(kind, cfg, fe):
    """Constructor kwargs of the reference ax model (and of the drop-in modules) for a case."""
    import dataclasses
    if kind == "ax":
        wn = dict(n_layers=cfg.n_layers, n_channels=cfg.n_channels, kernel_size=cfg.kernel_size, kernel_size_w=None,
                  n_layers_dilations_w=None, n_layers_dilations_h=1, speaker_embed_dim=cfg.wn_speaker_embed_dim, rezero=False,
                  cond_layers=1, cond_activation_func="none", negative_slope=None, cond_hidden_channels=256, cond_kernel_size=1,
                  cond_padding_mode="zeros", seperable_conv=cfg.seperable_conv, res_skip=True, merge_res_skip=False,
                  upsample_mode=cfg.upsample_mode)
        kw = dict(n_early_every=cfg.n_early_every, n_early_size=cfg.n_early_size, channel_mixing=cfg.channel_mixing,
                  mix_first=cfg.mix_first, waveflow=False)
    else:
        wn = dict(n_layers=cfg.n_layers, n_channels=cfg.n_channels, kernel_size_w=cfg.kernel_size_w,
                  kernel_size_h=cfg.kernel_size_h, n_layers_dilations_w=None, n_layers_dilations_h=1,
                  speaker_embed_dim=0, rezero=False, cond_layers=1, cond_activation_func="none", negative_slope=None,
                  cond_hidden_channels=256, cond_kernel_size=1, cond_padding_mode="zeros", seperable_conv=cfg.seperable_conv,
                  res_skip=True, merge_res_skip=False, upsample_mode=cfg.upsample_mode)
        kw = dict(n_early_every=cfg.n_flows * 2, n_early_size=2, channel_mixing="permuteheight", mix_first=True, waveflow=True)
    kw.update(n_mel_channels=cfg.n_mel_channels, n_flows=cfg.n_flows, n_group=cfg.n_group, memory_efficient=0.0,
              spect_scaling=False, upsample_mode="normal", upsample_first=True, WN_config=wn, win_length=cfg.win_length,
              hop_length=cfg.hop_length, sampling_rate=22050)
    kw["upsample_first"] = bool(getattr(cfg, "upsample_first", True))
    for f in dataclasses.fields(fe):
        if f.name not in ("n_mel_channels", "n_flows", "n_group", "hop_length", "upsample_mode", "upsample_first"):
            kw[f.name] = getattr(fe, f.name)
    return kw
