"""Seeded synthetic checkpoints and inputs with the reference models' exact state_dict layouts.

A data utility, not an algorithm: `bench.py`, `__graft_entry__.smoke()`, the golden generators and the tests all need
"a reference-shaped checkpoint with non-zero `end`" (the reference zero-initialises `end`, glow.py:141-144, which makes
the WN a no-op and every parity check vacuous - SURVEY Appendix B.1) regenerated identically on every machine from a
seed.  The oracle modules re-export these names; nothing here evaluates the model.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np


@dataclass
class ModelConfig:
    """Constructor arguments of the reference `WaveGlow` (glow.py:226-227) that the
    inverse pass depends on, plus `WN_config` (glow.py:116-117)."""
    n_mel_channels: int = 80
    n_flows: int = 12
    n_group: int = 8
    n_early_every: int = 4
    n_early_size: int = 2
    win_length: int = 1024
    hop_length: int = 256
    n_layers: int = 8
    n_channels: int = 256
    kernel_size: int = 3
    speaker_embed_dim: int = 0
    rezero: bool = False
    cond_hidden: int = 256  # literal `hidden_dim = 256`, glow.py:153
    end_scale: float = 1.0  # synthetic checkpoints only: multiplier on `end` ~ N(0, 0.02) ("trained-scale" stress cases)

    def flow_channels(self) -> List[Tuple[int, int]]:
        """(n_remaining_channels, n_half) per flow k, glow.py:251-264."""
        out = []
        n_half = self.n_group // 2
        n_rem = self.n_group
        for k in range(self.n_flows):
            if k % self.n_early_every == 0 and k > 0:
                n_half -= self.n_early_size // 2
                n_rem -= self.n_early_size
            out.append((n_rem, n_half))
        return out



# ----------------------------------------------------------------------------------------
# Deterministic synthetic checkpoints (shared by the golden generator, tests and bench)
# ----------------------------------------------------------------------------------------

def synthetic_state_dict(cfg: ModelConfig, seed: int = 1234) -> Dict[str, np.ndarray]:
    """A state_dict with the exact key/shape layout of the reference model
    (glow.py:136-186,238-241,74-83), filled from a seeded
    `numpy.random.RandomState` so every machine regenerates identical weights.
    `end` is NOT zero (the reference zero-inits it, glow.py:141-144, which would make
    the WN a no-op and parity vacuous)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}

    def uni(shape, bound):
        return rs.uniform(-bound, bound, size=shape).astype(np.float32)

    def wn_conv(prefix, co, ci, k):
        bound = 1.0 / np.sqrt(ci * k)
        v = uni((co, ci, k), bound)
        norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
        sd[prefix + ".bias"] = uni((co,), bound)
        sd[prefix + ".weight_g"] = (norm * rs.uniform(0.8, 1.2, size=norm.shape)).astype(np.float32)
        sd[prefix + ".weight_v"] = v

    M, G, C, L = cfg.n_mel_channels, cfg.n_group, cfg.n_channels, cfg.n_layers
    bound = 1.0 / np.sqrt(M * cfg.win_length / cfg.hop_length)
    sd["upsample.weight"] = uni((M, M, cfg.win_length), bound)
    sd["upsample.bias"] = uni((M,), bound)
    for k, (n_rem, n_half) in enumerate(cfg.flow_channels()):
        q1, _ = np.linalg.qr(rs.standard_normal((n_rem, n_rem)))
        q2, _ = np.linalg.qr(rs.standard_normal((n_rem, n_rem)))
        W = q1 @ np.diag(rs.uniform(0.7, 1.4, size=n_rem)) @ q2
        sd[f"convinv.{k}.conv.weight"] = W.astype(np.float32)[:, :, None]
        p = f"WN.{k}"
        wn_conv(p + ".start", C, n_half, 1)
        wn_conv(p + ".cond_layers.0", cfg.cond_hidden, M * G + cfg.speaker_embed_dim, 1)
        wn_conv(p + ".cond_layers.1", cfg.cond_hidden, cfg.cond_hidden, 1)
        wn_conv(p + ".cond_layers.2", 2 * C * L, cfg.cond_hidden, 1)
        for i in range(L):
            wn_conv(p + f".in_layers.{i}", 2 * C, C, cfg.kernel_size)
            wn_conv(p + f".res_skip_layers.{i}", 2 * C if i < L - 1 else C, C, 1)
            if cfg.rezero:
                sd[p + f".alpha_i.{i}"] = (rs.uniform(size=1) * 0.02 + 0.09).astype(np.float32)
        sd[p + ".end.weight"] = (rs.standard_normal((2 * n_half, C, 1)) * 0.02 * cfg.end_scale).astype(np.float32)
        sd[p + ".end.bias"] = (rs.standard_normal((2 * n_half,)) * 0.02 * cfg.end_scale).astype(np.float32)
        if cfg.speaker_embed_dim:       # nn.Embedding(512, E) scaled by 0.05 at init, glow.py:131-134
            sd[p + ".speaker_embed.weight"] = (rs.standard_normal((512, cfg.speaker_embed_dim)) * 0.5).astype(np.float32)
    return sd


def synthetic_inputs(cfg: ModelConfig, batch: int, t_mel: int, seed: int = 0):
    """Synthetic mel = clamp(N(-5, 2^2), -11.5129, 2.0) and z ~ N(0,1)."""
    rs = np.random.RandomState(seed)
    mel = np.clip(rs.standard_normal((batch, cfg.n_mel_channels, t_mel)) * 2.0 - 5.0, -11.5129, 2.0)
    z = rs.standard_normal((batch, t_mel * cfg.hop_length))
    return mel.astype(np.float32), z.astype(np.float32)



# ----------------------------------------------------------------------------------------
# WaveFlow (ax model, waveflow=True)
# ----------------------------------------------------------------------------------------

@dataclass
class WaveFlowConfig:
    n_mel_channels: int = 80
    n_flows: int = 8
    n_group: int = 16            # squeeze height h
    n_layers: int = 8
    n_channels: int = 128
    kernel_size_w: int = 3
    kernel_size_h: int = 3
    win_length: int = 1024
    hop_length: int = 256
    upsample_mode: str = "linear"   # WN_config['upsample_mode'] used by the model-level interpolate
    seperable_conv: bool = False    # in_layer = Sequential(depthwise, pointwise), glow_ax.py:525-531
    # ---- WN_config variants (the fp32 CUDA-core mode runs them; the defaults are what the tensor-core kernels take)
    gated_unit: str = "GTU"         # glow_ax.py:168-198
    dilations_w: object = None      # n_layers_dilations_w: None (2^i), one int, or a list (glow_ax.py:506-514)
    dilations_h: object = 1         # n_layers_dilations_h: one int or a list (:509-513)
    res_skip: bool = True
    merge_res_skip: bool = False
    wn_speaker_embed_dim: int = 0   # WN_config['speaker_embed_dim'], glow_ax.py:464-466
    upsample_first: bool = True     # False: the WN interpolates its cond-layer output (glow_ax.py:578-579)
    n_early_every: int = 0          # early outputs (efficient_model_ax.py:151-167); 0 = none
    n_early_size: int = 2
    mix_first: bool = True          # False: PermuteHeight before the coupling (efficient_model_ax.py:326-337)
    channel_mixing: str = "permuteheight"   # or "1x1conv": InvertibleConv1x1 over the height rows
    # WN cond path in its general form (glow_ax.py:476-505,:565-579); the defaults are the one linear 1x1 layer
    wn_cond_layers: int = 1
    wn_cond_hidden_channels: int = 256
    wn_cond_kernel_size: int = 1            # the conv has 2k - 1 taps
    wn_cond_padding_mode: str = "zeros"
    wn_cond_activation_func: str = "none"
    wn_negative_slope: object = None
    wn_cond_out_activation_func: bool = True
    wn_tconv_scales: object = None          # WN-level TransposedUpsampleNet (upsample_first=False)
    wn_tconv_hidden_dim: int = 256
    wn_tconv_kernel_size: object = 4

    def flow_rows(self):
        """Height rows each flow works on."""
        out, n = [], self.n_group
        for k in range(self.n_flows):
            if self.n_early_every and k % self.n_early_every == 0 and k > 0:
                n -= self.n_early_size
            out.append(n)
        return out

    def dilation_w(self, i: int) -> int:
        if self.dilations_w is None:
            return 2 ** i
        return int(self.dilations_w) if isinstance(self.dilations_w, int) else int(self.dilations_w[i])

    def dilation_h(self, i: int) -> int:
        return int(self.dilations_h) if isinstance(self.dilations_h, int) else int(self.dilations_h[i])



def waveflow_state_dict(cfg: WaveFlowConfig, seed: int = 1234, cond_in_channels=None) -> Dict[str, np.ndarray]:
    """Seeded checkpoint with the reference ax/WaveFlow key layout (probe-printed: WN.{k}.WN.*,
    4-D conv weights, no convinv parameters for permuteheight); `end` non-zero."""
    cin = cond_in_channels or cfg.n_mel_channels
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    C, L, kh, kw = cfg.n_channels, cfg.n_layers, cfg.kernel_size_h, cfg.kernel_size_w

    def wn(prefix, shape, fan_in):
        bound = 1.0 / np.sqrt(fan_in)
        v = rs.uniform(-bound, bound, size=shape).astype(np.float32)
        norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
        sd[prefix + ".bias"] = rs.uniform(-bound, bound, size=(shape[0],)).astype(np.float32)
        sd[prefix + ".weight_g"] = (norm * rs.uniform(0.8, 1.2, size=norm.shape)).astype(np.float32)
        sd[prefix + ".weight_v"] = v

    for k in range(cfg.n_flows):
        p = f"WN.{k}.WN."
        if cfg.channel_mixing == "1x1conv":
            n_rem = cfg.flow_rows()[k]
            q1, _ = np.linalg.qr(rs.standard_normal((n_rem, n_rem)))
            q2, _ = np.linalg.qr(rs.standard_normal((n_rem, n_rem)))
            sd[f"convinv.{k}.weight"] = (q1 @ np.diag(rs.uniform(0.7, 1.4, size=n_rem)) @ q2).astype(np.float32)[:, :, None]
        for i in range(L):
            if cfg.seperable_conv:
                wn(p + f"in_layers.{i}.0", (C, 1, kh, kw), kh * kw)
                wn(p + f"in_layers.{i}.1", (2 * C, C, 1, 1), C)
            else:
                wn(p + f"in_layers.{i}", (2 * C, C, kh, kw), C * kh * kw)
            if cfg.res_skip:
                wn(p + f"res_skip_layers.{i}", (2 * C if (i < L - 1 and not cfg.merge_res_skip) else C, C, 1, 1), C)
        wn(p + "start", (C, 1, 1, 1), 1)
        sd[p + "end.weight"] = (rs.standard_normal((2, C, 1, 1)) * 0.02).astype(np.float32)
        sd[p + "end.bias"] = (rs.standard_normal((2,)) * 0.02).astype(np.float32)
        kc = 2 * cfg.wn_cond_kernel_size - 1
        cond_out = cfg.wn_tconv_hidden_dim if cfg.wn_tconv_scales else 2 * C * L                       # glow_ax.py:481
        dims = [cin + cfg.wn_speaker_embed_dim] + [cfg.wn_cond_hidden_channels] * (cfg.wn_cond_layers - 1) + [cond_out]
        for i, (di, do) in enumerate(zip(dims[:-1], dims[1:])):
            wn(p + f"cond_layers.{i}", (do, di, kc), di * kc)
        if cfg.wn_speaker_embed_dim:
            sd[p + "speaker_embed.weight"] = rs.standard_normal((512, cfg.wn_speaker_embed_dim)).astype(np.float32)
        if cfg.wn_tconv_scales:                                          # TransposedUpsampleNet(hidden, 2CL, hidden, ...), :468-473
            n = len(cfg.wn_tconv_scales)
            for i, sc in enumerate(cfg.wn_tconv_scales):
                kk = cfg.wn_tconv_kernel_size[i] if isinstance(cfg.wn_tconv_kernel_size, (list, tuple)) else cfg.wn_tconv_kernel_size
                ci, co = cfg.wn_tconv_hidden_dim, (2 * C * L if i == n - 1 else cfg.wn_tconv_hidden_dim)
                bound = 1.0 / np.sqrt(ci * kk / sc)
                sd[p + f"upsample_net.t_convs.{2 * i}.weight"] = rs.uniform(-bound, bound, size=(ci, co, kk)).astype(np.float32)
                sd[p + f"upsample_net.t_convs.{2 * i}.bias"] = rs.uniform(-bound, bound, size=(co,)).astype(np.float32)
    return sd


def waveflow_reference_kwargs(cfg: WaveFlowConfig) -> dict:
    wn = dict(n_layers=cfg.n_layers, n_channels=cfg.n_channels, kernel_size_w=cfg.kernel_size_w,
              kernel_size_h=cfg.kernel_size_h, n_layers_dilations_w=cfg.dilations_w, n_layers_dilations_h=cfg.dilations_h,
              speaker_embed_dim=cfg.wn_speaker_embed_dim, rezero=False, cond_layers=cfg.wn_cond_layers,
              cond_activation_func=cfg.wn_cond_activation_func, negative_slope=cfg.wn_negative_slope,
              cond_hidden_channels=cfg.wn_cond_hidden_channels, cond_kernel_size=cfg.wn_cond_kernel_size,
              cond_padding_mode=cfg.wn_cond_padding_mode, seperable_conv=cfg.seperable_conv,
              res_skip=cfg.res_skip, merge_res_skip=cfg.merge_res_skip, upsample_mode=cfg.upsample_mode, gated_unit=cfg.gated_unit,
              cond_out_activation_func=cfg.wn_cond_out_activation_func)
    if cfg.wn_tconv_scales:
        wn.update(transposed_conv_scales=list(cfg.wn_tconv_scales), transposed_conv_hidden_dim=cfg.wn_tconv_hidden_dim,
                  transposed_conv_kernel_size=cfg.wn_tconv_kernel_size)
    return dict(n_mel_channels=cfg.n_mel_channels, n_flows=cfg.n_flows, n_group=cfg.n_group,
                n_early_every=cfg.n_early_every or cfg.n_flows * 2, n_early_size=cfg.n_early_size, memory_efficient=0.0, spect_scaling=False,
                upsample_mode="normal", upsample_first=cfg.upsample_first, speaker_embed=0, cond_layers=0,
                cond_hidden_channels=256, cond_output_channels=256, cond_kernel_size=1, cond_residual=False,
                cond_padding_mode="zeros", WN_config=wn, win_length=cfg.win_length, hop_length=cfg.hop_length,
                sampling_rate=22050, channel_mixing=cfg.channel_mixing, mix_first=cfg.mix_first, waveflow=True)




def notebook_ax_kwargs() -> dict:
    """Constructor kwargs of the one model the reference records an inference speed for: `waveglow_config` of
    scripts/WaveGlowFlow Inference Speed Testing.ipynb cell 2, merged with its data_config's win / hop length the way
    cell 3 does.  48 flows, n_group 24, early outputs every 16 flows, 8 x 256 WN with its own 96-dim speaker embedding,
    a 3-layer residual ReZero cond net, `upsample_first: false` (every WN interpolates its cond-layer output)."""
    wn = {"n_layers": 8, "n_channels": 256, "kernel_size_w": 3, "n_layers_dilations_w": None, "n_layers_dilations_h": 1,
          "speaker_embed_dim": 96, "rezero": False, "cond_layers": 1, "cond_activation_func": "none", "negative_slope": 0.5,
          "cond_hidden_channels": 256, "cond_padding_mode": "replicate", "seperable_conv": 0, "res_skip": True,
          "merge_res_skip": False, "upsample_mode": "linear", "cond_kernel_size": 1}
    return {"n_mel_channels": 160, "preceived_vol_scaling": False, "waveflow": False, "channel_mixing": "permute",
            "mix_first": False, "n_flows": 48, "n_group": 24, "n_early_every": 16, "n_early_size": 2, "memory_efficient": 0.0,
            "spect_scaling": False, "upsample_mode": "normal", "WN_config": wn, "speaker_embed": 96, "cond_layers": 3,
            "cond_activation_func": "lrelu", "negative_slope": 0.5, "cond_hidden_channels": 256, "cond_output_channels": 256,
            "cond_residual": True, "cond_res_rezero": True, "cond_padding_mode": "replicate", "upsample_first": False,
            "cond_kernel_size": 2, "win_length": 2400, "hop_length": 600}
