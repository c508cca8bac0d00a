"""CPU oracle for the 1-D ("WaveGlow", waveflow=False) configuration of the reference's ax model.

TEST INFRASTRUCTURE ONLY (same rules as waveglow_oracle.py).  numpy restatement of, in
/root/reference/CookieTTS/_4_mtw/waveglow/:
  efficient_model_ax.py:279-357  WaveGlow.inverse (explicit z; early-z split :319-340)
  efficient_model_ax.py:359-388  WaveGlow.infer (zero-frame pad, hop trim)
  efficient_modules.py:94-105    AffineCouplingBlock.inverse (non-memory-efficient branch)
  glow_ax.py:375-418             WN.forward (1-D, one 1x1 cond layer, GTU gate)
  efficient_modules.py:269-286   InvertibleConv1x1.inverse   |  :360-403 PermuteHeight.inverse
  glow_ax.py:284-286,378-381     WN-level speaker embedding   |  :361-373,389 WN._upsample_mels (upsample_first=False)
for the subset the B200 build supports: upsample_first=True with model-level F.interpolate or upsample_first=False
(every WN interpolates its own cond-layer output), channel_mixing '1x1conv' or 'permuteheight', mix_first True or
False, early outputs, n_group <= 32, optional WN-level speaker embedding, and the WN_config variants of the general fp32
mode (include/cwg.h cwg_axg_flow): the 14 gated units glow_ax.py:36-198, listed dilations :331-335, merge_res_skip /
res_skip=False :259-263,:352,:399-414, WN cond stacks of several layers with activations / kernel sizes / padding modes
:297-329,:383-387.  The model-level conditioning front-end (speaker embedding, cond layers, upsample net, group conv) is
restated in ax_frontend_oracle.py.

Parity status: PINNED by oracle/make_golden_waveflow.py (cases `waveglow_ax_*`, `waveglow_axv_*`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from .waveflow_oracle import _w, upsample_cond, permute_height


@dataclass
class AxConfig:
    n_mel_channels: int = 80
    n_flows: int = 12
    n_group: int = 8
    n_early_every: int = 4
    n_early_size: int = 2
    n_layers: int = 8
    n_channels: int = 256
    kernel_size: int = 3
    win_length: int = 1024
    hop_length: int = 256
    upsample_mode: str = "linear"
    channel_mixing: str = "1x1conv"      # or "permuteheight"
    mix_first: bool = True
    seperable_conv: bool = False         # in_layer = Sequential(depthwise, pointwise), glow_ax.py:350-358
    wn_speaker_embed_dim: int = 0        # WN_config['speaker_embed_dim'], glow_ax.py:255,284-286
    upsample_first: bool = True          # False: cond stays at frame rate until WN._upsample_mels, glow_ax.py:389
    # ---- WN_config variants (general fp32 mode); the defaults are the layout of the packed layer kernels
    gated_unit: str = "GTU"              # glow_ax.py:168-198
    dilations_w: Optional[List[int]] = None   # WN_config['n_layers_dilations_w'] (list, or one int for all layers), :331-335
    res_skip: bool = True
    merge_res_skip: bool = False
    wn_cond_layers: int = 1
    wn_cond_hidden_channels: int = 256
    wn_cond_kernel_size: int = 1         # the conv has 2k - 1 taps (:301)
    wn_cond_padding_mode: str = "zeros"
    wn_cond_activation_func: str = "none"
    wn_negative_slope: Optional[float] = None
    wn_cond_out_activation_func: bool = True
    # WN-level TransposedUpsampleNet (upsample_first=False; glow_ax.py:288-295,:361-373): cond stack -> hidden_dim channels ->
    # ConvTranspose1d chain (LeakyReLU(0.4) between, not after) -> 2CL channels -> F.interpolate or centre crop
    wn_tconv_scales: Optional[List[int]] = None
    wn_tconv_hidden_dim: int = 256
    wn_tconv_kernel_size: object = 4     # int or one per scale

    def dilation(self, i: int) -> int:
        if self.dilations_w is None:
            return 2 ** i
        return int(self.dilations_w) if isinstance(self.dilations_w, int) else int(self.dilations_w[i])

    def is_variant(self) -> bool:
        return (self.gated_unit.upper() != "GTU" or self.dilations_w is not None or not self.res_skip or self.merge_res_skip
                or self.wn_cond_layers != 1 or self.wn_cond_kernel_size != 1 or self.wn_cond_activation_func.lower() != "none"
                or bool(self.wn_tconv_scales) or self.n_group > 32)

    def flow_channels(self) -> List[int]:
        out, n_rem = [], self.n_group
        for k in range(self.n_flows):
            if k % self.n_early_every == 0 and k > 0:
                n_rem -= self.n_early_size
            out.append(n_rem)
        return out


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _selu(x):
    return 1.0507009873554804934193349852946 * np.where(x > 0, x, 1.6732632423543772848170429916717 * np.expm1(np.minimum(x, 0)))


def _softplus(x):                                            # F.softplus: beta 1, threshold 20
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))


def _lrelu(slope):
    return lambda x: np.where(x > 0, x, slope * x)


_tanhshrink = lambda x: x - np.tanh(x)
_sin16 = lambda x: np.sin(16 * x)                            # in_act[:, :n].mul_(16) then sin, glow_ax.py:113-114
_relu = lambda x: np.maximum(x, 0)
# name -> (unit of the first half, unit of the second half), glow_ax.py:36-166; rrelu(0.01, 0.1) in eval mode has slope 0.055
GATED_UNITS = {
    "GTU": (np.tanh, _sigmoid), "GTRU": (np.tanh, _relu), "GTLRU": (np.tanh, _lrelu(0.01)), "GLU": (lambda x: x, _sigmoid),
    "TTU": (np.tanh, np.tanh), "STU": (np.tanh, _selu), "GTSU": (_tanhshrink, _sigmoid), "SPTU": (np.tanh, _softplus),
    "GSIU": (np.sin, _sigmoid), "GSIRU": (_sin16, _sigmoid), "GTSRU": (_tanhshrink, _relu), "GSIRRU": (_sin16, _relu),
    "GSIRLRU": (_sin16, _lrelu(0.01)), "GSIRRLRU": (_sin16, _lrelu(0.055)),
}


def wn_cond_path(sd, p, cfg, cond_up, T, dtype, speaker_ids=None, crop_2d=False):
    """The cond path of one WN (glow_ax.py:378-389; WN_2d :565-579): WN-level speaker embedding, cond stack, and with
    upsample_first=False the WN's upsample net + crop / interpolation (_upsample_mels :361-373; WN_2d's crop rule :545-553 with
    `crop_2d`).  `cfg` is an AxConfig or a WaveFlowConfig (same field names).  Returns [B, 2CL, T]."""
    C, L = cfg.n_channels, cfg.n_layers
    if cfg.wn_speaker_embed_dim and speaker_ids is not None:             # :378-381
        emb = np.asarray(sd[p + "speaker_embed.weight"], dtype)[np.asarray(speaker_ids)]
        cond_up = np.concatenate([cond_up, np.repeat(emb[:, :, None], cond_up.shape[2], axis=2)], axis=1)
    if cfg.wn_cond_layers == 1 and cfg.wn_cond_kernel_size == 1 and cfg.wn_cond_activation_func.lower() == "none" and not cfg.wn_tconv_scales:
        spect = np.einsum("oc,bct->bot", _w(sd, p + "cond_layers.0", dtype)[:, :, 0], cond_up, optimize=True) \
            + np.asarray(sd[p + "cond_layers.0.bias"], dtype)[None, :, None]
    else:                                                                # general cond stack, :297-329 / :383-387
        from .ax_frontend_oracle import conv1d, _act
        spect = cond_up
        pad = (2 * cfg.wn_cond_kernel_size - 1 - 1) // 2
        for i in range(cfg.wn_cond_layers):
            spect = conv1d(spect, _w(sd, p + f"cond_layers.{i}", dtype), np.asarray(sd[p + f"cond_layers.{i}.bias"], dtype),
                           pad, cfg.wn_cond_padding_mode)
            if cfg.wn_cond_activation_func.lower() != "none" and (cfg.wn_cond_out_activation_func or i != cfg.wn_cond_layers - 1):
                spect = _act(spect, cfg.wn_cond_activation_func, cfg.wn_negative_slope)
    if not cfg.upsample_first and cfg.wn_tconv_scales:                   # :389 -> _upsample_mels :361-373 with the WN's upsample net
        from .ax_frontend_oracle import conv_transpose1d
        ks, n = cfg.wn_tconv_kernel_size, len(cfg.wn_tconv_scales)
        for i, sc in enumerate(cfg.wn_tconv_scales):
            kk = ks[i] if isinstance(ks, (list, tuple)) else ks
            spect = conv_transpose1d(spect, np.asarray(sd[p + f"upsample_net.t_convs.{2 * i}.weight"], dtype),
                                     np.asarray(sd[p + f"upsample_net.t_convs.{2 * i}.bias"], dtype), sc, (kk - sc) // 2)
            if i < n - 1:                                                # use_last_layer_act_func=False (:294)
                spect = np.where(spect >= 0, spect, spect * dtype(0.4))
        if int(np.prod(cfg.wn_tconv_scales)) != cfg.hop_length // cfg.n_group and spect.shape[2] != T:   # interpolation_required
            spect = upsample_cond(spect, T, cfg.upsample_mode)
        else:                                                            # centre crop :367-372
            pad_l, pad_r = (spect.shape[2] - T) // 2, (-(T - spect.shape[2])) // 2
            if crop_2d:                                                  # WN_2d: cond[:, :, pad:-(pad + pad % 2)]
                pad_r = pad_l + pad_l % 2
            spect = spect[:, :, pad_l:spect.shape[2] - pad_r] if pad_r else spect[:, :, pad_l:0]
    elif not cfg.upsample_first:                                         # no WN upsample net: interpolation_required,
        spect = upsample_cond(spect, T, cfg.upsample_mode)               # F.interpolate to the audio length
    return spect


def wn_forward(sd, k, cfg: AxConfig, audio0, cond_up, dtype, speaker_ids=None):
    """glow_ax.WN.forward (:375-418): returns (log_s, t).  `cond_up` is at T' rate (upsample_first=True) or at frame
    rate (upsample_first=False: interpolated after the cond layer, :389)."""
    p = f"WN.{k}.WN."
    C, L = cfg.n_channels, cfg.n_layers
    audio = np.einsum("oc,bct->bot", _w(sd, p + "start", dtype)[:, :, 0], audio0, optimize=True) \
        + np.asarray(sd[p + "start.bias"], dtype)[None, :, None]
    B, _, T = audio.shape
    spect = wn_cond_path(sd, p, cfg, cond_up, T, dtype, speaker_ids)
    output = None
    unit_a, unit_b = GATED_UNITS[cfg.gated_unit.upper()]
    split = cfg.res_skip and not cfg.merge_res_skip
    for i in range(L):
        d = cfg.dilation(i)
        sep = (p + f"in_layers.{i}.0.weight_v") in sd
        w_in = _w(sd, p + (f"in_layers.{i}.0" if sep else f"in_layers.{i}"), dtype)
        ks = w_in.shape[2]
        pad = (ks * d - d) // 2
        xp = np.zeros((B, C, T + 2 * pad), dtype); xp[:, :, pad:pad + T] = audio
        if sep:                                                          # depthwise, then pointwise (no activation between)
            dw = np.zeros((B, C, T), dtype)
            for j in range(ks):
                dw += w_in[None, :, 0, j, None] * xp[:, :, j * d:j * d + T]
            dw += np.asarray(sd[p + f"in_layers.{i}.0.bias"], dtype)[None, :, None]
            acts = np.einsum("oc,bct->bot", _w(sd, p + f"in_layers.{i}.1", dtype)[:, :, 0], dw, optimize=True) \
                + np.asarray(sd[p + f"in_layers.{i}.1.bias"], dtype)[None, :, None]
        else:
            acts = np.zeros((B, 2 * C, T), dtype)
            for j in range(ks):
                acts += np.einsum("oc,bct->bot", w_in[:, :, j], xp[:, :, j * d:j * d + T], optimize=True)
            acts += np.asarray(sd[p + f"in_layers.{i}.bias"], dtype)[None, :, None]
        acts += spect[:, 2 * C * i:2 * C * (i + 1)]
        g = unit_a(acts[:, :C]) * unit_b(acts[:, C:])
        if cfg.res_skip:
            rs = np.einsum("oc,bct->bot", _w(sd, p + f"res_skip_layers.{i}", dtype)[:, :, 0], g, optimize=True) \
                + np.asarray(sd[p + f"res_skip_layers.{i}.bias"], dtype)[None, :, None]
        else:
            rs = g                                                       # :399
        if split and i < L - 1:                                          # :402-404, :409-411
            audio = audio + rs[:, :C]
            skip = rs[:, C:]
        else:                                                            # last layer, or merged: the hidden tensor stays (:406, :413)
            skip = rs
        output = skip if output is None else output + skip
    end = np.einsum("oc,bct->bot", np.asarray(sd[p + "end.weight"], dtype)[:, :, 0], output, optimize=True) \
        + np.asarray(sd[p + "end.bias"], dtype)[None, :, None]
    n = end.shape[1] // 2
    return end[:, :n], end[:, n:]                                        # chunk(2, 1): (log_s, t)


def mix_inverse(sd, k, cfg: AxConfig, z, dtype):
    if cfg.channel_mixing == "permuteheight":
        return permute_height(z, k)
    W = np.asarray(sd[f"convinv.{k}.weight"], dtype)[:, :, 0]
    W_inv = np.linalg.inv(W.astype(np.float64)).astype(dtype) if dtype == np.float64 else np.linalg.inv(W.astype(np.float32))
    return np.einsum("oc,bct->bot", W_inv, z, optimize=True)


def inverse(sd, cfg: AxConfig, z, cond, dtype=np.float32, cond_up=None, ignore_nan=True, speaker_ids=None):
    """`cond_up` (one [B, C, T'] array, or one per flow) replaces the plain interpolation when the model has a
    conditioning front-end (oracle/ax_frontend_oracle.py); with upsample_first=False it is at frame rate."""
    z = np.asarray(z, dtype)
    B = z.shape[0]
    zz = z.reshape(B, -1, cfg.n_group).transpose(0, 2, 1)               # :310
    if cond_up is None:
        cond_up = np.asarray(cond, dtype)
        if cfg.upsample_first:
            cond_up = upsample_cond(cond_up, zz.shape[2], cfg.upsample_mode)   # :313-314
    n_early = sum(1 for k in range(cfg.n_flows) if k % cfg.n_early_every == 0 and k > 0)
    sizes = [cfg.n_early_size] * n_early + [cfg.n_group - cfg.n_early_size * n_early]
    parts, off = [], 0
    for s in sizes:                                                      # :319-322
        parts.append(zz[:, off:off + s]); off += s
    *remained, zz = parts
    for k in reversed(range(cfg.n_flows)):                               # :325
        if not cfg.mix_first:
            zz = mix_inverse(sd, k, cfg, zz, dtype)
        n_half = zz.shape[1] // 2
        a0, a1 = zz[:, :n_half], zz[:, n_half:]
        k_cond = cond_up[k] if isinstance(cond_up, (list, tuple)) else cond_up      # :328
        log_s, t = wn_forward(sd, k, cfg, a0, k_cond, dtype, speaker_ids)   # efficient_modules.py:99-105
        with np.errstate(all="ignore"):
            zz = np.concatenate([a0, (a1 - t) / np.exp(log_s)], axis=1)
        if ignore_nan:                                                   # :331-332 (masked_fill_(isnan, 0))
            zz = np.where(np.isnan(zz), np.zeros_like(zz), zz)
        if cfg.mix_first:
            zz = mix_inverse(sd, k, cfg, zz, dtype)
        if k % cfg.n_early_every == 0 and k:
            zz = np.concatenate([remained.pop(), zz], axis=1)            # :339-340
    return np.ascontiguousarray(zz.transpose(0, 2, 1)).reshape(B, -1)


def infer_with_z(sd, cfg: AxConfig, spect, z, sigma, artifact_trimming=1, dtype=np.float32, speaker_ids=None):
    spect = np.asarray(spect, dtype)
    if artifact_trimming > 0:
        spect = np.concatenate([spect, np.zeros(spect.shape[:2] + (artifact_trimming,), dtype)], axis=2)
    samples = (spect.shape[2] - 1) * cfg.hop_length
    samples -= samples % cfg.n_group
    assert z.shape[1] == samples
    audio = inverse(sd, cfg, np.asarray(z, dtype) * dtype(sigma), spect, dtype, speaker_ids=speaker_ids)
    return audio[:, :-artifact_trimming * cfg.hop_length] if artifact_trimming > 0 else audio


def synthetic_state_dict(cfg: AxConfig, seed: int = 1234, cond_in_channels=None) -> Dict[str, np.ndarray]:
    """Reference ax key layout for waveflow=False (probe-printed): `convinv.{k}.weight` (1x1conv mixing
    only), `WN.{k}.WN.{in_layers,res_skip_layers,start,end,cond_layers.0}` with 3-D conv weights."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    C, L, ks = cfg.n_channels, cfg.n_layers, cfg.kernel_size

    def wn(prefix, shape, fan_in):
        bound = 1.0 / np.sqrt(fan_in)
        v = rs.uniform(-bound, bound, size=shape).astype(np.float32)
        norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
        sd[prefix + ".bias"] = rs.uniform(-bound, bound, size=(shape[0],)).astype(np.float32)
        sd[prefix + ".weight_g"] = (norm * rs.uniform(0.8, 1.2, size=norm.shape)).astype(np.float32)
        sd[prefix + ".weight_v"] = v

    for k, n_rem in enumerate(cfg.flow_channels()):
        n_half = n_rem // 2
        if cfg.channel_mixing == "1x1conv":
            q1, _ = np.linalg.qr(rs.standard_normal((n_rem, n_rem)))
            q2, _ = np.linalg.qr(rs.standard_normal((n_rem, n_rem)))
            sd[f"convinv.{k}.weight"] = (q1 @ np.diag(rs.uniform(0.7, 1.4, size=n_rem)) @ q2).astype(np.float32)[:, :, None]
        p = f"WN.{k}.WN."
        for i in range(L):
            if cfg.seperable_conv and ks > 1:
                wn(p + f"in_layers.{i}.0", (C, 1, ks), ks)
                wn(p + f"in_layers.{i}.1", (2 * C, C, 1), C)
            else:
                wn(p + f"in_layers.{i}", (2 * C, C, ks), C * ks)
            if cfg.res_skip:
                wn(p + f"res_skip_layers.{i}", (2 * C if (i < L - 1 and not cfg.merge_res_skip) else C, C, 1), C)
        wn(p + "start", (C, n_half, 1), n_half)
        sd[p + "end.weight"] = (rs.standard_normal((2 * n_half, C, 1)) * 0.02).astype(np.float32)
        sd[p + "end.bias"] = (rs.standard_normal((2 * n_half,)) * 0.02).astype(np.float32)
        cin = (cond_in_channels or cfg.n_mel_channels) + cfg.wn_speaker_embed_dim
        kc = 2 * cfg.wn_cond_kernel_size - 1
        cond_out = cfg.wn_tconv_hidden_dim if cfg.wn_tconv_scales else 2 * C * L        # glow_ax.py:302
        dims = [cin] + [cfg.wn_cond_hidden_channels] * (cfg.wn_cond_layers - 1) + [cond_out]
        for i, (di, do) in enumerate(zip(dims[:-1], dims[1:])):
            wn(p + f"cond_layers.{i}", (do, di, kc), di * kc)
        if cfg.wn_tconv_scales:                                          # TransposedUpsampleNet(hidden, 2CL, hidden, ...), :290-294
            n = len(cfg.wn_tconv_scales)
            for i, sc in enumerate(cfg.wn_tconv_scales):
                kk = cfg.wn_tconv_kernel_size[i] if isinstance(cfg.wn_tconv_kernel_size, (list, tuple)) else cfg.wn_tconv_kernel_size
                ci, co = cfg.wn_tconv_hidden_dim, (2 * C * L if i == n - 1 else cfg.wn_tconv_hidden_dim)
                bound = 1.0 / np.sqrt(ci * kk / sc)
                sd[p + f"upsample_net.t_convs.{2 * i}.weight"] = rs.uniform(-bound, bound, size=(ci, co, kk)).astype(np.float32)
                sd[p + f"upsample_net.t_convs.{2 * i}.bias"] = rs.uniform(-bound, bound, size=(co,)).astype(np.float32)
        if cfg.wn_speaker_embed_dim:
            sd[p + "speaker_embed.weight"] = rs.standard_normal((512, cfg.wn_speaker_embed_dim)).astype(np.float32)
    return sd
