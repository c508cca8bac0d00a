"""CPU oracle for the conditioning front-end and output filters of the reference's ax models.

TEST INFRASTRUCTURE ONLY (same rules as waveglow_oracle.py).  numpy restatement of, in
/root/reference/CookieTTS/_4_mtw/waveglow/:
  efficient_model_ax.py:62-134   constructor: speaker_embed, cond_layers, res_conv, alpha, upsample_net, n_flow_group_conv
  efficient_model_ax.py:280-317  inverse(): shift/scale, speaker concat, cond layers, residual, upsample, group conv
  efficient_model_ax.py:171-182  _upsample_mels (upsample_net, interpolate or centre crop)
  glow_ax.py:201-242             TransposedUpsampleNet
  efficient_model_ax.py:343-355  inverse volume map, inverse pre-emphasis (scipy.signal.lfilter)
Parity status: PINNED by oracle/make_golden_ax_frontend.py (the unmodified reference model with these options
switched on; tests/golden/axfe_*.npz); tests/test_ax_frontend_oracle.py holds this file to them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .waveflow_oracle import _w, upsample_cond


@dataclass
class FrontEndConfig:
    n_mel_channels: int = 80
    n_flows: int = 4
    n_group: int = 8
    hop_length: int = 16
    upsample_mode: str = "linear"
    upsample_first: bool = True            # False: the model-level upsampling is skipped (efficient_model_ax.py:313)
    speaker_embed: int = 0
    cond_layers: int = 0
    cond_hidden_channels: int = 16
    cond_output_channels: int = 16
    cond_kernel_size: int = 1
    cond_residual: object = False          # False / True / '1x1conv'
    cond_padding_mode: str = "zeros"
    cond_res_rezero: bool = False
    cond_activation_func: str = "none"
    negative_slope: Optional[float] = None
    transposed_conv_hidden_dim: int = 16
    transposed_conv_kernel_size: object = 4
    transposed_conv_scales: Optional[List[int]] = None
    transposed_conv_output_dim: Optional[int] = 16
    transposed_conv_residual: bool = False
    transposed_conv_residual_linear: bool = False
    transposed_conv_res_rezero: bool = False
    group_conv_output_dim: Optional[int] = None
    group_conv_groupped: bool = True
    shift_spect: float = 0.0
    scale_spect: float = 1.0
    preempthasis: Optional[float] = None
    preceived_vol_scaling: bool = False

    def in_channels(self) -> int:
        return self.n_mel_channels + self.speaker_embed

    def cond_out(self) -> int:
        if not self.cond_layers:
            return self.in_channels()
        return self.in_channels() if self.cond_residual is True else self.cond_output_channels

    def has_upsample_net(self) -> bool:
        return bool(self.transposed_conv_scales)

    def cond_channels(self) -> int:
        """Channels of the tensor handed to the flows (before n_flow_group_conv)."""
        ch = self.cond_out()
        if self.has_upsample_net():
            ch = self.transposed_conv_output_dim if self.transposed_conv_output_dim is not None else ch
        return ch

    def wn_cond_in(self) -> int:
        return self.group_conv_output_dim or self.cond_channels()


def _act(x, name, slope):
    name = name.lower()
    if name == "none":
        return x
    if name == "lrelu":                                    # efficient_model_ax.py:98-99 maps 'lrelu' to relu
        return np.maximum(x, 0)
    if name == "relu":                                     # :100-102 maps 'relu' to LeakyReLU(negative_slope)
        return np.where(x >= 0, x, x * slope)
    if name == "tanh":
        return np.tanh(x)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    raise NotImplementedError(name)


def conv1d(x, w, b, padding, mode):
    """nn.Conv1d(stride 1) with padding_mode zeros / replicate / reflect / circular."""
    B, C, T = x.shape
    np_mode = {"zeros": "constant", "replicate": "edge", "reflect": "reflect", "circular": "wrap"}[mode]
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)), mode=np_mode)
    k = w.shape[2]
    To = T + 2 * padding - (k - 1)
    y = np.zeros((B, w.shape[0], To), x.dtype)
    for j in range(k):
        y += np.einsum("oc,bct->bot", w[:, :, j], xp[:, :, j:j + To], optimize=True)
    return y + b[None, :, None]


def conv_transpose1d(x, w, b, stride, padding):
    """nn.ConvTranspose1d: w [c_in, c_out, k]."""
    B, C, T = x.shape
    k = w.shape[2]
    full = np.zeros((B, w.shape[1], (T - 1) * stride + k), x.dtype)
    for j in range(k):
        full[:, :, j:j + (T - 1) * stride + 1:stride] += np.einsum("co,bct->bot", w[:, :, j], x, optimize=True)
    out = full[:, :, padding:full.shape[2] - padding] if padding else full
    return out + b[None, :, None]


def interp_linear_half_pixel(x, scale: int):
    """F.interpolate(x, scale_factor=scale, mode='linear', align_corners=False) (glow_ax.py:231)."""
    Tin = x.shape[2]
    Tout = Tin * scale
    src = np.maximum((np.arange(Tout, dtype=np.float64) + 0.5) / scale - 0.5, 0.0)
    i0 = np.minimum(np.floor(src).astype(np.int64), Tin - 1)
    i1 = np.minimum(i0 + 1, Tin - 1)
    l1 = (src - i0).astype(x.dtype)
    return x[:, :, i0] * (1 - l1) + x[:, :, i1] * l1


def frontend(sd, fe: FrontEndConfig, spect, speaker_ids, n_steps: int, dtype=np.float64):
    """efficient_model_ax.py:280-317: spect [B, n_mel, frames] -> cond at T' rate ([B, C, T'] or one per flow)."""
    cond = np.asarray(spect, dtype)
    if fe.shift_spect != 0.0:
        cond = cond + dtype(fe.shift_spect)
    if fe.scale_spect != 1.0:
        cond = cond * dtype(fe.scale_spect)
    if fe.speaker_embed:
        emb = np.asarray(sd["speaker_embed.weight"], dtype)[np.asarray(speaker_ids)]
        cond = np.concatenate([cond, np.repeat(emb[:, :, None], cond.shape[2], axis=2)], axis=1)
    cond_res = cond
    pad = (2 * fe.cond_kernel_size - 1 - 1) // 2
    for i in range(fe.cond_layers):
        cond_res = conv1d(cond_res, _w(sd, f"cond_layers.{i}", dtype), np.asarray(sd[f"cond_layers.{i}.bias"], dtype), pad, fe.cond_padding_mode)
        cond_res = _act(cond_res, fe.cond_activation_func, fe.negative_slope)
    if fe.cond_res_rezero:
        cond_res = cond_res * np.asarray(sd["alpha"], dtype)
    if fe.cond_residual:
        if "res_conv.weight" in sd:
            cond = conv1d(cond, np.asarray(sd["res_conv.weight"], dtype), np.asarray(sd["res_conv.bias"], dtype), 0, "zeros")
        cond = cond + cond_res
    else:
        cond = cond_res
    interpolation_required = True
    if not fe.upsample_first:                              # efficient_model_ax.py:313: `if self.upsample_early is True`
        pass
    elif fe.has_upsample_net():                            # glow_ax.py:228-242
        scales = fe.transposed_conv_scales
        x = cond
        ks = fe.transposed_conv_kernel_size
        idx = 0
        for i, s in enumerate(scales):
            k = ks[i] if isinstance(ks, (list, tuple)) else ks
            x = conv_transpose1d(x, np.asarray(sd[f"upsample_net.t_convs.{idx}.weight"], dtype),
                                 np.asarray(sd[f"upsample_net.t_convs.{idx}.bias"], dtype), s, (k - s) // 2)
            x = np.where(x >= 0, x, x * dtype(0.4))        # LeakyReLU(0.4) after every layer (use_last_layer_act_func=True)
            idx += 2
        if fe.transposed_conv_residual:
            if fe.transposed_conv_res_rezero and float(np.asarray(sd["upsample_net.res_weight"]).reshape(-1)[0]) != 0.0:
                x = x * np.asarray(sd["upsample_net.res_weight"], dtype)
            rc = min(cond.shape[1], x.shape[1])
            x[:, :rc] += interp_linear_half_pixel(cond, int(np.prod(scales)))[:, :rc]
        cond = x
        interpolation_required = int(np.prod(scales)) != fe.hop_length // fe.n_group
    if not fe.upsample_first:
        pass
    elif interpolation_required and cond.shape[2] != n_steps:   # efficient_model_ax.py:174-175
        cond = upsample_cond(cond, n_steps, fe.upsample_mode)
    else:                                                  # :176-181
        pad_l = (cond.shape[2] - n_steps) // 2
        pad_r = (-(n_steps - cond.shape[2])) // 2
        cond = cond[:, :, pad_l:cond.shape[2] - pad_r] if pad_r else cond[:, :, pad_l:0]
    if fe.group_conv_output_dim:                           # :316-317
        wg = np.asarray(sd["n_flow_group_conv.weight"], dtype)[:, :, 0]
        bg = np.asarray(sd["n_flow_group_conv.bias"], dtype)
        g = fe.group_conv_output_dim
        outs = []
        per = cond.shape[1] // fe.n_flows
        for k in range(fe.n_flows):
            xin = cond[:, k * per:(k + 1) * per] if fe.group_conv_groupped else cond
            outs.append(np.einsum("oc,bct->bot", wg[k * g:(k + 1) * g], xin, optimize=True) + bg[k * g:(k + 1) * g][None, :, None])
        return outs
    return cond


def post_filter(fe: FrontEndConfig, audio, dtype=np.float64):
    """efficient_model_ax.py:343-355: z[z>0] = 10**log2(z) (and the mirrored negative branch), then
    scipy.signal.lfilter([1], [1, -coef]) per utterance (fp64) cast back to the model dtype."""
    audio = np.array(audio, dtype)
    if fe.preceived_vol_scaling:
        pos, neg = audio > 0, audio < 0
        audio[pos] = 10.0 ** np.log2(audio[pos])
        audio[neg] = -(10.0 ** np.log2(-audio[neg]))
    if fe.preempthasis:
        from scipy import signal
        audio = np.stack([signal.lfilter([1], [1, -float(fe.preempthasis)], row) for row in audio]).astype(dtype)
    return audio


def synthetic_frontend_state_dict(fe: FrontEndConfig, seed: int) -> Dict[str, np.ndarray]:
    """Seeded front-end parameters with the reference's key names (pinned by strict load in make_golden_ax_frontend.py)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}

    def conv(prefix, shape, fan_in, weight_norm):
        bound = 1.0 / np.sqrt(fan_in)
        v = rs.uniform(-bound, bound, size=shape).astype(np.float32)
        sd[prefix + ".bias"] = rs.uniform(-bound, bound, size=(shape[0] if weight_norm is not None else shape[0],)).astype(np.float32)
        if weight_norm:
            norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
            sd[prefix + ".weight_g"] = (norm * rs.uniform(0.8, 1.2, size=norm.shape)).astype(np.float32)
            sd[prefix + ".weight_v"] = v
        else:
            sd[prefix + ".weight"] = v

    if fe.speaker_embed:
        sd["speaker_embed.weight"] = rs.standard_normal((512, fe.speaker_embed)).astype(np.float32)
    ch = fe.in_channels()
    if fe.cond_res_rezero:
        sd["alpha"] = np.array([0.7], np.float32)
    if fe.cond_layers:
        ks = 2 * fe.cond_kernel_size - 1
        if fe.cond_residual == "1x1conv":
            conv("res_conv", (fe.cond_out(), ch, 1), ch, False)
        dims = [ch] + [fe.cond_hidden_channels] * (fe.cond_layers - 1) + [fe.cond_out()]
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            conv(f"cond_layers.{i}", (cout, cin, ks), cin * ks, True)
        ch = fe.cond_out()
    if fe.has_upsample_net():
        ks = fe.transposed_conv_kernel_size
        n = len(fe.transposed_conv_scales)
        t_out = fe.cond_channels()
        for i in range(n):
            k = ks[i] if isinstance(ks, (list, tuple)) else ks
            cin = ch if i == 0 else fe.transposed_conv_hidden_dim
            cout = t_out if i == n - 1 else fe.transposed_conv_hidden_dim
            bound = 1.0 / np.sqrt(cin * k / fe.transposed_conv_scales[i])
            sd[f"upsample_net.t_convs.{2 * i}.weight"] = rs.uniform(-bound, bound, size=(cin, cout, k)).astype(np.float32)
            sd[f"upsample_net.t_convs.{2 * i}.bias"] = rs.uniform(-bound, bound, size=(cout,)).astype(np.float32)
        if fe.transposed_conv_res_rezero:
            sd["upsample_net.res_weight"] = np.array([0.6], np.float32)
        ch = t_out
    if fe.group_conv_output_dim:
        groups = fe.n_flows if fe.group_conv_groupped else 1
        conv("n_flow_group_conv", (fe.group_conv_output_dim * fe.n_flows, ch // groups, 1), ch // groups, False)
    return sd
