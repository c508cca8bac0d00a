"""Conditioning front-end and output filters of the reference's "ax" models, shared by
`WaveGlowAx` and `WaveFlow`.

What `efficient_model_ax.py::WaveGlow` does to the mel before the flows (:62-134 construction, :280-317 in
`inverse`) and to the waveform after them (:343-355):
  * model-level speaker embedding, concatenated to the mel as constant channels (:64-66,287-292)
  * model-level `cond_layers` (weight-normed Conv1d chain, optional activation after every layer, ReZero `alpha`,
    residual with an optional 1x1 `res_conv`; :74-113,294-307)
  * `TransposedUpsampleNet` (ConvTranspose1d + LeakyReLU(0.4) chain, optional linear-interpolated residual;
    glow_ax.py:201-242), then `F.interpolate` or the centre crop of `_upsample_mels` (:171-182)
  * `n_flow_group_conv` (:128-131,316-317): a 1x1 conv in front of each flow's own 1x1 cond layer - folded into that
    layer's weights at pack time (both are linear), see `group_conv_fold`
  * inverse perceived-volume map and inverse pre-emphasis (the reference runs scipy.signal.lfilter on the host)
Parameter names match the reference's state_dict.  All arithmetic runs in the fp32 kernels of csrc/cwg_condnet.cu
through the C ABI; torch only allocates, gathers embedding rows and concatenates.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _cabi
from .packing import effective_weight, _np

PAD_MODES = {"zeros": 0, "replicate": 1, "reflect": 2, "circular": 3}
# efficient_model_ax.py:98-109 - the reference maps 'lrelu' to relu and 'relu' to LeakyReLU(negative_slope)
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4


def _cond_act(name: str, negative_slope):
    name = str(name).lower()
    if name == "none":
        return ACT_NONE, 0.0
    if name == "lrelu":
        return ACT_RELU, 0.0
    if name == "relu":
        assert negative_slope, "negative_slope not defined in wn_config"
        return ACT_LRELU, float(negative_slope)
    if name == "tanh":
        return ACT_TANH, 0.0
    if name == "sigmoid":
        return ACT_SIGMOID, 0.0
    raise NotImplementedError(name)


class TransposedUpsampleNet(nn.Module):
    """Parameter holder with the layout of glow_ax.py:201-227 (`t_convs.{i}` interleaved with LeakyReLU slots)."""

    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size=3, scales=(16, 16),
                 use_last_layer_act_func=False, residual=False, residual_linear=False, rezero=False):
        super().__init__()
        self.residual, self.residual_linear = residual, residual_linear
        self.res_weight = nn.Parameter(torch.rand(1) * 0.02 + 0.01) if rezero else None
        self.scales = list(scales)
        self.t_convs = nn.ModuleList()
        self.layers = []                                  # (index in t_convs, k, stride, padding, has_act)
        for i, scale in enumerate(self.scales):
            last = i + 1 == len(self.scales)
            k = kernel_size[i] if isinstance(kernel_size, (list, tuple)) else kernel_size
            self.t_convs.append(nn.ConvTranspose1d(in_channels if i == 0 else hidden_channels,
                                                   out_channels if last else hidden_channels, k, scale,
                                                   padding=(k - scale) // 2))
            act = (not last) or use_last_layer_act_func
            self.layers.append((len(self.t_convs) - 1, k, scale, (k - scale) // 2, act))
            if act:
                self.t_convs.append(nn.LeakyReLU(negative_slope=0.4, inplace=True))
        self.res_channels = min(in_channels, out_channels)


def repack_conv_transpose(w: np.ndarray, stride: int) -> np.ndarray:
    """torch ConvTranspose1d weight [c_in, c_out, k] -> [stride, c_out, c_in, ceil(k/stride)] (include/cwg.h)."""
    c_in, c_out, k = w.shape
    jmax = (k + stride - 1) // stride
    out = np.zeros((stride, c_out, c_in, jmax), np.float32)
    for r in range(stride):
        for jj in range(jmax):
            if r + stride * jj < k:
                out[r, :, :, jj] = w[:, :, r + stride * jj].T
    return out


class AxFrontEndMixin:
    """Mixed into the ax drop-in modules; the submodules it creates are attributes of the model itself so that the
    state_dict keys are the reference's (`speaker_embed.weight`, `cond_layers.{i}.*`, `res_conv.*`, `alpha`,
    `upsample_net.t_convs.{i}.*`, `upsample_net.res_weight`, `n_flow_group_conv.*`)."""

    def _fe_build(self, a: dict, wn: dict) -> int:
        """Creates the front-end parameters from the constructor arguments `a`; returns the channel count of the cond
        tensor handed to the flows (before `n_flow_group_conv`, which is folded into the per-flow cond layer)."""
        def need(cond, msg):
            if not cond:
                raise NotImplementedError(f"cookietts_b200.{type(self).__name__}: {msg}")
        need(not a["iso226_empthasis"], "ISO-226 emphasis needs a module the reference downloads at import (iso226.py:3-9)")
        need(not a["load_hidden_from_disk"] and not a["spect_scaling"] and not a["memory_efficient"], "unsupported input/training options")
        self.speaker_embed_dim = int(a["speaker_embed"] or 0)
        self.has_logvar_channels = bool(a["use_logvar_channels"])
        self.vol_scaling = bool(a["preceived_vol_scaling"])
        self.preempthasis = a["preempthasis"]
        n_mel_in = a["n_mel_channels"] * (2 if self.has_logvar_channels else 1)
        if self.speaker_embed_dim:
            self.speaker_embed = nn.Embedding(512, self.speaker_embed_dim)
        ch = n_mel_in + self.speaker_embed_dim
        self._fe_in_channels = ch
        # ---- cond layers (efficient_model_ax.py:71-113)
        self.cond_residual = a["cond_residual"]
        cond_out = a["cond_output_channels"]
        if self.cond_residual is True or (type(self.cond_residual) is int and self.cond_residual == 1):   # :72-73
            cond_out = ch
        self.cond_res_rezero = bool(a["cond_res_rezero"])
        if self.cond_res_rezero:
            self.alpha = nn.Parameter(torch.rand(1) * 0.02 + 0.01)
        self.cond_layers = nn.ModuleList()
        self._fe_cond_act = (ACT_NONE, 0.0)
        need(a["cond_layers"] or not (self.cond_residual or self.cond_res_rezero),
             "cond_residual / cond_res_rezero without cond_layers (the reference aliases cond with itself there)")
        if a["cond_layers"]:
            if self.cond_residual == "1x1conv":
                self.res_conv = nn.Conv1d(ch, cond_out, 1)
            ks = 2 * a["cond_kernel_size"] - 1
            need(a["cond_padding_mode"] in PAD_MODES, "cond_padding_mode must be zeros / replicate / reflect / circular")
            self._fe_cond_pad = ((ks - 1) // 2, PAD_MODES[a["cond_padding_mode"]])
            dims = [ch] + [a["cond_hidden_channels"]] * (a["cond_layers"] - 1) + [cond_out]
            for cin, cout in zip(dims[:-1], dims[1:]):
                self.cond_layers.append(nn.utils.weight_norm(
                    nn.Conv1d(cin, cout, ks, padding=(ks - 1) // 2, padding_mode=a["cond_padding_mode"]), name="weight"))
            ch = cond_out
            self._fe_cond_act = _cond_act(a["cond_activation_func"], a["negative_slope"])
        # ---- TransposedUpsampleNet (efficient_model_ax.py:115-126)
        self.upsample_factor = a["hop_length"] // a["n_group"]
        self.interpolation_required = True
        scales = a["transposed_conv_scales"]
        if scales is not None and len(scales) > 0 and a["transposed_conv_hidden_dim"] and a["transposed_conv_kernel_size"]:
            t_out = a["transposed_conv_output_dim"] if a["transposed_conv_output_dim"] is not None else ch
            need(not a["transposed_conv_residual"] or a["transposed_conv_residual_linear"],
                 "transposed_conv_residual needs residual_linear=True (the reference's nearest branch raises in F.interpolate)")
            self.upsample_net = TransposedUpsampleNet(ch, t_out, a["transposed_conv_hidden_dim"], a["transposed_conv_kernel_size"],
                                                      scales, use_last_layer_act_func=True, residual=a["transposed_conv_residual"],
                                                      residual_linear=a["transposed_conv_residual_linear"],
                                                      rezero=a["transposed_conv_res_rezero"])
            self.interpolation_required = bool(int(np.prod(scales)) != self.upsample_factor)
            ch = t_out
        need(not wn.get("transposed_conv_scales") or getattr(self, "general", False),
             "a WN-level TransposedUpsampleNet is served by the general fp32 modes only")
        # ---- n_flow_group_conv (efficient_model_ax.py:128-131)
        self._fe_group = None
        if a["group_conv_output_dim"]:
            g = a["n_flows"] if a["group_conv_groupped"] else 1
            self.n_flow_group_conv = nn.Conv1d(ch, a["group_conv_output_dim"] * a["n_flows"], 1, groups=g)
            self._fe_group = (int(a["group_conv_output_dim"]), bool(a["group_conv_groupped"]))
        self._fe_cond_channels = ch
        self._fe_packed = None
        self._fe_key = None
        return ch

    @property
    def wn_cond_in_channels(self) -> int:
        """`cond_in_channels` of every flow's WN (efficient_model_ax.py:166-167)."""
        return self._fe_group[0] if self._fe_group else self._fe_cond_channels

    def group_conv_fold(self, k: int, w_c: np.ndarray, b_c: np.ndarray, sd):
        """Folds flow k's slice of `n_flow_group_conv` into its WN cond layer: returns (W [2CL, cond_channels], b)."""
        if not self._fe_group:
            return w_c, b_c
        gdim, grouped = self._fe_group
        wg = _np(sd["n_flow_group_conv.weight"])[:, :, 0]            # [gdim*F, ch/groups]
        bg = _np(sd["n_flow_group_conv.bias"])
        ch = self._fe_cond_channels
        full = np.zeros((gdim, ch))
        rows = wg[k * gdim:(k + 1) * gdim]
        if grouped:
            per = ch // self.n_flows
            full[:, k * per:(k + 1) * per] = rows
        else:
            full[:] = rows
        return w_c @ full, b_c + w_c @ bg[k * gdim:(k + 1) * gdim]

    # ------------------------------------------------------------------ device weights
    def _fe_params(self):
        mods = [getattr(self, n, None) for n in ("speaker_embed", "cond_layers", "res_conv", "upsample_net")]
        ps = [p for m in mods if m is not None for p in m.parameters()]
        if hasattr(self, "alpha"):
            ps.append(self.alpha)
        return ps

    def _fe_ensure_packed(self, dev):
        key = tuple((p.data_ptr(), p._version) for p in self._fe_params())
        if self._fe_packed is not None and self._fe_key == key:
            return self._fe_packed
        sd = {k: v.detach().float().cpu().numpy() for k, v in self.state_dict().items() if not k.startswith(("WN.", "convinv."))}
        up = lambda arr: torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(dev)
        pk = {"cond": [], "tconv": []}
        for i in range(len(self.cond_layers)):
            pk["cond"].append((up(effective_weight(sd, f"cond_layers.{i}")), up(sd[f"cond_layers.{i}.bias"])))
        if hasattr(self, "res_conv"):
            pk["res_conv"] = (up(sd["res_conv.weight"]), up(sd["res_conv.bias"]))
        if hasattr(self, "upsample_net"):
            for (idx, k, s, p, act) in self.upsample_net.layers:
                w = sd[f"upsample_net.t_convs.{idx}.weight"]
                pk["tconv"].append((up(repack_conv_transpose(w, s)), up(sd[f"upsample_net.t_convs.{idx}.bias"]),
                                    w.shape[0], w.shape[1], k, s, p, act))
        # host copies of the scalar gates, so that applying the front-end never synchronises (CUDA-graph capturable)
        pk["alpha"] = float(np.asarray(sd["alpha"]).reshape(-1)[0]) if hasattr(self, "alpha") else 1.0
        if hasattr(self, "upsample_net") and self.upsample_net.res_weight is not None:
            pk["res_weight"] = float(np.asarray(sd["upsample_net.res_weight"]).reshape(-1)[0])
        self._fe_packed, self._fe_key = pk, key
        return pk

    # ------------------------------------------------------------------ kernels
    @staticmethod
    def _conv1d(lib, x, w, b, padding, pad_mode, act, slope, out_scale=1.0, res=None):
        B, cin, T = x.shape
        cout, _, k = w.shape
        y = torch.empty(B, cout, T + 2 * padding - (k - 1), device=x.device, dtype=torch.float32)
        _cabi.check(lib.cwg_conv1d(x.data_ptr(), B, cin, T, w.data_ptr(), b.data_ptr(), cout, k, padding, pad_mode, act,
                                   float(slope), float(out_scale), res.data_ptr() if res is not None else None,
                                   y.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
        return y

    @staticmethod
    def _tconv_chain(lib, h, tconvs, stream, last_scale=1.0):
        """TransposedUpsampleNet.forward without the residual (glow_ax.py:233-236): ConvTranspose1d (+ LeakyReLU(0.4)) per
        entry of `tconvs` = (w_phases, bias, c_in, c_out, k, stride, padding, has_act); the last output is scaled by
        `last_scale` (ReZero res_weight)."""
        n = len(tconvs)
        for i, (w, b, cin, cout, k, s, p, has_act) in enumerate(tconvs):
            Bh, _, T = h.shape
            y = torch.empty(Bh, cout, (T - 1) * s - 2 * p + k, device=h.device, dtype=torch.float32)
            _cabi.check(lib.cwg_conv_transpose1d(h.data_ptr(), Bh, cin, T, w.data_ptr(), b.data_ptr(), cout, k, s, p,
                                                 ACT_LRELU if has_act else ACT_NONE, 0.4, last_scale if i == n - 1 else 1.0,
                                                 y.data_ptr(), stream))
            h = y
        return h

    def _wn_resample(self, lib, x, t_out, stream):
        """F.interpolate(x, size=t_out, mode, align_corners=True for 'linear') - efficient_model_ax.py:175, glow_ax.py:365"""
        B, ch, t_in = x.shape
        y = torch.empty(B, ch, t_out, device=x.device, dtype=torch.float32)
        _cabi.check(lib.cwg_resample1d(x.data_ptr(), B, ch, t_in, ch * t_in, y.data_ptr(), t_out, ch * t_out,
                                       1 if self.upsample_linear else 0, t_out, 0, 0.0, 0, stream))
        return y

    def _wn_cond_apply(self, lib, cond, f, ids, Tp, stream, flow, crop_2d=False):
        """The cond path of ONE flow's WN in its general form (glow_ax.py:378-389 / WN_2d :565-579): flow slice of
        n_flow_group_conv, WN-level speaker embedding, the cond stack (self._wn_cond: depth, kernel, padding mode, activation,
        out_act), then - with upsample_first=False - the WN's own upsample net and crop / interpolation (_upsample_mels :361-373;
        WN_2d crops with its own rule :545-553).  `f` holds this flow's device weights: group, emb, cond [(w, b)], tconv.
        Returns [B, 2*C*L, Tp]."""
        wc = self._wn_cond
        pad, pad_mode = (2 * wc["kernel_size"] - 2) // 2, PAD_MODES[wc["padding_mode"]]
        act, slope = wc["act"]
        x = cond
        B = x.shape[0]
        if f.get("group") is not None:                       # efficient_model_ax.py:316-317,:328
            x = self._conv1d(lib, x, *f["group"], 0, 0, ACT_NONE, 0.0)
        if f.get("emb") is not None:                         # glow_ax.py:378-381
            x = torch.cat([x, f["emb"][ids][:, :, None].expand(-1, -1, x.shape[2])], dim=1).contiguous()
        n = len(f["cond"])
        for i, (w, b) in enumerate(f["cond"]):               # :383-387
            a = act if (wc["out_act"] or i != n - 1) else ACT_NONE
            x = self._conv1d(lib, x, w, b, pad, pad_mode, a, slope)
        if not self.upsample_first and f.get("tconv"):       # :389, _upsample_mels with the WN's upsample net
            x = self._tconv_chain(lib, x, f["tconv"], stream)
            if int(np.prod(self._wn_tconv["scales"])) != self.hop_length // self.n_group and x.shape[2] != Tp:   # interpolation_required
                x = self._wn_resample(lib, x, Tp, stream)
            else:                                            # centre crop
                diff = x.shape[2] - Tp
                pad_l = diff // 2
                pad_r = pad_l + (pad_l % 2) if crop_2d else pad_l     # WN_2d: cond[:, :, pad:-(pad + pad % 2)], glow_ax.py:550-552
                if diff <= 0 or pad_r == 0 or x.shape[2] - pad_l - pad_r != Tp:
                    raise RuntimeError(f"WN {flow}: upsampled cond length {x.shape[2]} cannot be cropped to {Tp} group-steps "
                                       "(the reference's slice is empty or mis-sized here too)")
                y = torch.empty(B, x.shape[1], Tp, device=x.device, dtype=torch.float32)
                _cabi.check(lib.cwg_resample1d(x.data_ptr(), B, x.shape[1], x.shape[2], x.shape[1] * x.shape[2], y.data_ptr(),
                                               Tp, x.shape[1] * Tp, 0, x.shape[2], pad_l, 0.0, 0, stream))
                x = y
        elif not self.upsample_first and x.shape[2] != Tp:   # ... without one: interpolation_required, F.interpolate
            x = self._wn_resample(lib, x, Tp, stream)
        return x

    @torch.no_grad()
    def _fe_apply(self, spect: torch.Tensor, speaker_ids, n_steps: int) -> torch.Tensor:
        """spect [B, n_mel, frames] fp32 on the device (already shifted/scaled and zero-padded by `infer`) -> the cond
        tensor [B, cond_channels, frames'] the flows interpolate (frames' != n_steps) or use as is (frames' == n_steps).
        efficient_model_ax.py:286-314."""
        dev = spect.device
        lib = _cabi.load()
        stream = torch.cuda.current_stream(dev).cuda_stream
        cond = spect.contiguous()
        B = cond.shape[0]
        if self.speaker_embed_dim:
            if speaker_ids is None:
                raise Exception("This WaveFlow/WaveGlow model requires speaker ids or speaker embeddings.")
            emb = self.speaker_embed.weight.detach().float()[torch.as_tensor(speaker_ids, device=dev).long().view(-1)]
            cond = torch.cat([cond, emb[:, :, None].expand(-1, -1, cond.shape[2])], dim=1).contiguous()
        if cond.shape[1] != self._fe_in_channels:
            raise ValueError(f"spect must have {self._fe_in_channels - self.speaker_embed_dim} channels, got {spect.shape[1]}")
        pk = self._fe_ensure_packed(dev)
        if len(self.cond_layers):
            pad, pad_mode = self._fe_cond_pad
            act, slope = self._fe_cond_act
            res = None
            if self.cond_residual:
                res = self._conv1d(lib, cond, *pk["res_conv"], 0, 0, ACT_NONE, 0.0) if "res_conv" in pk else cond
            alpha = pk["alpha"]
            h = cond
            n = len(pk["cond"])
            for i, (w, b) in enumerate(pk["cond"]):
                last = i == n - 1
                h = self._conv1d(lib, h, w, b, pad, pad_mode, act, slope, alpha if last else 1.0, res if last else None)
            cond = h
        if hasattr(self, "upsample_net"):
            net = self.upsample_net
            x_in = cond
            h = cond
            n = len(pk["tconv"])
            last_scale = pk["res_weight"] if (net.residual and pk.get("res_weight", 0.0) != 0.0) else 1.0
            h = self._tconv_chain(lib, h, pk["tconv"], stream, last_scale)
            if net.residual:                                   # glow_ax.py:229-241
                sf = int(np.prod(net.scales))
                t_virtual = x_in.shape[2] * sf
                if t_virtual != h.shape[2]:
                    raise RuntimeError("TransposedUpsampleNet residual: interpolated and transposed-conv lengths differ "
                                       "(the reference fails here too)")
                rc = net.res_channels
                _cabi.check(lib.cwg_resample1d(x_in.data_ptr(), B, rc, x_in.shape[2], x_in.shape[1] * x_in.shape[2],
                                               h.data_ptr(), h.shape[2], h.shape[1] * h.shape[2], 2, t_virtual, 0,
                                               float(sf), 1, stream))
            cond = h
            if not (self.interpolation_required and cond.shape[2] != n_steps):   # centre crop, efficient_model_ax.py:176-181
                diff = cond.shape[2] - n_steps
                pad_l, pad_r = diff // 2, diff // 2
                if diff <= 0 or pad_r == 0 or cond.shape[2] - pad_l - pad_r != n_steps:
                    raise RuntimeError(f"upsampled cond length {cond.shape[2]} cannot be cropped to {n_steps} group-steps "
                                       "(the reference's slice is empty or mis-sized here too)")
                y = torch.empty(B, cond.shape[1], n_steps, device=dev, dtype=torch.float32)
                _cabi.check(lib.cwg_resample1d(cond.data_ptr(), B, cond.shape[1], cond.shape[2], cond.shape[1] * cond.shape[2],
                                               y.data_ptr(), n_steps, cond.shape[1] * n_steps, 0, cond.shape[2], pad_l,
                                               0.0, 0, stream))
                cond = y
        return cond

    @torch.no_grad()
    def _fe_post(self, audio: torch.Tensor) -> torch.Tensor:
        """efficient_model_ax.py:343-355 on the device: inverse volume map, then y[n] = x[n] + coef*y[n-1]."""
        if not self.vol_scaling and not self.preempthasis:
            return audio
        lib = _cabi.load()
        out = torch.empty_like(audio)
        _cabi.check(lib.cwg_deemphasis(audio.data_ptr(), audio.shape[0], audio.shape[1], float(self.preempthasis or 0.0),
                                       int(self.vol_scaling), out.data_ptr(), torch.cuda.current_stream(audio.device).cuda_stream))
        return out
