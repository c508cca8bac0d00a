"""ORACLE (test infrastructure, not the product): numpy restatement of the mel-domain flow decoders' inverse pass
 with the latent passed in.

Follows, line by line,
  CookieTTS/_2_ttm/flowtts/waveglow/glow.py:133-172   WN.forward (start, WN-level cond layers, GTU, res/skip or merged, end)
  CookieTTS/_2_ttm/flowtts/waveglow/glow.py:302-343   FlowDecoder.inverse (split z, flows in reverse, mix_first, early z)
  CookieTTS/_2_ttm/flowtts/waveglow/modules.py:36-47  AffineCouplingBlock.inverse  (log_s, t = WN(...); (x1 - t) / exp(log_s))
  CookieTTS/_2_ttm/flowtts/waveglow/modules.py:234-250 InvertibleConv1x1.inverse   (conv1d with W^-1)
  CookieTTS/_2_ttm/untts/waveglow/glow.py:126-127      the untts variant pads the hidden tensor of the FIRST flow's WN with
                                                       `decoder_padding_value` instead of zeros
Pinned by tests/golden/fd_*.npz, which oracle/make_golden_flow_decoder.py generates from the unmodified reference classes.
Only tests/ may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass
class FlowDecoderConfig:
    n_mel_channels: int = 160
    n_flows: int = 10
    n_group: int = 160
    n_early_every: int = 4
    n_early_size: int = 20
    mix_first: bool = True
    cond_channels: int = 768          # encoder_LSTM_dim + speaker_embedding_dim (flowtts) / cond_input_dim (untts)
    wn_n_channels: int = 256
    wn_kernel_size: int = 3
    wn_n_layers: int = 1
    wn_dilations_w: Optional[List[int]] = None     # None: 2^i; an int in the reference means a constant dilation
    wn_res_skip: bool = False
    wn_merge_res_skip: bool = True
    first_pad_value: float = 0.0      # untts decoder_padding_value (hidden-tensor padding of flow 0's WN)
    end_std: float = 0.05             # synthetic checkpoints only: std of the (non-zero) `end` weights
    # WN variants (glow.py:74-101,:109-126,:145-148): separable in_layers, cond stacks with an activation after EVERY layer
    wn_seperable_conv: bool = False
    wn_cond_layers: int = 1
    wn_cond_hidden_channels: int = 256
    wn_cond_kernel_size: int = 1      # taps of the cond convs (odd; 'same' padding)
    wn_cond_padding_mode: str = "zeros"
    wn_cond_act_func: str = "none"    # 'lrelu' (= relu in the reference), 'tanh', 'sigmoid'; 'relu' raises NameError upstream

    def dilations(self) -> List[int]:
        if self.wn_dilations_w is None:
            return [2 ** i for i in range(self.wn_n_layers)]
        if isinstance(self.wn_dilations_w, int):
            return [self.wn_dilations_w] * self.wn_n_layers
        return list(self.wn_dilations_w)

    def flow_channels(self) -> List[int]:
        """n_remaining_channels of flow k (glow.py:205-220)."""
        out, n_rem = [], self.n_group
        for k in range(self.n_flows):
            if k % self.n_early_every == 0 and k > 0:
                n_rem -= self.n_early_size
            out.append(n_rem)
        return out

    def rs_channels(self, i: int) -> int:
        return 2 * self.wn_n_channels if (i < self.wn_n_layers - 1 and not self.wn_merge_res_skip) else self.wn_n_channels


def _eff(sd, prefix):
    if prefix + ".weight_g" in sd:
        g, v = sd[prefix + ".weight_g"].astype(np.float64), sd[prefix + ".weight_v"].astype(np.float64)
        return g * v / np.sqrt((v ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return sd[prefix + ".weight"].astype(np.float64)


def _conv1d(x, w, b, dilation=1, pad_value=0.0):
    """x [B, Ci, T], w [Co, Ci, k] ('same' padding k*d - d over 2 per side, constant pad value), b [Co]"""
    B, Ci, T = x.shape
    Co, _, k = w.shape
    p = (k * dilation - dilation) // 2
    xp = np.full((B, Ci, T + 2 * p), pad_value, dtype=x.dtype)
    xp[:, :, p:p + T] = x
    y = np.zeros((B, Co, T), dtype=x.dtype)
    for j in range(k):
        y += np.einsum("oc,bct->bot", w[:, :, j], xp[:, :, j * dilation:j * dilation + T])
    return y + b[None, :, None]


def wn_forward(sd, k, cfg: FlowDecoderConfig, z0, cond):
    """glow.py:133-172.  Returns (log_s, t)."""
    p = f"WN.{k}.WN."
    C = cfg.wn_n_channels
    h = _conv1d(z0, _eff(sd, p + "start"), sd[p + "start.bias"].astype(np.float64))
    c = cond
    for j in range(cfg.wn_cond_layers):                                  # glow.py:145-148
        w_c, b_c = _eff(sd, p + f"cond_layers.{j}"), sd[p + f"cond_layers.{j}.bias"].astype(np.float64)
        kc = w_c.shape[2]
        mode = {"zeros": "constant", "replicate": "edge", "reflect": "reflect", "circular": "wrap"}[cfg.wn_cond_padding_mode]
        cp = np.pad(c, ((0, 0), (0, 0), ((kc - 1) // 2,) * 2), mode=mode)
        c = sum(np.einsum("oc,bct->bot", w_c[:, :, t], cp[:, :, t:t + c.shape[2]]) for t in range(kc)) + b_c[None, :, None]
        act = cfg.wn_cond_act_func.lower()
        if act == "lrelu":
            c = np.maximum(c, 0)                                         # 'lrelu' maps to relu, glow.py:93-94
        elif act == "tanh":
            c = np.tanh(c)
        elif act == "sigmoid":
            c = 1.0 / (1.0 + np.exp(-c))
        else:
            assert act == "none", act
    out = np.zeros_like(h)
    pad = cfg.first_pad_value if k == 0 else 0.0
    for i, d in enumerate(cfg.dilations()):
        if cfg.wn_seperable_conv and cfg.wn_kernel_size > 1:            # depthwise then pointwise, glow.py:115-121
            dw = _eff(sd, p + f"in_layers.{i}.0")                        # [C, 1, k]
            ks = dw.shape[2]
            pp = (ks * d - d) // 2
            hp = np.full(h.shape[:2] + (h.shape[2] + 2 * pp,), pad, dtype=h.dtype)
            hp[:, :, pp:pp + h.shape[2]] = h
            dwo = sum(dw[None, :, 0, t, None] * hp[:, :, t * d:t * d + h.shape[2]] for t in range(ks)) \
                + sd[p + f"in_layers.{i}.0.bias"].astype(np.float64)[None, :, None]
            pre = _conv1d(dwo, _eff(sd, p + f"in_layers.{i}.1"), sd[p + f"in_layers.{i}.1.bias"].astype(np.float64))
        else:
            pre = _conv1d(h, _eff(sd, p + f"in_layers.{i}"), sd[p + f"in_layers.{i}.bias"].astype(np.float64), d, pad)
        pre = pre + c[:, 2 * C * i:2 * C * (i + 1)]
        acts = np.tanh(pre[:, :C]) / (1.0 + np.exp(-pre[:, C:]))
        if cfg.wn_res_skip:
            rs = _conv1d(acts, _eff(sd, p + f"res_skip_layers.{i}"), sd[p + f"res_skip_layers.{i}.bias"].astype(np.float64))
        else:
            rs = acts
        if cfg.wn_merge_res_skip:
            h = h + rs
        elif i < cfg.wn_n_layers - 1:
            h = h + rs[:, :C]
            out = out + rs[:, C:]
        else:
            out = out + rs
    if cfg.wn_merge_res_skip:
        out = h
    e = _conv1d(out, sd[p + "end.weight"].astype(np.float64), sd[p + "end.bias"].astype(np.float64))
    n = e.shape[1] // 2
    return e[:, :n], e[:, n:]


def inverse(sd: Dict[str, np.ndarray], cfg: FlowDecoderConfig, z: np.ndarray, cond: np.ndarray) -> np.ndarray:
    """FlowDecoder.inverse (glow.py:302-343) in fp64.  z [B, n_mel, frames] (already scaled by sigma), cond [B, Cc, T]."""
    B = z.shape[0]
    z = z.astype(np.float64).reshape(B, cfg.n_group, -1)
    cond = cond.astype(np.float64)
    sizes = []
    for k in range(cfg.n_flows):
        if k % cfg.n_early_every == 0 and k > 0:
            sizes.append(cfg.n_early_size)
    n_last = cfg.flow_channels()[-1]
    parts, off = [], 0
    for s in sizes + [n_last]:
        parts.append(z[:, off:off + s].copy())
        off += s
    *remained, z = parts
    for k in range(cfg.n_flows - 1, -1, -1):
        Winv = np.linalg.inv(sd[f"convinv.{k}.weight"].astype(np.float64)[:, :, 0])
        if not cfg.mix_first:
            z = np.einsum("oc,bct->bot", Winv, z)
        n_half = z.shape[1] // 2
        z0, z1 = z[:, :n_half], z[:, n_half:]
        log_s, t = wn_forward(sd, k, cfg, z0, cond)
        z = np.concatenate([z0, (z1 - t) / np.exp(log_s)], axis=1)
        if cfg.mix_first:
            z = np.einsum("oc,bct->bot", Winv, z)
        if k % cfg.n_early_every == 0 and k:
            z = np.concatenate([remained.pop(), z], axis=1)
    return z.reshape(B, cfg.n_mel_channels, -1)


def synthetic_state_dict(cfg: FlowDecoderConfig, seed: int) -> Dict[str, np.ndarray]:
    """Seeded checkpoint with the reference FlowDecoder's state_dict layout; `end` is non-zero (the reference
    zero-initialises it, glow.py:62-66, which would make every WN a no-op)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}

    def wn_conv(prefix, co, ci, k):
        bound = 1.0 / np.sqrt(ci * k)
        v = rs.uniform(-bound, bound, size=(co, ci, k)).astype(np.float32)
        norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
        sd[prefix + ".bias"] = rs.uniform(-bound, bound, size=(co,)).astype(np.float32)
        sd[prefix + ".weight_g"] = (norm * rs.uniform(0.8, 1.2, size=norm.shape)).astype(np.float32)
        sd[prefix + ".weight_v"] = v

    C, L, ks = cfg.wn_n_channels, cfg.wn_n_layers, cfg.wn_kernel_size
    for k, n_rem in enumerate(cfg.flow_channels()):
        n_half = n_rem // 2
        q1, _ = np.linalg.qr(rs.standard_normal((n_rem, n_rem)))
        q2, _ = np.linalg.qr(rs.standard_normal((n_rem, n_rem)))
        sd[f"convinv.{k}.weight"] = (q1 @ np.diag(rs.uniform(0.7, 1.4, size=n_rem)) @ q2).astype(np.float32)[:, :, None]
        p = f"WN.{k}.WN"
        for i in range(L):
            if cfg.wn_seperable_conv and ks > 1:
                wn_conv(p + f".in_layers.{i}.0", C, 1, ks)
                wn_conv(p + f".in_layers.{i}.1", 2 * C, C, 1)
            else:
                wn_conv(p + f".in_layers.{i}", 2 * C, C, ks)
        wn_conv(p + ".start", C, n_half, 1)
        sd[p + ".end.weight"] = (rs.standard_normal((2 * n_half, C, 1)) * cfg.end_std).astype(np.float32)
        sd[p + ".end.bias"] = (rs.standard_normal((2 * n_half,)) * cfg.end_std).astype(np.float32)
        dims = [cfg.cond_channels] + [cfg.wn_cond_hidden_channels] * (cfg.wn_cond_layers - 1) + [2 * C * L]
        for j, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
            wn_conv(p + f".cond_layers.{j}", co, ci, cfg.wn_cond_kernel_size)
        if cfg.wn_res_skip:
            for i in range(L):
                wn_conv(p + f".res_skip_layers.{i}", cfg.rs_channels(i), C, 1)
    return sd


def synthetic_inputs(cfg: FlowDecoderConfig, batch: int, frames: int, seed: int):
    rs = np.random.RandomState(seed)
    T = cfg.n_mel_channels * frames // cfg.n_group
    cond = rs.standard_normal((batch, cfg.cond_channels, T)).astype(np.float32)
    z = rs.standard_normal((batch, cfg.n_mel_channels, frames)).astype(np.float32)
    return cond, z


def snr_db(ref, out):
    ref, out = np.asarray(ref, np.float64), np.asarray(out, np.float64)
    return float(10 * np.log10((ref ** 2).sum() / max(((ref - out) ** 2).sum(), 1e-300)))
