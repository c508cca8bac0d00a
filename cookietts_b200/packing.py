"""Weight packing for the B200 WaveGlow inverse pass (host-side, numpy fp64).

Turns a reference `state_dict` (`CookieTTS/_4_mtw/waveglow/glow.py`
:136-186 for WN, :74-83 for the invertible 1x1 conv, :238-241 for the upsampler) into the
flat arrays `include/cwg.h::cwg_weights` describes.  Three exact algebraic rewrites are
applied once, in fp64, so the kernels do less work per group-step:

1. weight-norm is folded: w = g * v / ||v||  (the reference re-evaluates it on every
   forward because `remove_weightnorm` is broken, glow.py:352-360);
2. the purely linear cond chain `cond_layers[0..1]` (glow.py:198-199) is folded with the
   ConvTranspose1d upsampler + squeeze (glow.py:318-324) into one per-phase map from
   J = win/hop mel frames to the H=256 hidden cond vector; `cond_layers[2]` stays as extra K
   columns of every layer's in_layer GEMM;
3. `end` (glow.py:222) is linear, so `end(sum_i skip_i)` becomes `sum_i (W_end W_skip_i) acts_i`
   and the C-wide skip accumulator shrinks to 2*n_half <= 16 columns.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, List, Tuple

import numpy as np

EO_PAD = 16          # narrow layout (n_group <= 16), also what the WaveFlow kernels use
MAX_GROUP = 32


def group_pad(n_group: int) -> int:
    """include/cwg.h CWG_GROUP_PAD: padding MG of the arrays indexed by latent channels (start_w [F][C][MG/2],
    winv [F][MG][MG], eo_b [F][MG], w2 [F][L][C + MG][C]): 16, or 32 for 16 < n_group <= 32."""
    return 16 if n_group <= 16 else 32


@dataclass(frozen=True)
class PackConfig:
    n_mel: int = 80
    n_flows: int = 12
    n_group: int = 8
    n_early_every: int = 4
    n_early_size: int = 2
    win_length: int = 1024
    hop_length: int = 256
    n_layers: int = 8
    n_channels: int = 256
    kernel_size: int = 3
    cond_hidden: int = 256
    speaker_embed_dim: int = 0
    rezero: bool = False

    def flow_channels(self) -> List[Tuple[int, int]]:
        """(n_remaining_channels, n_half) per flow, glow.py:251-264."""
        out, n_half, n_rem = [], self.n_group // 2, self.n_group
        for k in range(self.n_flows):
            if k % self.n_early_every == 0 and k > 0:
                n_half -= self.n_early_size // 2
                n_rem -= self.n_early_size
            out.append((n_rem, n_half))
        return out

    @property
    def phases(self) -> int:
        return self.hop_length // self.n_group

    @property
    def taps(self) -> int:
        return -(-self.win_length // self.hop_length)

    @property
    def k1(self) -> int:
        return self.kernel_size * self.n_channels + self.cond_hidden

    def validate(self) -> None:
        if self.n_group % 2 or self.n_group > MAX_GROUP:
            raise ValueError(f"n_group must be even and <= {MAX_GROUP}, got {self.n_group}")
        if self.hop_length % self.n_group:
            raise ValueError("hop_length must be a multiple of n_group")
        if self.kernel_size % 2 != 1:
            raise ValueError("kernel_size must be odd (glow.py:120)")
        if self.n_channels % 2:
            raise ValueError("n_channels must be even (glow.py:121)")
        if self.n_early_size % 2:
            raise ValueError("n_early_size must be even")
        n_rem, n_half = self.flow_channels()[-1]
        if n_half < 1:
            raise ValueError("too many early outputs for n_group")


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def split_hi_lo(w64: np.ndarray):
    """w ~= hi + lo with both in bf16 (lo = bf16(w - hi)); returns uint16 bit planes."""
    w32 = w64.astype(np.float32)
    hi = f32_to_bf16_bits(w32)
    lo = f32_to_bf16_bits((w64 - bf16_bits_to_f32(hi).astype(np.float64)).astype(np.float32))
    return hi, lo


# ---- "f16f8" mode (CWG_MODE_F16F8): q ~= fp16 hi + fp16 lo for the operands of 3-pass GEMMs, and for the in_layer
# GEMM one fp16 pass plus two fp8 (e5m2) correction passes: a*w ~= a16*w16 + e5m2(a_lo*2^P)*e5m2(w16*2^-P)
#                                                                        + e5m2(a16*2^-Q)*e5m2(w_lo*2^Q)
F8_P, F8_Q = 6, 8


def f32_to_e5m2_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> e5m2 (what cvt.rn.satfinite.e5m2x2.f32 does), uint8 bit patterns."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    t = t.clamp(-57344.0, 57344.0)
    return t.to(torch.float8_e5m2).view(torch.uint8).numpy()


def e5m2_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint16) << 8).view(np.float16).astype(np.float32)


def split_f16(w64: np.ndarray):
    """w ~= hi + lo with both in fp16; returns uint16 bit planes.  numpy converts float64 -> float16 with ONE rounding,
    like the device's cvt.rn.f16.f64 (cwg_pack.cu); torch's CPU conversion goes through float32 and rounds twice, which
    moves values on an fp16 tie by one ulp - up to 6e-8 in the subnormal range the lo planes of small weights live in."""
    w64 = np.ascontiguousarray(w64, dtype=np.float64)
    hi = w64.astype(np.float16)
    lo = (w64 - hi.astype(np.float64)).astype(np.float16)
    return hi.view(np.uint16), lo.view(np.uint16)


def f8_correction_planes(w64: np.ndarray):
    """(h8, l8) = (e5m2(w16 * 2^-P), e5m2((w - w16) * 2^Q)) as uint8 bit planes."""
    import torch
    w = torch.from_numpy(np.ascontiguousarray(w64, dtype=np.float64))
    w16 = torch.from_numpy(w64.astype(np.float16).astype(np.float64))      # one rounding (see split_f16)
    h = (w16 * 2.0 ** -F8_P).to(torch.float32).clamp_(-57344.0, 57344.0).to(torch.float8_e5m2).view(torch.uint8).numpy()
    l = ((w - w16) * 2.0 ** F8_Q).to(torch.float32).clamp_(-57344.0, 57344.0).to(torch.float8_e5m2).view(torch.uint8).numpy()
    return h, l


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float64)


def effective_weight(sd, prefix: str) -> np.ndarray:
    """Accepts the weight-normed (`weight_g`/`weight_v`) or plain (`weight`) layout."""
    if prefix + ".weight_g" in sd:
        g, v = _np(sd[prefix + ".weight_g"]), _np(sd[prefix + ".weight_v"])
        norm = np.sqrt((v ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
        return g * v / norm
    return _np(sd[prefix + ".weight"])


def in_layer_weight_bias(sd, prefix: str):
    """Dense (weight, bias) of one WN in_layer.  The ax models' `seperable_conv` in_layer is
    `nn.Sequential(depthwise, pointwise)` with nothing in between (glow_ax.py:350-358, 2-D: :525-531), i.e. a dense
    conv with W[o, c, ...] = P[o, c] * D[c, 0, ...] and bias P @ b_d + b_p: folded here (exact) so that it runs on
    the same implicit-GEMM kernels."""
    if (prefix + ".0.weight_g" in sd) or (prefix + ".0.weight" in sd):
        d = effective_weight(sd, prefix + ".0")                   # [C, 1, k] or [C, 1, kh, kw]
        pw = effective_weight(sd, prefix + ".1")                  # [2C, C, 1] or [2C, C, 1, 1]
        pw2 = pw.reshape(pw.shape[0], pw.shape[1])
        w = pw2.reshape(pw2.shape + (1,) * (d.ndim - 2)) * d[:, 0][None]
        b = pw2 @ _np(sd[prefix + ".0.bias"]) + _np(sd[prefix + ".1.bias"])
        return w, b
    return effective_weight(sd, prefix), _np(sd[prefix + ".bias"])


def pack_state_dict(sd, cfg: PackConfig, planes=("f32", "hi", "lo")) -> Dict[str, np.ndarray]:
    """Returns the arrays of `cwg_weights` (numpy, host) plus
    `cond_b_base` [F][H] and, for multispeaker models, `cond_w_spk` [F][H][E]
    (the speaker part of the folded cond chain, glow.py:193-196)."""
    cfg.validate()
    F, L, C, H = cfg.n_flows, cfg.n_layers, cfg.n_channels, cfg.cond_hidden
    M, G, P, J, ks = cfg.n_mel, cfg.n_group, cfg.phases, cfg.taps, cfg.kernel_size
    MG = group_pad(G)
    K1, N2, E = cfg.k1, C + MG, cfg.speaker_embed_dim

    w_up = _np(sd["upsample.weight"])            # [M_in, M_out, win]
    b_up = _np(sd["upsample.bias"])
    if w_up.shape != (M, M, cfg.win_length):
        raise ValueError(f"upsample.weight shape {w_up.shape}: only upsample_mode='normal' is supported")
    w_up_p = np.zeros((M, M, J * cfg.hop_length))
    w_up_p[:, :, :cfg.win_length] = w_up
    w_up5 = w_up_p.reshape(M, M, J, P, G)        # index = j*hop + p*G + g

    KCp = -(-(J * M) // 64) * 64          # row pitch of cond_w: K zero-padded to a multiple of 64
    cond_w = np.zeros((F, P * H, KCp))
    cond_b = np.zeros((F, H))
    cond_w_spk = np.zeros((F, H, max(E, 1)))
    w1 = np.zeros((F, L, 2 * C, K1))
    b1 = np.zeros((F, L, 2 * C))
    w2 = np.zeros((F, L, N2, C))
    b2 = np.zeros((F, L, C))
    eo_b = np.zeros((F, MG))
    w0 = np.zeros((F, 2 * C, 48))                 # layer-0 fold: in_layers.0 * start, col = tap*16 + j (see include/cwg.h)
    start_w = np.zeros((F, C, MG // 2))
    start_b = np.zeros((F, C))
    winv = np.zeros((F, MG, MG))

    for k, (n_rem, n_half) in enumerate(cfg.flow_channels()):
        p = f"WN.{k}."
        # --- cond chain folded with upsample + squeeze -------------------------------
        c0 = effective_weight(sd, p + "cond_layers.0")[:, :, 0]     # [H, M*G + E]
        c1 = effective_weight(sd, p + "cond_layers.1")[:, :, 0]     # [H, H]
        c2 = effective_weight(sd, p + "cond_layers.2")[:, :, 0]     # [2CL, H]
        cb0, cb1, cb2 = (_np(sd[p + f"cond_layers.{j}.bias"]) for j in range(3))
        if c0.shape[0] != H or c2.shape != (2 * C * L, H):
            raise ValueError("unexpected cond_layers shapes")
        w21 = c1 @ c0                                                # [H, M*G + E]
        w21_mel = w21[:, :M * G].reshape(H, M, G)
        a = np.einsum("hmg,cmjpg->phjc", w21_mel, w_up5, optimize=True)   # [P, H, J, M_in]
        cond_w[k, :, :J * M] = a.reshape(P * H, J * M)
        cond_b[k] = np.einsum("hmg,m->h", w21_mel, b_up) + c1 @ cb0 + cb1
        if E:
            cond_w_spk[k] = w21[:, M * G:]
        # --- per layer ---------------------------------------------------------------
        w_end = _np(sd[p + "end.weight"])[:, :, 0]                   # [2 n_half, C]
        b_end = _np(sd[p + "end.bias"])
        eo_bias = b_end.copy()
        for i in range(L):
            w_in = effective_weight(sd, p + f"in_layers.{i}")        # [2C, C, ks]
            w1[k, i, :, :ks * C] = w_in.transpose(0, 2, 1).reshape(2 * C, ks * C)   # col = tap*C + c
            w1[k, i, :, ks * C:] = c2[2 * C * i:2 * C * (i + 1)]
            b1[k, i] = _np(sd[p + f"in_layers.{i}.bias"]) + cb2[2 * C * i:2 * C * (i + 1)]
            w_rs = effective_weight(sd, p + f"res_skip_layers.{i}")[:, :, 0]
            b_rs = _np(sd[p + f"res_skip_layers.{i}.bias"])
            if cfg.rezero:                                           # glow.py:211-212
                alpha = float(_np(sd[p + f"alpha_i.{i}"]).reshape(-1)[0])
                w_rs, b_rs = w_rs * alpha, b_rs * alpha
            if i < L - 1:
                w2[k, i, :C] = w_rs[:C]
                b2[k, i] = b_rs[:C]
                w_skip, b_skip = w_rs[C:], b_rs[C:]
            else:
                w_skip, b_skip = w_rs, b_rs
            w2[k, i, C:C + 2 * n_half] = w_end @ w_skip
            eo_bias += w_end @ b_skip
        eo_b[k, :2 * n_half] = eo_bias
        # --- start / inverse 1x1 -------------------------------------------------------
        start_w[k, :, :n_half] = effective_weight(sd, p + "start")[:, :, 0]
        start_b[k] = _np(sd[p + "start.bias"])
        if ks == 3 and MG == 16:
            w_in0 = effective_weight(sd, p + "in_layers.0")              # [2C, C, 3]
            for tap in range(3):
                w0[k, :, tap * 16:tap * 16 + n_half] = w_in0[:, :, tap] @ start_w[k, :, :n_half]
                w0[k, :, tap * 16 + n_half] = w_in0[:, :, tap] @ start_b[k]
        W = _np(sd[f"convinv.{k}.conv.weight"]).reshape(n_rem, n_rem)
        # the reference inverts in fp32 (glow.py:93); fp64 here is closer to the exact inverse
        winv[k, :n_rem, :n_rem] = np.linalg.inv(W)

    out: Dict[str, np.ndarray] = {
        "b1": b1.astype(np.float32), "b2": b2.astype(np.float32), "eo_b": eo_b.astype(np.float32),
        "start_w": start_w.astype(np.float32), "start_b": start_b.astype(np.float32),
        "winv": winv.astype(np.float32),
        "cond_b_base": cond_b.astype(np.float32), "cond_w_spk": cond_w_spk.astype(np.float32),
    }
    if ks == 3 and C == 256 and MG == 16 and "f32" not in planes:
        out["w0_hi"], out["w0_lo"] = split_f16(w0) if "f16f8" in planes else split_hi_lo(w0)
    for name, arr in (("cond_w", cond_w), ("w1", w1), ("w2", w2)):
        if "f32" in planes:
            out[name + "_f32"] = arr.astype(np.float32)
        if "f16f8" in planes:
            out[name + "_hi"], out[name + "_lo"] = split_f16(arr)
            if name in ("w1", "w2"):          # GEMMs issued as one fp16 pass + two e5m2 correction passes
                out[name + "_h8"], out[name + "_l8"] = f8_correction_planes(arr)
        elif "hi" in planes or "lo" in planes:
            hi, lo = split_hi_lo(arr)
            out[name + "_hi"] = hi
            if "lo" in planes:
                out[name + "_lo"] = lo
    return out


def config_dict(cfg: PackConfig) -> dict:
    return asdict(cfg)
