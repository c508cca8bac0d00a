"""CPU oracle for the WaveFlow inverse pass of the reference's "ax" model (BASELINE config 5).

TEST INFRASTRUCTURE ONLY (same rules as waveglow_oracle.py).  numpy restatement of, in
/root/reference/CookieTTS/_4_mtw/waveglow/:
  efficient_model_ax.py:279-357  WaveGlow.inverse (explicit z)           -> `inverse`
  efficient_model_ax.py:359-388  WaveGlow.infer (zero-frame pad, trim)   -> `infer_with_z`
  efficient_model_ax.py:171-182  _upsample_mels (F.interpolate)          -> `upsample_cond`
  efficient_modules.py:42-65     WaveFlowCoupling.inverse (AR over height)-> `coupling_inverse`
  glow_ax.py:556-635             WN_2d.forward with conv queues          -> `wn2d_step`
  efficient_modules.py:360-403   PermuteHeight.inverse                   -> `permute_height`
for the configuration subset the B200 build supports (and config 5 pins):
waveflow=True, channel_mixing='permuteheight', mix_first=True, upsample_first=True, no model-level
cond layers / upsample net / speaker embedding, WN_2d with one 1x1 cond layer, full (non-separable)
kernel (kh, kw), dilation_h = 1, dilation_w = 2^i, GTU gate, res_skip without merge - plus the WN_config variants the fp32
CUDA-core mode runs: the 14 gated units (glow_ax.py:168-198), listed width / height dilations (:506-517), merge_res_skip /
res_skip=False (:541-553,:610-626), WN-level speaker embeddings (:464-466,:567-570), upsample_first=False (:578-579).

Parity status: PINNED - `oracle/make_golden_waveflow.py` runs the unmodified reference model
(`inverse` with explicit z and `infer` with the RNG draw replaced) and stores
tests/golden/waveflow_*.npz; tests/test_waveflow_oracle.py checks this file against them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np

from .waveglow_oracle import weight_norm_effective


from cookietts_b200.synthetic import WaveFlowConfig, waveflow_state_dict as synthetic_state_dict  # noqa: E402,F401


def _w(sd, prefix, dtype):
    if prefix + ".weight_g" in sd:
        return weight_norm_effective(np.asarray(sd[prefix + ".weight_g"], np.float64),
                                     np.asarray(sd[prefix + ".weight_v"], np.float64)).astype(dtype)
    return np.asarray(sd[prefix + ".weight"], dtype)


def upsample_cond(cond: np.ndarray, size: int, mode: str) -> np.ndarray:
    """F.interpolate(cond [B,C,Tm], size, mode, align_corners=True if linear)
    (efficient_model_ax.py:174-175)."""
    B, C, Tm = cond.shape
    if mode == "linear":
        src = np.arange(size, dtype=np.float64) * ((Tm - 1) / (size - 1) if size > 1 else 0.0)
        i0 = np.minimum(np.floor(src).astype(np.int64), Tm - 1)
        i1 = np.minimum(i0 + 1, Tm - 1)
        w = (src - i0).astype(cond.dtype)
        return cond[:, :, i0] * (1 - w) + cond[:, :, i1] * w
    if mode == "nearest":
        idx = np.minimum(np.floor(np.arange(size) * (Tm / size)).astype(np.int64), Tm - 1)
        return cond[:, :, idx]
    raise NotImplementedError(mode)


def permute_height(x: np.ndarray, k: int) -> np.ndarray:
    """PermuteHeight(k).inverse == forward (an involution): full reverse, or for k % 4 in (2,3)
    reverse each half (efficient_modules.py:341-353,377-383,400-403)."""
    h = x.shape[1]
    idx = list(range(h))
    if k % 4 in (2, 3):
        half = h // 2
        idx = idx[:half][::-1] + idx[half:][::-1]
    else:
        idx = idx[::-1]
    return x[:, idx]


def _gated_units():
    from .waveglow_ax_oracle import GATED_UNITS
    return GATED_UNITS


def wn2d_step(sd, k, cfg: WaveFlowConfig, row: np.ndarray, spec_all: np.ndarray, queues, dtype):
    """One autoregressive step of WN_2d (glow_ax.py:556-635) with conv queues.
    row [B, T'] is the newest height row; spec_all [B, 2CL, T'] the cond-layer output;
    queues[i] holds the last (kh-1) input rows of layer i, [B, C, kh-1, T'] (zeros at start).
    Returns (log_s, t) each [B, T']."""
    p = f"WN.{k}.WN."
    C, L, kh, kw = cfg.n_channels, cfg.n_layers, cfg.kernel_size_h, cfg.kernel_size_w
    B, T = row.shape
    w_s = _w(sd, p + "start", dtype).reshape(C)
    audio = w_s[None, :, None] * row[:, None, :] + np.asarray(sd[p + "start.bias"], dtype)[None, :, None]   # :558
    output = np.zeros_like(audio)
    unit_a, unit_b = _gated_units()[cfg.gated_unit.upper()]
    split = cfg.res_skip and not cfg.merge_res_skip
    for i in range(L):
        d, dh = cfg.dilation_w(i), cfg.dilation_h(i)
        pad_h = (kh - 1) * dh                                            # :517
        if queues[i] is None:                                            # :597-599
            queues[i] = np.zeros((B, C, pad_h, T), dtype)
        full = np.concatenate([queues[i], audio[:, :, None, :]], axis=2)    # [B, C, pad_h + 1, T']  :602
        queues[i] = full[:, :, 1:]
        stack = full[:, :, ::dh]                                         # the kh rows a height-dilated kernel touches
        sep = (p + f"in_layers.{i}.0.weight_v") in sd
        w_in = _w(sd, p + (f"in_layers.{i}.0" if sep else f"in_layers.{i}"), dtype)   # [2C, C, kh, kw] / depthwise [C, 1, kh, kw]
        pad = ((kw - 1) * d) // 2
        sp = np.zeros((B, C, kh, T + 2 * pad), dtype)
        sp[:, :, :, pad:pad + T] = stack
        if sep:                                                          # depthwise, then pointwise (no activation between)
            dw = np.zeros((B, C, T), dtype)
            for a in range(kh):
                for b in range(kw):
                    dw += w_in[None, :, 0, a, b, None] * sp[:, :, a, b * d:b * d + T]
            dw += np.asarray(sd[p + f"in_layers.{i}.0.bias"], dtype)[None, :, None]
            acts = np.einsum("oc,bct->bot", _w(sd, p + f"in_layers.{i}.1", dtype)[:, :, 0, 0], dw, optimize=True) \
                + np.asarray(sd[p + f"in_layers.{i}.1.bias"], dtype)[None, :, None]
        else:
            acts = np.zeros((B, 2 * C, T), dtype)
            for a in range(kh):
                for b in range(kw):
                    acts += np.einsum("oc,bct->bot", w_in[:, :, a, b], sp[:, :, a, b * d:b * d + T], optimize=True)
            acts += np.asarray(sd[p + f"in_layers.{i}.bias"], dtype)[None, :, None]
        acts += spec_all[:, 2 * C * i:2 * C * (i + 1)]                   # :585-608 (GTU: add, tanh*sigmoid)
        g = unit_a(acts[:, :C]) * unit_b(acts[:, C:])
        if cfg.res_skip:
            w_rs = _w(sd, p + f"res_skip_layers.{i}", dtype)[:, :, 0, 0]
            rs = np.einsum("oc,bct->bot", w_rs, g, optimize=True) + np.asarray(sd[p + f"res_skip_layers.{i}.bias"], dtype)[None, :, None]
        else:
            rs = g                                                       # :610
        if split and i < L - 1:                                          # :613-626 (merged: the hidden tensor stays)
            audio = audio + rs[:, :C]
            output = output + rs[:, C:]
        else:
            output = output + rs
    w_e = np.asarray(sd[p + "end.weight"], dtype)[:, :, 0, 0]            # [2, C]
    out = np.einsum("oc,bct->bot", w_e, output, optimize=True) + np.asarray(sd[p + "end.bias"], dtype)[None, :, None]
    return out[:, 0], out[:, 1]                                          # log_s, t  (:628, efficient_modules.py:62)


def coupling_inverse(sd, k, cfg, audio_out: np.ndarray, spec_all: np.ndarray, dtype) -> np.ndarray:
    """WaveFlowCoupling.inverse (efficient_modules.py:42-65): audio_out [B, h, T'] -> z."""
    h = audio_out.shape[1]
    z = [audio_out[:, 0]]
    queues = [None] * cfg.n_layers
    for i in range(h - 1):
        log_s, t = wn2d_step(sd, k, cfg, z[-1], spec_all, queues, dtype)
        z.append((audio_out[:, i + 1] - t) / np.exp(log_s))
    return np.stack(z, axis=1)


def inverse(sd, cfg: WaveFlowConfig, z: np.ndarray, cond: np.ndarray, dtype=np.float32, cond_up=None, speaker_ids=None) -> np.ndarray:
    """WaveGlow.inverse(z, cond) (efficient_model_ax.py:279-357): z [B, T] (already scaled by
    sigma), cond [B, n_mel, frames] -> audio [B, T].  `cond_up` (one [B, C, T'] array, or one per flow) replaces
    the plain interpolation when the model has a conditioning front-end (oracle/ax_frontend_oracle.py)."""
    z = np.asarray(z, dtype)
    B = z.shape[0]
    zz = z.reshape(B, -1, cfg.n_group).transpose(0, 2, 1)               # :310
    Tp = zz.shape[2]
    if cond_up is None:
        cond_up = np.asarray(cond, dtype)
        if cfg.upsample_first:
            cond_up = upsample_cond(cond_up, Tp, cfg.upsample_mode)      # :313-314
    C, L = cfg.n_channels, cfg.n_layers
    rows = cfg.flow_rows()
    n_early = sum(1 for k in range(cfg.n_flows) if cfg.n_early_every and k % cfg.n_early_every == 0 and k > 0)
    remained, off = [], 0
    for _ in range(n_early):                                             # :319-322
        remained.append(zz[:, off:off + cfg.n_early_size]); off += cfg.n_early_size
    zz = zz[:, off:]
    assert zz.shape[1] == rows[-1]
    for k in reversed(range(cfg.n_flows)):                               # :325
        k_cond = cond_up[k] if isinstance(cond_up, (list, tuple)) else cond_up      # :328
        from .waveglow_ax_oracle import wn_cond_path                     # speaker embedding, cond stack, WN-level upsampling
        spec_all = wn_cond_path(sd, f"WN.{k}.WN.", cfg, k_cond, Tp, dtype, speaker_ids, crop_2d=True)   # glow_ax.py:565-579
        def mix(x):                                                      # PermuteHeight.inverse or InvertibleConv1x1.inverse
            if cfg.channel_mixing == "permuteheight":
                return permute_height(x, k)
            W = np.asarray(sd[f"convinv.{k}.weight"], np.float64)[:, :, 0]      # efficient_modules.py:269-286
            W_inv = np.linalg.inv(W) if dtype == np.float64 else np.linalg.inv(W.astype(np.float32))
            return np.einsum("oc,bct->bot", W_inv.astype(dtype), x, optimize=True)
        if not cfg.mix_first:
            zz = mix(zz)                                                 # :326-327
        zz = coupling_inverse(sd, k, cfg, zz, spec_all, dtype)           # :331
        if cfg.mix_first:
            zz = mix(zz)                                                 # :336-337
        if cfg.n_early_every and k % cfg.n_early_every == 0 and k:
            zz = np.concatenate([remained.pop(), zz], axis=1)            # :339-340
    return np.ascontiguousarray(zz.transpose(0, 2, 1)).reshape(B, -1)    # :346


def infer_with_z(sd, cfg: WaveFlowConfig, spect: np.ndarray, z: np.ndarray, sigma: float,
                 artifact_trimming: int = 1, dtype=np.float32) -> np.ndarray:
    """WaveGlow.infer (efficient_model_ax.py:359-388) with the latent passed in: z is standard
    normal [B, frames*hop]; returns [B, (frames*hop) - artifact_trimming*hop... ] exactly as the
    reference: pad `artifact_trimming` zero frames, run inverse on (steps-1)*hop samples, trim."""
    spect = np.asarray(spect, dtype)
    if artifact_trimming > 0:
        spect = np.concatenate([spect, np.zeros(spect.shape[:2] + (artifact_trimming,), dtype)], axis=2)   # :370-371
    steps = spect.shape[2]
    samples = (steps - 1) * cfg.hop_length
    samples -= samples % cfg.n_group
    assert z.shape[1] == samples, (z.shape, samples)
    audio = inverse(sd, cfg, np.asarray(z, dtype) * dtype(sigma), spect, dtype)
    if artifact_trimming > 0:
        audio = audio[:, :-artifact_trimming * cfg.hop_length]           # :381-383
    return audio
