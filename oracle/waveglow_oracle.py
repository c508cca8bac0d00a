"""CPU oracle for the WaveGlow inverse pass (classic `WaveGlow.infer`).

TEST INFRASTRUCTURE ONLY.  This file is a numpy restatement of the reference
algorithm in `/root/reference/CookieTTS/_4_mtw/waveglow/glow.py`.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference arm
may import it; the product package `cookietts_b200` never does.

Parity status: PINNED.  `oracle/make_golden.py` runs the unmodified reference
`WaveGlow.infer` (imported in place from /root/reference, RNG draws replaced by
pre-drawn z) and stores its outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against those vectors.

Every function cites the reference lines it restates.  Layout follows the
reference (channel-major `[B, C, T]`); z is explicit (the reference draws it
inside `infer`, glow.py:326,343-347).

z convention used across this repo: `z` is `[B, T]` (audio shaped, T = T_mel*hop),
standard normal, and the latent channel c of group-step s is `z[b, s*n_group + c]`.
Channels are ordered as the reference stacks them at the end of `infer`
(SURVEY Appendix C): [early drawn at the lowest k, ..., early drawn at the highest k,
main].  `split_z` converts to the per-draw tensors the reference would have drawn.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np


from cookietts_b200.synthetic import ModelConfig as OracleConfig, synthetic_state_dict, synthetic_inputs  # noqa: E402,F401


def weight_norm_effective(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """`torch.nn.utils.weight_norm(dim=0)`: w = g * v / ||v||, norm over all dims but 0
    (used by glow.py:136-137,155-165,171-173,181-182)."""
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return (g.astype(np.float64) * v.astype(np.float64) / norm).astype(v.dtype)


def _conv_weight(sd: Dict[str, np.ndarray], prefix: str, dtype) -> np.ndarray:
    """Effective conv weight for `prefix`, accepting both the weight-normed
    (`weight_g`/`weight_v`) and the plain (`weight`) checkpoint layouts."""
    if prefix + ".weight_g" in sd:
        w = weight_norm_effective(np.asarray(sd[prefix + ".weight_g"], dtype=np.float64),
                                  np.asarray(sd[prefix + ".weight_v"], dtype=np.float64))
    else:
        w = np.asarray(sd[prefix + ".weight"], dtype=np.float64)
    return w.astype(dtype)


def conv1x1(x: np.ndarray, w: np.ndarray, b: np.ndarray | None) -> np.ndarray:
    """`nn.Conv1d(kernel_size=1)`: x [B,Ci,T], w [Co,Ci,1] -> [B,Co,T]."""
    y = np.einsum("oc,bct->bot", w[:, :, 0], x, optimize=True)
    if b is not None:
        y = y + b[None, :, None]
    return y


def dilated_conv(x: np.ndarray, w: np.ndarray, b: np.ndarray, dilation: int) -> np.ndarray:
    """`nn.Conv1d(C, 2C, k, dilation=d, padding=(k*d-d)/2)` (glow.py:167-171):
    y[:, :, t] = b + sum_j w[:, :, j] @ x[:, :, t + (j - (k-1)/2)*d], zero padded."""
    B, C, T = x.shape
    k = w.shape[2]
    pad = (k * dilation - dilation) // 2
    xp = np.zeros((B, C, T + 2 * pad), dtype=x.dtype)
    xp[:, :, pad:pad + T] = x
    y = np.zeros((B, w.shape[0], T), dtype=x.dtype)
    for j in range(k):
        y += np.einsum("oc,bct->bot", w[:, :, j], xp[:, :, j * dilation:j * dilation + T], optimize=True)
    return y + b[None, :, None]


def upsample_mel(mel: np.ndarray, w: np.ndarray, b: np.ndarray, hop: int, win: int) -> np.ndarray:
    """`nn.ConvTranspose1d(n_mel, n_mel, win, stride=hop)` followed by dropping the last
    `win - hop` samples (glow.py:238-241,318-321).
    out[b, co, hop*f + r] += sum_ci mel[b, ci, f] * w[ci, co, r]  for r in [0, win)."""
    B, C, Tm = mel.shape
    full = np.zeros((B, w.shape[1], (Tm - 1) * hop + win), dtype=mel.dtype)
    # contribution of every frame f: [B, Co, win] placed at hop*f
    contrib = np.einsum("bcf,cor->bfor", mel, w, optimize=True)
    for f in range(Tm):
        full[:, :, f * hop:f * hop + win] += contrib[:, f]
    full += b[None, :, None]
    cutoff = win - hop
    return full[:, :, :-cutoff] if cutoff > 0 else full


def squeeze_cond(up: np.ndarray, n_group: int) -> np.ndarray:
    """`unfold(2, G, G).permute(0,2,1,3).view(B, T', -1).permute(0,2,1)` (glow.py:323-324):
    cond[b, m*G + g, s] = up[b, m, s*G + g]."""
    B, M, T = up.shape
    Tp = T // n_group
    x = up[:, :, :Tp * n_group].reshape(B, M, Tp, n_group)      # [B, M, T', G]
    return np.ascontiguousarray(x.transpose(0, 1, 3, 2)).reshape(B, M * n_group, Tp)


def gate(a: np.ndarray, b: np.ndarray, n_channels: int) -> np.ndarray:
    """`fused_add_tanh_sigmoid_multiply` (glow.py:34-41)."""
    s = a + b
    return np.tanh(s[:, :n_channels]) * (1.0 / (1.0 + np.exp(-s[:, n_channels:])))


def wn_forward(sd: Dict[str, np.ndarray], k: int, cfg: OracleConfig, audio0: np.ndarray,
               cond: np.ndarray, dtype, speaker_id=None) -> Tuple[np.ndarray, np.ndarray]:
    """`WN.forward` (glow.py:188-222) for flow k; returns (b, s)."""
    p = f"WN.{k}."
    C, L = cfg.n_channels, cfg.n_layers
    x = conv1x1(audio0, _conv_weight(sd, p + "start", dtype), np.asarray(sd[p + "start.bias"], dtype))
    output = np.zeros_like(x)                                            # :190
    spect = cond
    if cfg.speaker_embed_dim and speaker_id is not None:                 # :193-196
        emb = np.asarray(sd[p + "speaker_embed.weight"], dtype)[np.asarray(speaker_id)]      # [B, E]
        emb = np.repeat(emb[:, :, None], cond.shape[2], axis=2)
        spect = np.concatenate([spect, emb], axis=1)
    for j in range(3):                                                   # :198-199 (linear chain)
        spect = conv1x1(spect, _conv_weight(sd, p + f"cond_layers.{j}", dtype),
                        np.asarray(sd[p + f"cond_layers.{j}.bias"], dtype))
    for i in range(L):                                                   # :201-220
        w_in = _conv_weight(sd, p + f"in_layers.{i}", dtype)
        acts = gate(dilated_conv(x, w_in, np.asarray(sd[p + f"in_layers.{i}.bias"], dtype), 2 ** i),
                    spect[:, 2 * C * i:2 * C * (i + 1)], C)
        rs = conv1x1(acts, _conv_weight(sd, p + f"res_skip_layers.{i}", dtype),
                     np.asarray(sd[p + f"res_skip_layers.{i}.bias"], dtype))
        if cfg.rezero:                                                   # :211-212
            rs = rs * np.asarray(sd[p + f"alpha_i.{i}"], dtype).reshape(())
        if i < L - 1:
            x = x + rs[:, :C]
            output = output + rs[:, C:]
        else:
            output = output + rs
    end = conv1x1(output, np.asarray(sd[p + "end.weight"], dtype), np.asarray(sd[p + "end.bias"], dtype))
    n_half = end.shape[1] // 2
    return end[:, :n_half], end[:, n_half:]                              # chunk(2, 1), :222


def split_z(z: np.ndarray, cfg: OracleConfig):
    """[B, T] latent -> (z_main [B, n_rem_last, T'], {k: z_early_k [B, n_early_size, T']}).
    Channel order of the stacked latent: SURVEY Appendix C."""
    B, T = z.shape
    G = cfg.n_group
    zf = z.reshape(B, T // G, G).transpose(0, 2, 1)                     # [B, G, T']
    early_ks = [k for k in range(cfg.n_flows) if k % cfg.n_early_every == 0 and k > 0]
    early = {}
    off = 0
    for k in early_ks:                                                   # lowest k first
        early[k] = zf[:, off:off + cfg.n_early_size]
        off += cfg.n_early_size
    return zf[:, off:], early


def infer_with_z(sd: Dict[str, np.ndarray], cfg: OracleConfig, mel: np.ndarray, z: np.ndarray,
                 sigma: float, dtype=np.float32, speaker_id=None) -> np.ndarray:
    """`WaveGlow.infer` (glow.py:314-350) with the latent passed in.
    mel [B, n_mel, T_mel]; z [B, T_mel*hop] standard normal; returns audio [B, T_mel*hop]."""
    mel = np.asarray(mel, dtype)
    z = np.asarray(z, dtype)
    sigma = dtype(sigma)
    up = upsample_mel(mel, np.asarray(sd["upsample.weight"], dtype), np.asarray(sd["upsample.bias"], dtype),
                      cfg.hop_length, cfg.win_length)                    # :318-321
    cond = squeeze_cond(up, cfg.n_group)                                 # :323-324
    z_main, z_early = split_z(z, cfg)
    audio = sigma * z_main                                               # :326
    for k in reversed(range(cfg.n_flows)):                               # :328
        n_half = audio.shape[1] // 2
        a0, a1 = audio[:, :n_half], audio[:, n_half:]
        b, s = wn_forward(sd, k, cfg, a0, cond, dtype, speaker_id)       # :333
        a1 = (a1 - b) / np.exp(s)                                        # :337
        audio = np.concatenate([a0, a1], axis=1)
        W = np.asarray(sd[f"convinv.{k}.conv.weight"], dtype)[:, :, 0]
        W_inv = np.linalg.inv(W.astype(np.float64) if dtype == np.float64 else W.astype(np.float32))
        audio = np.einsum("oc,bct->bot", W_inv.astype(dtype), audio, optimize=True)   # :85-99
        if k % cfg.n_early_every == 0 and k > 0:                         # :342-347
            audio = np.concatenate([sigma * z_early[k], audio], axis=1)
    B = audio.shape[0]
    return np.ascontiguousarray(audio.transpose(0, 2, 1)).reshape(B, -1)  # :349


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    err = np.asarray(test, np.float64) - ref
    return float(10.0 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-300)))
