"""CPU port of the reference `WaveGlow.infer` on the same torch ops the reference calls.

TEST / BASELINE INFRASTRUCTURE ONLY (same rules as waveglow_oracle.py).  The numpy oracle is
the definition; this port exists because `bench.py`'s CPU baseline should run at the speed of
the reference's own CPU path (torch conv1d / conv_transpose1d on MKL-DNN, all host threads),
and /root/reference cannot travel to the GPU box.  It follows glow.py line by line with the
latent passed in, and `tests/test_oracle_golden.py` pins it to the same golden vectors.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .waveglow_oracle import OracleConfig, split_z


def _w(sd, prefix, dtype):
    if prefix + ".weight_g" in sd:
        g = torch.as_tensor(sd[prefix + ".weight_g"]).to(dtype)
        v = torch.as_tensor(sd[prefix + ".weight_v"]).to(dtype)
        return torch._weight_norm(v, g, 0)               # what nn.utils.weight_norm evaluates
    return torch.as_tensor(sd[prefix + ".weight"]).to(dtype)


def _b(sd, name, dtype):
    return torch.as_tensor(sd[name]).to(dtype)


class TorchPort:
    """Holds the checkpoint tensors; `infer` re-evaluates weight-norm on every call exactly as
    the reference does (its `remove_weightnorm` is broken, glow.py:352-360)."""

    def __init__(self, sd, cfg: OracleConfig, dtype=torch.float32):
        self.sd, self.cfg, self.dtype = sd, cfg, dtype
        self.w_inv = {}
        for k in range(cfg.n_flows):                      # glow.py:91-97 (cached after first use)
            W = torch.as_tensor(sd[f"convinv.{k}.conv.weight"]).to(dtype).squeeze(-1)
            self.w_inv[k] = W.float().inverse().to(dtype)[..., None] if dtype == torch.float32 else W.inverse()[..., None]

    def wn(self, k, audio0, spect):                       # glow.py:188-222
        sd, cfg, dt = self.sd, self.cfg, self.dtype
        p = f"WN.{k}."
        C = cfg.n_channels
        audio = F.conv1d(audio0, _w(sd, p + "start", dt), _b(sd, p + "start.bias", dt))
        output = torch.zeros_like(audio)
        for j in range(3):
            spect = F.conv1d(spect, _w(sd, p + f"cond_layers.{j}", dt), _b(sd, p + f"cond_layers.{j}.bias", dt))
        for i in range(cfg.n_layers):
            d = 2 ** i
            pad = (cfg.kernel_size * d - d) // 2
            in_act = F.conv1d(audio, _w(sd, p + f"in_layers.{i}", dt), _b(sd, p + f"in_layers.{i}.bias", dt),
                              dilation=d, padding=pad) + spect[:, 2 * C * i:2 * C * (i + 1)]
            acts = torch.tanh(in_act[:, :C]) * torch.sigmoid(in_act[:, C:])       # glow.py:34-41
            rs = F.conv1d(acts, _w(sd, p + f"res_skip_layers.{i}", dt), _b(sd, p + f"res_skip_layers.{i}.bias", dt))
            if cfg.rezero:
                rs = rs * _b(sd, p + f"alpha_i.{i}", dt)
            if i < cfg.n_layers - 1:
                audio = audio + rs[:, :C]
                output = output + rs[:, C:]
            else:
                output = output + rs
        end = F.conv1d(output, _b(sd, p + "end.weight", dt), _b(sd, p + "end.bias", dt))
        return end.chunk(2, 1)

    @torch.no_grad()
    def infer(self, mel, z, sigma):                       # glow.py:314-350 with explicit z
        cfg, dt = self.cfg, self.dtype
        mel = torch.as_tensor(mel).to(dt)
        z_main, z_early = split_z(np.asarray(z), cfg)
        spect = F.conv_transpose1d(mel, _b(self.sd, "upsample.weight", dt), _b(self.sd, "upsample.bias", dt),
                                   stride=cfg.hop_length)
        spect = spect[:, :, :-(cfg.win_length - cfg.hop_length)]
        spect = spect.unfold(2, cfg.n_group, cfg.n_group).permute(0, 2, 1, 3)
        spect = spect.contiguous().view(spect.size(0), spect.size(1), -1).permute(0, 2, 1)
        audio = torch.as_tensor(np.ascontiguousarray(z_main)).to(dt) * sigma
        for k in reversed(range(cfg.n_flows)):
            n_half = audio.size(1) // 2
            a0, a1 = audio[:, :n_half], audio[:, n_half:]
            b, s = self.wn(k, a0, spect)
            a1 = (a1 - b) / torch.exp(s)
            audio = torch.cat([a0, a1], 1)
            audio = F.conv1d(audio, self.w_inv[k])
            if k % cfg.n_early_every == 0 and k > 0:
                audio = torch.cat((sigma * torch.as_tensor(np.ascontiguousarray(z_early[k])).to(dt), audio), 1)
        return audio.permute(0, 2, 1).contiguous().view(audio.size(0), -1).numpy()
