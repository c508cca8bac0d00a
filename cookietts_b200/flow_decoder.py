"""Drop-in `FlowDecoder` (inverse pass) for the reference's mel-domain flow decoders:

    CookieTTS/_2_ttm/flowtts/waveglow/glow.py:174-363  FlowDecoder(hparams)   (flowtts)
    CookieTTS/_2_ttm/untts/waveglow/glow.py:178-379    FlowDecoder(hparams)   (untts "MelGlow")

Same constructor (an `hparams` object with the reference's attribute names), same parameter names / shapes (so a
reference checkpoint's decoder state_dict loads with strict=True), same `inverse(z, cond)` / `infer(cond, sigma)` calls.
The computation runs in libcwg.so (`cwg_fd_inverse`, csrc/cwg_fd.cu): fp32 CUDA-core kernels on the reference's
channels-first layout - these decoders run at mel-frame rate, a few GFLOP per utterance.  `forward` (the training
direction) raises; there is no CPU fallback.

Supported: the hparams defaults of both models and their WN variants - any n_group / n_flows / early outputs, mix_first on
or off, 1..16 layers with 2^i / constant / listed dilations, res_skip and merge_res_skip on or off, the untts
`decoder_padding_value`, depthwise-separable in_layers (folded to dense convs at pack time - exact, also under the constant
padding), WN cond stacks of several layers / kernel sizes / padding modes with an activation after every layer (evaluated
by the module with `cwg_conv1d` and handed over as `cwg_fd_weights.c_all`).  Decoder-level cond layers (cond_layers > 0,
cond_residual, cond_res_rezero) raise NotImplementedError: the reference's own constructor fails on them (glow.py:196-201
reads undefined names), so there is nothing to be faithful to.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _cabi

FD_MAX_LAYERS = 16


class CwgFdConfig(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("n_group", "n_flows", "n_early_every", "n_early_size", "mix_first", "cond_channels",
                                           "n_layers", "n_channels", "kernel_size")]
                + [("dilations", C.c_int32 * FD_MAX_LAYERS), ("res_skip", C.c_int32), ("merge_res_skip", C.c_int32),
                   ("first_pad_value", C.c_float)])


FD_WEIGHT_FIELDS = ("start_w", "start_b", "cond_w", "cond_b", "in_w", "in_b", "rs_w", "rs_b", "end_w", "end_b", "winv")
# glow.py:89-101: 'lrelu' is relu; 'relu' (meant to be LeakyReLU(0.2)) raises NameError in the reference's constructor
FD_COND_ACTS = {"none": 0, "lrelu": 1, "tanh": 3, "sigmoid": 4}
FD_PAD_MODES = {"zeros": 0, "replicate": 1, "reflect": 2, "circular": 3}


class CwgFdWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in FD_WEIGHT_FIELDS + ("c_all",)]


def _bind(lib):
    if getattr(lib, "_fd_bound", False):
        return
    lib.cwg_fd_workspace_bytes.restype = C.c_size_t
    lib.cwg_fd_workspace_bytes.argtypes = [C.POINTER(CwgFdConfig), C.c_int, C.c_int]
    lib.cwg_fd_launch_count.restype = C.c_int
    lib.cwg_fd_launch_count.argtypes = [C.POINTER(CwgFdConfig)]
    lib.cwg_fd_inverse.restype = C.c_int
    lib.cwg_fd_inverse.argtypes = [C.POINTER(CwgFdConfig), C.POINTER(CwgFdWeights), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    lib._fd_bound = True


def _eff(sd, prefix):
    """weight-norm folded weight (fp64); accepts the plain `weight` layout too"""
    if prefix + ".weight_g" in sd:
        g, v = sd[prefix + ".weight_g"].double(), sd[prefix + ".weight_v"].double()
        return g * v / v.pow(2).sum(dim=tuple(range(1, v.dim())), keepdim=True).sqrt()
    return sd[prefix + ".weight"].double()


class _WN(nn.Module):
    """Parameter holder with the layout of flowtts glow.py:43-131 (untts: weight_norm global, same keys)."""

    def __init__(self, n_in_channels, cond_in_channels, hp):
        super().__init__()
        wn = nn.utils.weight_norm
        C_, L, ks = hp.wn_n_channels, hp.wn_n_layers, hp.wn_kernel_size
        self.in_layers = nn.ModuleList()
        self.start = wn(nn.Conv1d(n_in_channels, C_, 1), name="weight")
        self.end = nn.Conv1d(C_, 2 * n_in_channels, 1)
        self.end.weight.data.zero_()
        self.end.bias.data.zero_()
        kc = hp.wn_cond_kernel_size
        dims = [cond_in_channels] + [hp.wn_cond_hidden_channels] * (hp.wn_cond_layers - 1) + [2 * C_ * L]
        self.cond_layers = nn.ModuleList([                    # glow.py:74-87
            wn(nn.Conv1d(di, do, kc, padding=(kc - 1) // 2, padding_mode=hp.wn_cond_padding_mode), name="weight")
            for di, do in zip(dims[:-1], dims[1:])])
        self.res_skip_layers = nn.ModuleList()
        dil = hp.wn_dilations_w
        for i in range(L):
            d = 2 ** i if dil is None else (dil if isinstance(dil, int) else dil[i])
            pad = (ks * d - d) // 2 if not getattr(hp, "_explicit_pad", False) else 0
            if not hp.wn_seperable_conv or ks == 1:
                self.in_layers.append(wn(nn.Conv1d(C_, 2 * C_, ks, dilation=d, padding=pad), name="weight"))
            else:                                            # glow.py:115-121
                self.in_layers.append(nn.Sequential(wn(nn.Conv1d(C_, C_, ks, dilation=d, padding=pad, groups=C_), name="weight"),
                                                    wn(nn.Conv1d(C_, 2 * C_, 1), name="weight")))
            if hp.wn_res_skip:
                rs = 2 * C_ if (i < L - 1 and not hp.wn_merge_res_skip) else C_
                self.res_skip_layers.append(wn(nn.Conv1d(C_, rs, 1), name="weight"))


class _Coupling(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.WN = _WN(**kw)


class _InvConv(nn.Conv1d):
    def __init__(self, c):
        super().__init__(c, c, 1, bias=False)
        w = torch.linalg.qr(torch.randn(c, c))[0]
        if torch.det(w) < 0:
            w[:, 0] = -w[:, 0]
        self.weight.data = w.view(c, c, 1).contiguous()


class FlowDecoder(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        hp = hparams

        def need(cond, msg):
            if not cond:
                raise NotImplementedError(f"cookietts_b200.FlowDecoder: {msg}")
        assert hp.n_group % 2 == 0
        need(getattr(hp, "cond_layers", 0) == 0, "decoder-level cond_layers > 0")
        need(not getattr(hp, "cond_res_rezero", False), "cond_res_rezero")
        assert hp.wn_cond_layers > 0, "cond_layers must be greater than 0"      # glow.py:53
        need(hp.wn_cond_kernel_size % 2 == 1, "even wn_cond_kernel_size (the reference's output length then differs from its input's)")
        need(str(hp.wn_cond_act_func).lower() in FD_COND_ACTS, "wn_cond_act_func must be none / lrelu / tanh / sigmoid "
                                                                 "('relu' raises NameError in the reference, glow.py:96)")
        need(hp.wn_cond_padding_mode in FD_PAD_MODES, "wn_cond_padding_mode must be zeros / replicate / reflect / circular")
        self.cond_external = not (hp.wn_cond_layers == 1 and hp.wn_cond_kernel_size == 1 and str(hp.wn_cond_act_func).lower() == "none")
        self.cond_stack = dict(layers=hp.wn_cond_layers, pad=(hp.wn_cond_kernel_size - 1) // 2,
                               pad_mode=FD_PAD_MODES[hp.wn_cond_padding_mode], act=FD_COND_ACTS[str(hp.wn_cond_act_func).lower()])
        need(hp.wn_res_skip or hp.wn_merge_res_skip, "wn_res_skip=False needs wn_merge_res_skip=True (glow.py:53)")
        need(hp.wn_n_layers <= FD_MAX_LAYERS, f"more than {FD_MAX_LAYERS} WN layers")
        self.n_flows, self.n_group = hp.n_flows, hp.n_group
        self.n_early_every, self.n_early_size = hp.n_early_every, hp.n_early_size
        self.n_mel_channels, self.mix_first = hp.n_mel_channels, bool(hp.mix_first)
        # flowtts: encoder_LSTM_dim + speaker_embedding_dim (glow.py:183); untts: cond_input_dim (glow.py:157)
        self.cond_in_channels = (hp.cond_input_dim if hasattr(hp, "cond_input_dim")
                                 else hp.encoder_LSTM_dim + hp.speaker_embedding_dim)
        self.first_pad_value = float(getattr(hp, "decoder_padding_value", 0.0)) if hasattr(hp, "cond_input_dim") else 0.0
        self.wn = dict(n_layers=hp.wn_n_layers, n_channels=hp.wn_n_channels, kernel_size=hp.wn_kernel_size,
                       res_skip=bool(hp.wn_res_skip), merge=bool(hp.wn_merge_res_skip))
        dil = hp.wn_dilations_w
        self.dilations = [2 ** i if dil is None else (dil if isinstance(dil, int) else dil[i]) for i in range(hp.wn_n_layers)]
        self.convinv, self.WN = nn.ModuleList(), nn.ModuleList()
        self.z_split_sizes, self.flow_channels = [], []
        n_rem = hp.n_group
        for k in range(hp.n_flows):
            if k % self.n_early_every == 0 and k > 0:
                n_rem -= self.n_early_size
                self.z_split_sizes.append(self.n_early_size)
            assert n_rem > 0, "n_remaining_channels is 0"
            self.flow_channels.append(n_rem)
            self.convinv.append(_InvConv(n_rem))
            self.WN.append(_Coupling(n_in_channels=n_rem // 2, cond_in_channels=self.cond_in_channels, hp=hp))
        self.z_split_sizes.append(n_rem)
        self._packed, self._packed_key, self._workspace = None, None, None

    def forward(self, *a, **kw):
        raise NotImplementedError("only the inverse pass (inverse / infer) is in scope of this implementation")

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._packed = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def invalidate(self):
        self._packed, self._packed_key = None, None

    def _device(self):
        return self.convinv[0].weight.device

    def _ensure_packed(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and self._packed_key == key:
            return
        dev = self._device()
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        w, L, Cc = self.wn, self.wn["n_layers"], self.wn["n_channels"]
        F, ks = self.n_flows, w["kernel_size"]
        parts = {n: [] for n in FD_WEIGHT_FIELDS}
        for k, n_rem in enumerate(self.flow_channels):
            p = f"WN.{k}.WN."
            parts["start_w"].append(_eff(sd, p + "start")[:, :, 0].reshape(-1))
            parts["start_b"].append(sd[p + "start.bias"].double())
            if not self.cond_external:
                parts["cond_w"].append(_eff(sd, p + "cond_layers.0")[:, :, 0].reshape(-1))
                parts["cond_b"].append(sd[p + "cond_layers.0.bias"].double())
            for i in range(L):
                if p + f"in_layers.{i}.0.bias" in sd:        # separable: W[o, c, j] = P[o, c] * D[c, j], b = P b_d + b_p (exact)
                    dwt, pwt = _eff(sd, p + f"in_layers.{i}.0")[:, 0], _eff(sd, p + f"in_layers.{i}.1")[:, :, 0]
                    parts["in_w"].append((pwt[:, :, None] * dwt[None]).reshape(-1))
                    parts["in_b"].append(pwt @ sd[p + f"in_layers.{i}.0.bias"].double() + sd[p + f"in_layers.{i}.1.bias"].double())
                else:
                    parts["in_w"].append(_eff(sd, p + f"in_layers.{i}").reshape(-1))
                    parts["in_b"].append(sd[p + f"in_layers.{i}.bias"].double())
                if w["res_skip"]:
                    rw = torch.zeros(2 * Cc, Cc, dtype=torch.float64)
                    rb = torch.zeros(2 * Cc, dtype=torch.float64)
                    e = _eff(sd, p + f"res_skip_layers.{i}")[:, :, 0]
                    rw[:e.shape[0]] = e
                    rb[:e.shape[0]] = sd[p + f"res_skip_layers.{i}.bias"].double()
                    parts["rs_w"].append(rw.reshape(-1))
                    parts["rs_b"].append(rb)
            parts["end_w"].append(sd[p + "end.weight"].double()[:, :, 0].reshape(-1))
            parts["end_b"].append(sd[p + "end.bias"].double())
            # the reference inverts in fp32 (modules.py:236-238); fp64 here is closer to the exact inverse
            parts["winv"].append(torch.linalg.inv(sd[f"convinv.{k}.weight"].double()[:, :, 0]).reshape(-1))
        dev_pk, cw = {}, CwgFdWeights()
        for n in FD_WEIGHT_FIELDS:
            if parts[n]:
                dev_pk[n] = torch.cat(parts[n]).float().contiguous().to(dev)
                setattr(cw, n, dev_pk[n].data_ptr())
        cfg = CwgFdConfig(n_group=self.n_group, n_flows=F, n_early_every=self.n_early_every, n_early_size=self.n_early_size,
                          mix_first=int(self.mix_first), cond_channels=self.cond_in_channels, n_layers=L, n_channels=Cc,
                          kernel_size=ks, res_skip=int(w["res_skip"]), merge_res_skip=int(w["merge"]),
                          first_pad_value=self.first_pad_value)
        for i, d in enumerate(self.dilations):
            cfg.dilations[i] = int(d)
        if self.cond_external:                               # per-flow cond stacks, evaluated in `inverse` with cwg_conv1d
            dev_pk["cond_stack"] = [[(_eff(sd, f"WN.{k}.WN.cond_layers.{j}").float().contiguous().to(dev),
                                      sd[f"WN.{k}.WN.cond_layers.{j}.bias"].float().contiguous().to(dev))
                                     for j in range(self.cond_stack["layers"])] for k in range(F)]
        self._packed, self._packed_key, self._cw, self._ccfg = dev_pk, key, cw, cfg

    @torch.no_grad()
    def inverse(self, z, cond, speaker_ids=None):
        """z [B, n_mel, frames] (the latent, already scaled by sigma) , cond [B, cond_channels, T] with
        T = n_mel * frames / n_group -> (mel [B, n_mel, frames], None), like glow.py:302-343."""
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("cookietts_b200.FlowDecoder needs the module on a CUDA device (no CPU fallback)")
        lib = _cabi.load()
        _bind(lib)
        B, n_mel, frames = z.shape
        if n_mel != self.n_mel_channels or (n_mel * frames) % self.n_group:
            raise ValueError(f"z must be [B, {self.n_mel_channels}, frames] with n_mel * frames divisible by n_group")
        T = n_mel * frames // self.n_group
        if tuple(cond.shape) != (B, self.cond_in_channels, T):
            raise ValueError(f"cond must be [B, {self.cond_in_channels}, {T}], got {tuple(cond.shape)}")
        if B == 0 or T == 0:
            return z.new_zeros(z.shape), None
        with torch.cuda.device(dev):
            self._ensure_packed()
            zz = z.to(device=dev, dtype=torch.float32).contiguous().clone().view(B, self.n_group, T)   # z.view(B, n_group, -1)
            cc = cond.to(device=dev, dtype=torch.float32).contiguous()
            nbytes = lib.cwg_fd_workspace_bytes(self._ccfg, B, T)
            if nbytes == 0:
                raise _cabi.CwgError(lib.cwg_last_error().decode())
            if self._workspace is None or self._workspace.numel() < nbytes + 256 or self._workspace.device != dev:
                self._workspace = None
                self._workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            ws_ptr = (self._workspace.data_ptr() + 255) // 256 * 256
            c_all = None
            if self.cond_external:                           # glow.py:145-148: every layer is followed by the activation
                stream = torch.cuda.current_stream(dev).cuda_stream
                cs = self.cond_stack
                c_all = torch.empty(self.n_flows, B, 2 * self.wn["n_channels"] * self.wn["n_layers"], T, device=dev, dtype=torch.float32)
                for k in range(self.n_flows):
                    x = cc
                    layers = self._packed["cond_stack"][k]
                    for j, (wj, bj) in enumerate(layers):
                        last = j == len(layers) - 1
                        y = c_all[k] if last else torch.empty(B, wj.shape[0], T, device=dev, dtype=torch.float32)
                        _cabi.check(lib.cwg_conv1d(x.data_ptr(), B, wj.shape[1], T, wj.data_ptr(), bj.data_ptr(), wj.shape[0],
                                                   wj.shape[2], cs["pad"], cs["pad_mode"], cs["act"], 0.0, 1.0, None, y.data_ptr(), stream))
                        x = y
            self._cw.c_all = c_all.data_ptr() if c_all is not None else None
            _cabi.check(lib.cwg_fd_inverse(self._ccfg, self._cw, cc.data_ptr(), zz.data_ptr(), ws_ptr, nbytes, B, T,
                                           torch.cuda.current_stream(dev).cuda_stream))
        return zz.view(B, self.n_mel_channels, -1), None

    @torch.no_grad()
    def infer(self, cond, speaker_ids=None, sigma=1., *, z: Optional[torch.Tensor] = None):
        """glow.py:345-352: z ~ N(0, sigma^2) [B, n_mel, frames], then `inverse`.  `z` injects a standard-normal latent."""
        B, _, T = cond.shape
        if z is None:
            # the reference draws [B, n_mel, cond frames] (glow.py:347-348), which only fits its own .view when
            # n_group == n_mel; in general the latent that matches cond's T steps has T * n_group / n_mel frames
            z = torch.randn(B, self.n_mel_channels, T * self.n_group // self.n_mel_channels, device=self._device())
        out, _ = self.inverse(z.to(self._device()) * float(sigma), cond, speaker_ids)
        return out

    def launch_count(self) -> int:
        lib = _cabi.load()
        _bind(lib)
        self._ensure_packed()
        return int(lib.cwg_fd_launch_count(self._ccfg))
