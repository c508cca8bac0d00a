"""Drop-in `WaveGlow` module whose `infer` runs on hand-written sm_100a kernels.

Mirrors the interface of the reference class `CookieTTS/_4_mtw/waveglow/glow.py:225-350`:
same constructor kwargs (`WaveGlow(**waveglow_config)`, glow.py:226-227 and `WN_config`
:116-117), same parameter names/shapes (so `load_state_dict(checkpoint['model'])` of a
reference checkpoint works unchanged), same
`infer(spect, speaker_id=None, sigma=1.0)` call.  Only the inverse pass is implemented
(north_star scope): `forward` (the training direction) raises.

There is no CPU or PyTorch fallback: `infer` needs a CUDA device and the in-tree
`libcwg.so`; otherwise it raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _cabi, _torch_ops
from .packing import PackConfig


class Invertible1x1Conv(nn.Module):
    """Parameter holder for glow.py:65-107 (key `convinv.{k}.conv.weight`).  `W_inverse`
    may be set/deleted by callers as in the reference (train.py:332-334); it is not used -
    the inverse is recomputed from `conv.weight` whenever the weights change."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv1d(c, c, kernel_size=1, stride=1, padding=0, bias=False)
        w = torch.linalg.qr(torch.randn(c, c))[0]
        if torch.det(w) < 0:
            w[:, 0] = -w[:, 0]
        self.conv.weight.data = w.view(c, c, 1).contiguous()


class WN(nn.Module):
    """Parameter holder with the layout of glow.py:110-186."""

    def __init__(self, n_in_channels, n_mel_channels, n_layers, n_channels, kernel_size,
                 speaker_embed_dim=0, rezero=False):
        super().__init__()
        assert kernel_size % 2 == 1 and n_channels % 2 == 0
        wn = nn.utils.weight_norm          # legacy API on purpose: yields weight_g / weight_v keys
        self.n_layers, self.n_channels, self.speaker_embed_dim = n_layers, n_channels, speaker_embed_dim
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        if rezero:
            self.alpha_i = nn.ParameterList()
        if speaker_embed_dim:
            self.speaker_embed = nn.Embedding(512, speaker_embed_dim)
            self.speaker_embed.weight.data.mul_(0.05)
        self.start = wn(nn.Conv1d(n_in_channels, n_channels, 1), name="weight")
        self.end = nn.Conv1d(n_channels, 2 * n_in_channels, 1)
        self.end.weight.data.zero_()
        self.end.bias.data.zero_()
        hidden = 256                        # literal in the reference, glow.py:153
        self.cond_layers = nn.ModuleList([
            wn(nn.Conv1d(n_mel_channels + speaker_embed_dim, hidden, 1), name="weight"),
            wn(nn.Conv1d(hidden, hidden, 1), name="weight"),
            wn(nn.Conv1d(hidden, 2 * n_channels * n_layers, 1), name="weight")])
        for i in range(n_layers):
            d = 2 ** i
            self.in_layers.append(wn(nn.Conv1d(n_channels, 2 * n_channels, kernel_size, dilation=d,
                                               padding=(kernel_size * d - d) // 2), name="weight"))
            rs = 2 * n_channels if i < n_layers - 1 else n_channels
            self.res_skip_layers.append(wn(nn.Conv1d(n_channels, rs, 1), name="weight"))
            if rezero:
                self.alpha_i.append(nn.Parameter(torch.rand(1) * 0.02 + 0.09))


class WaveGlow(nn.Module):
    GRAPH_MAX_FRAMES = 2048
    GRAPH_CACHE = 8

    def __init__(self, yoyo=False, yoyo_WN=False, n_mel_channels=80, n_flows=12, n_group=8,
                 n_early_every=4, n_early_size=2, memory_efficient=False, spect_scaling=False,
                 upsample_mode="normal", WN_config=None, win_length=1024, hop_length=256,
                 precision: str = "auto", range_guard: bool = True, graphs="auto"):
        super().__init__()
        if yoyo or yoyo_WN:
            raise ValueError("yoyo models select a different reference class (efficient_model*), not glow.WaveGlow")
        if memory_efficient:
            raise ValueError("memory_efficient=True builds no layers in the reference (glow.py:263-264)")
        if spect_scaling:
            raise ValueError("spect_scaling=True is unusable in the reference (glow.py:232-235,315-316)")
        if upsample_mode != "normal":
            raise ValueError("only upsample_mode='normal' is supported")
        WN_config = dict(WN_config or dict(n_layers=8, n_channels=256, kernel_size=3, speaker_embed_dim=0, rezero=False))
        WN_config.setdefault("speaker_embed_dim", 0)
        WN_config.setdefault("rezero", False)
        assert n_group % 2 == 0
        self.spect_scaling = False
        self.multispeaker = WN_config["speaker_embed_dim"] > 0
        self.n_flows, self.n_group = n_flows, n_group
        self.n_early_every, self.n_early_size = n_early_every, n_early_size
        # "auto": the fastest fp32-accurate mode the kernels have for this width - f16f8 (guarded by the fp16 range check with
        # its bf16x3 fallback) for 256 channels and kernel size 3, bf16x3 for 512 channels, fp32 CUDA cores otherwise
        if precision == "auto":
            c, ks = WN_config["n_channels"], WN_config["kernel_size"]
            precision = "f16f8" if (c == 256 and ks == 3) else ("bf16x3" if (c == 512 and ks == 3) else "ffma")
        self.precision = precision
        self.range_guard = range_guard     # f16f8: check the fp16 range after every infer and fall back to bf16x3
        self.last_status = 0
        self.fallbacks = 0
        # CUDA-graph replay of repeated small shapes (launch-bound regime: ~125 launches per call): "auto" = calls of at
        # most GRAPH_MAX_FRAMES mel frames in total, True = always, False = never
        self.graphs = graphs
        self._graphs = {}
        self._graph_seen = set()
        self.upsample = nn.ConvTranspose1d(n_mel_channels, n_mel_channels, win_length, stride=hop_length)
        self.WN = nn.ModuleList()
        self.convinv = nn.ModuleList()
        self.pack_config = PackConfig(
            n_mel=n_mel_channels, n_flows=n_flows, n_group=n_group, n_early_every=n_early_every,
            n_early_size=n_early_size, win_length=win_length, hop_length=hop_length,
            n_layers=WN_config["n_layers"], n_channels=WN_config["n_channels"],
            kernel_size=WN_config["kernel_size"], cond_hidden=256,
            speaker_embed_dim=WN_config["speaker_embed_dim"], rezero=bool(WN_config["rezero"]))
        self.pack_config.validate()
        for n_rem, n_half in self.pack_config.flow_channels():
            self.convinv.append(Invertible1x1Conv(n_rem))
            self.WN.append(WN(n_half, n_mel_channels * n_group, **WN_config))
        self.n_remaining_channels = self.pack_config.flow_channels()[-1][0]
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._packs = {}

    # ------------------------------------------------------------------ checkpoint compatibility
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts the weight-normed layout (weight_g/weight_v) and the plain `weight` layout
        (after `remove_weight_norm`): a plain weight w becomes v = w, g = ||w||."""
        own = set(self.state_dict().keys())
        sd = {}
        for k, v in state_dict.items():
            if k.endswith(".weight") and k not in own and k[:-7] + ".weight_v" in own:
                norm = v.float().pow(2).sum(dim=tuple(range(1, v.dim())), keepdim=True).sqrt()
                sd[k[:-7] + ".weight_v"] = v
                sd[k[:-7] + ".weight_g"] = norm.to(v.dtype)
            else:
                sd[k] = v
        self._packed, self._packed_key, self._packs = None, None, {}
        return super().load_state_dict(sd, strict=strict, **kw)

    @staticmethod
    def remove_weightnorm(model):
        """The reference version is broken (glow.py:352-360); weight-norm is folded at pack
        time here, so this is a no-op kept for call-site compatibility."""
        return model

    def forward(self, *a, **kw):
        raise NotImplementedError("only the inverse pass (infer) is in scope of this implementation")

    # ------------------------------------------------------------------ packing cache
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _device(self) -> torch.device:
        return self.upsample.weight.device

    def invalidate(self):
        """Drop the packed weights; the next `infer` re-packs from the current parameters.  `load_state_dict` and
        in-place ops on a Parameter bump its version counter and are noticed automatically; edits made through
        `.data` (`weight.data.zero_()`, a reference-code habit) are not - call this after them."""
        self._packed, self._packed_key, self._packs = None, None, {}
    repack = invalidate

    def _ensure_packed(self, precision: Optional[str] = None):
        """Packs the checkpoint on the device (torch.ops.cookietts_b200.waveglow_pack -> cwg_pack_weights, fp64).
        One blob per precision mode in use (the module's own, plus bf16x3 once the f16f8 range guard needs it)."""
        precision = precision or self.precision
        key = self._weights_key()
        if self._packed_key != key:
            self._packs, self._packed_key, self._graphs = {}, key, {}
        if precision not in self._packs:
            ops, lib = _torch_ops.load(), _cabi.load()
            mode = _cabi.MODES[precision]
            sd = self.state_dict()
            names = list(sd.keys())
            self._cfg_list = _torch_ops.config_list(self.pack_config)
            self._embed_dim = self.pack_config.speaker_embed_dim
            self._n_speakers = int(self.WN[0].speaker_embed.weight.shape[0]) if self.multispeaker else 0
            blob = ops.waveglow_pack([sd[k] for k in names], "\n".join(names), self._cfg_list, mode)
            self._ccfg = _cabi.make_config(self.pack_config)
            w = _cabi.CwgWeights()      # ctypes view of the same blob for the stage-level entry points (tests, tools)
            base = (blob.data_ptr() + 255) // 256 * 256
            _cabi.check(lib.cwg_packed_view(self._ccfg, mode, self._embed_dim, self._n_speakers, base, blob.numel() - 256, w))
            self._packs[precision] = (blob, w)
        if precision == self.precision:
            self._packed, self._cw = self._packs[precision]

    def packed_views(self) -> Dict[str, torch.Tensor]:
        """name -> tensor views into the packed blob (the arrays of cwg_weights this mode uses), for tests and tools."""
        self._ensure_packed()
        pc, blob, w = self.pack_config, self._packed, self._cw
        F, L, C, H = pc.n_flows, pc.n_layers, pc.n_channels, pc.cond_hidden
        kcp = -(-(pc.taps * pc.n_mel) // 64) * 64
        E, S = max(self._embed_dim, 1), self._n_speakers
        mg = 16 if pc.n_group <= 16 else 32                  # include/cwg.h CWG_GROUP_PAD
        shapes = {"cond_w": (F, pc.phases * H, kcp), "w1": (F, L, 2 * C, pc.k1), "w2": (F, L, C + mg, C)}
        small = {"b1": (F, L, 2 * C), "b2": (F, L, C), "eo_b": (F, mg), "start_w": (F, C, mg // 2),
                 "start_b": (F, C), "winv": (F, mg, mg), "cond_b_base": (F, H), "cond_w_spk": (F, H, E)}
        if self._embed_dim:
            small["spk_embed"] = (F, S, self._embed_dim)
        f16 = torch.float16 if self.precision == "f16f8" else torch.bfloat16

        def at(ptr, shape, dtype):
            n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
            off = ptr - blob.data_ptr()
            return blob[off:off + n].view(dtype).view(shape)
        out = {}
        for name, shape in shapes.items():
            for suffix, dtype in (("_f32", torch.float32), ("_hi", f16), ("_lo", f16), ("_h8", torch.float8_e5m2), ("_l8", torch.float8_e5m2)):
                ptr = getattr(w, name + suffix, None)
                if ptr:
                    out[name + suffix] = at(ptr, shape, dtype)
        for name, shape in small.items():
            out[name] = at(getattr(w, name), shape, torch.float32)
        for name in ("w0_hi", "w0_lo"):
            if getattr(w, name):
                out[name] = at(getattr(w, name), (F, 2 * C, 48), f16)
        return out

    def _cond_bias(self, batch: int, speaker_ids) -> torch.Tensor:
        """[B, F, H] folded cond-chain bias (cwg_cond_bias; carries the speaker-embedding branch, glow.py:193-196)."""
        lib = _cabi.load()
        dev = self._device()
        out = torch.empty(batch, self.pack_config.n_flows, self.pack_config.cond_hidden, device=dev)
        ids = None
        if self.multispeaker and speaker_ids is not None:
            ids = torch.as_tensor(speaker_ids).reshape(-1).to(device=dev, dtype=torch.long)
            if ids.numel() == 1 and batch > 1:
                ids = ids.expand(batch)
            ids = ids.contiguous()
        _cabi.check(lib.cwg_cond_bias(self._ccfg, self._cw, ids.data_ptr() if ids is not None else None, batch,
                                      out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return out

    def draw_z(self, batch: int, t_mel: int, generator=None) -> torch.Tensor:
        """Standard-normal latent [B, T] drawn in the order of the reference's draws
        (main latent, then the early outputs at descending k; glow.py:326,342-347)."""
        pc, dev = self.pack_config, self._device()
        tp = t_mel * pc.phases
        z = torch.empty(batch, tp, pc.n_group, device=dev)
        n_rem_last = self.n_remaining_channels
        z[:, :, pc.n_group - n_rem_last:] = torch.randn(batch, n_rem_last, tp, device=dev, generator=generator).transpose(1, 2)
        early = [k for k in range(pc.n_flows) if k % pc.n_early_every == 0 and k > 0]
        off = pc.n_group - n_rem_last
        for _ in sorted(early, reverse=True):
            off -= pc.n_early_size
            z[:, :, off:off + pc.n_early_size] = torch.randn(batch, pc.n_early_size, tp, device=dev, generator=generator).transpose(1, 2)
        return z.view(batch, tp * pc.n_group)

    # ------------------------------------------------------------------ the hot path
    @torch.no_grad()
    def infer(self, spect, speaker_id=None, sigma=1.0, *, speaker_ids=None, z=None, layer_events=None):
        """mel [B, n_mel, T_mel] -> audio [B, T_mel*hop] (fp32, on the module's device).

        `speaker_id` is the reference keyword (glow.py:314); `speaker_ids` is what
        `Denoiser`/notebooks pass (denoiser.py:39).  `z` ([B, T], standard normal) injects the
        latent the reference draws internally; None draws it here.  `layer_events` (optional,
        (begin, end) lists of torch.cuda.Event) are recorded around each WN-layer launch."""
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("cookietts_b200.WaveGlow.infer needs the module on a CUDA device (no CPU fallback)")
        if self.precision not in _cabi.MODES:
            raise ValueError(f"precision must be one of {list(_cabi.MODES)}")
        mode = _cabi.MODES[self.precision]
        if speaker_id is None:
            speaker_id = speaker_ids
        if self.multispeaker and speaker_id is None:
            # the reference fails on a shape mismatch in cond_layers[0] here (glow.py:193-199)
            raise ValueError("this model has speaker embeddings: pass speaker_id / speaker_ids")
        if self.multispeaker:
            speaker_id = torch.as_tensor(speaker_id).reshape(-1).cpu()          # host copy: the range check costs no kernel
            if speaker_id.numel() not in (1, spect.shape[0]):
                # glow.py:195 concatenates the embedding onto the mel along channels: one id per utterance
                raise ValueError(f"speaker_id must hold 1 or batch = {spect.shape[0]} entries, got {speaker_id.numel()}")
            n_spk = self.WN[0].speaker_embed.weight.shape[0]
            if int(speaker_id.min()) < 0 or int(speaker_id.max()) >= n_spk:
                raise IndexError(f"speaker_id out of range for an embedding table of {n_spk} rows")   # nn.Embedding raises too
            if speaker_id.numel() == 1 and spect.shape[0] > 1:
                speaker_id = speaker_id.expand(spect.shape[0])
            speaker_id = speaker_id.long().contiguous()
        else:
            speaker_id = None
        pc = self.pack_config
        if spect.dim() != 3 or spect.shape[1] != pc.n_mel:
            raise ValueError(f"spect must be [B, {pc.n_mel}, T_mel], got {tuple(spect.shape)}")
        mel = spect.to(device=dev, dtype=torch.float32).contiguous()
        batch, _, t_mel = mel.shape
        if t_mel == 0 or batch == 0:
            return torch.zeros(batch, t_mel * pc.hop_length, device=dev)
        with torch.cuda.device(dev):
            self._ensure_packed()
            ops = _torch_ops.load()
            T = t_mel * pc.hop_length
            if z is None:
                z = self.draw_z(batch, t_mel)
            z = z.to(device=dev, dtype=torch.float32).contiguous()
            if tuple(z.shape) != (batch, T):
                raise ValueError(f"z must be [B, T_mel*hop] = {(batch, T)}, got {tuple(z.shape)}")
            ev_b, ev_e = [], []
            if layer_events is not None:
                begin, end = layer_events
                for e in list(begin) + list(end):          # torch creates the handle lazily
                    if not e.cuda_event:
                        e.record(torch.cuda.current_stream(dev))
                ev_b, ev_e = [int(e.cuda_event) for e in begin], [int(e.cuda_event) for e in end]
            use_graph = (layer_events is None and not torch.cuda.is_current_stream_capturing() and
                         (self.graphs is True or (self.graphs == "auto" and batch * t_mel <= self.GRAPH_MAX_FRAMES)))
            if use_graph and self.graphs == "auto":
                # "auto" captures a shape the second time it is seen: a stream of ever-new lengths (a TTS server) would pay
                # a warm-up run and a capture per call for nothing
                gkey = (tuple(mel.shape), float(sigma), mode, speaker_id is not None)
                if gkey not in self._graphs and gkey not in self._graph_seen:
                    if len(self._graph_seen) > 64:
                        self._graph_seen.clear()
                    self._graph_seen.add(gkey)
                    use_graph = False
            if use_graph:
                audio, status = self._graph_infer(ops, mode, mel, speaker_id, z, float(sigma))
            else:
                audio, status = ops.waveglow_infer(self._packed, self._cfg_list, mode, self._embed_dim, self._n_speakers, mel,
                                                   speaker_id, z, float(sigma), ev_b, ev_e)
            if self.precision == "f16f8" and self.range_guard and not torch.cuda.is_current_stream_capturing():
                # fp16 hi planes: values beyond +-65504 (or a NaN / Inf waveform) are flagged on the device (cwg_infer_status);
                # reading the flag synchronises this call.  On a hit the call is repeated in bf16x3 (fp32 exponent range).
                self.last_status = int(status.item())
                if self.last_status:
                    import warnings
                    warnings.warn(f"cookietts_b200.WaveGlow: f16f8 range guard tripped (status {self.last_status}); "
                                  "re-running this call in bf16x3", RuntimeWarning)
                    self.fallbacks += 1
                    self._ensure_packed("bf16x3")
                    audio, _ = ops.waveglow_infer(self._packs["bf16x3"][0], self._cfg_list, _cabi.MODES["bf16x3"], self._embed_dim,
                                                  self._n_speakers, mel, speaker_id, z, float(sigma), [], [])
        return audio

    def _graph_infer(self, ops, mode, mel, speaker_id, z, sigma):
        """Replays the captured launch sequence of this (shape, sigma) on static input buffers; captures it on first use."""
        dev = mel.device
        if speaker_id is not None:
            speaker_id = speaker_id.to(dev)
        key = (tuple(mel.shape), sigma, mode, speaker_id is not None)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.GRAPH_CACHE:
                self._graphs.pop(next(iter(self._graphs)))
            s_mel, s_z = mel.clone(), z.clone()
            s_spk = speaker_id.clone() if speaker_id is not None else None

            def run():
                return ops.waveglow_infer(self._packed, self._cfg_list, mode, self._embed_dim, self._n_speakers, s_mel, s_spk,
                                          s_z, sigma, [], [])
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                run()                                   # warm-up outside the capture (tensor-map cache, lazy module loads)
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = run()
            ent = (g, s_mel, s_z, s_spk, out)
            self._graphs[key] = ent
        else:
            ent[1].copy_(mel); ent[2].copy_(z)
            if speaker_id is not None:
                ent[3].copy_(speaker_id)
        ent[0].replay()
        return ent[4][0].clone(), ent[4][1]

    def launch_count(self) -> int:
        """Kernels launched by one `infer` call in the current precision mode."""
        lib = _cabi.load()
        # cwg_infer's kernels + the k_cond_bias launch of cwg_cond_bias that waveglow_infer issues before it
        return int(lib.cwg_launch_count(_cabi.make_config(self.pack_config), _cabi.MODES[self.precision])) + 1
