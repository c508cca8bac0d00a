"""Drop-in `Denoiser` (the post-filter every reference WaveGlow call site applies after `infer`).

Mirrors `CookieTTS/_4_mtw/waveglow/denoiser.py:7-71`: same constructor arguments, the same
`bias_spec` buffer ([n_speakers or 1, cutoff, 1]) and `forward(wg_audio, speaker_ids=None,
strength=0.1)` returning `[B, 1, T_out]`.  The reference does the STFT round trip on the CPU
(`stft_device='cpu'`) with the conv-based STFT of `CookieTTS/utils/audio/stft.py:44-151`; here it stays on
the GPU behind the C ABI (`cwg_denoise`, `cwg_stft_mean_magnitude` in include/cwg.h).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _cabi


def _hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) (stft.py:67)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def _pad_center(x: np.ndarray, size: int) -> np.ndarray:
    lpad = (size - x.shape[-1]) // 2
    return np.pad(x, (lpad, size - x.shape[-1] - lpad))


def stft_bases(filter_length: int, hop_length: int, win_length: int):
    """Windowed forward basis [2*cutoff, fl] and inverse (pinv) basis [2*cutoff, fl] of stft.py:44-78 (fp32)."""
    assert filter_length >= win_length
    k = np.arange(filter_length)
    cutoff = filter_length // 2 + 1
    ang = -2.0 * np.pi * np.outer(np.arange(cutoff), k) / filter_length
    fb = np.vstack([np.cos(ang), np.sin(ang)])
    w = _pad_center(_hann_periodic(win_length), filter_length).astype(np.float32)
    fwd = fb.astype(np.float32) * w[None, :]
    inv = np.linalg.pinv(filter_length / hop_length * fb).T.astype(np.float32) * w[None, :]
    return fwd, inv


def window_sumsquare(n_frames: int, filter_length: int, hop_length: int, win_length: int) -> np.ndarray:
    """audio_processing.py:7-57 with norm=None, dtype float32: the squared window overlap-added per frame."""
    n = filter_length + hop_length * (n_frames - 1)
    win_sq = _pad_center(_hann_periodic(win_length) ** 2, filter_length)
    x = np.zeros(n, dtype=np.float32)
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + filter_length)] += win_sq[:max(0, min(filter_length, n - s))].astype(np.float32)
    return x


class Denoiser(nn.Module):
    """Removes model bias from audio produced with waveglow (denoiser.py:7-71), on the GPU."""

    def __init__(self, waveglow, sampling_rate=48000, filter_length=None, hop_length=None, win_length=None,
                 n_mel_channels=160, n_frames=20, mu=0, var=0.01, wg_sigma=0.01, stft_device=None,
                 speaker_dependant=False, speaker_id=0):
        super().__init__()
        filter_length = filter_length or sampling_rate // 40
        win_length = win_length or sampling_rate // 40
        hop_length = hop_length or sampling_rate // 400
        self.filter_length, self.hop_length, self.win_length = filter_length, hop_length, win_length
        self.cutoff = filter_length // 2 + 1
        p = next(waveglow.parameters())
        dev = p.device
        # `stft_device` is accepted for call-site compatibility (denoiser.py:12); the transform always runs on
        # the vocoder's CUDA device here.
        self.stft_device = dev
        fwd, inv = stft_bases(filter_length, hop_length, win_length)
        self.register_buffer("forward_basis", torch.from_numpy(fwd).to(dev), persistent=False)
        self.register_buffer("inverse_basis_t", torch.from_numpy(np.ascontiguousarray(inv.T)).to(dev), persistent=False)
        self._window_sums = {}
        self._workspace = None

        mel_input = torch.randn((1, n_mel_channels, n_frames), dtype=p.dtype, device=dev) * float(var) + float(mu)
        with torch.no_grad():
            if speaker_dependant:                                       # denoiser.py:31-45
                if hasattr(waveglow, "speaker_embed"):
                    n_speakers = waveglow.speaker_embed.num_embeddings
                elif hasattr(waveglow, "WN") and hasattr(getattr(waveglow.WN[0], "WN", None), "speaker_embed"):
                    n_speakers = waveglow.WN[0].WN.speaker_embed.num_embeddings
                else:
                    n_speakers = 1
                rows = []
                for s in range(n_speakers):
                    ids = torch.tensor([s], device=dev, dtype=torch.int64)
                    rows.append(self._checked(waveglow.infer(mel_input, speaker_ids=ids, sigma=wg_sigma)))
                bias_audio = torch.cat(rows, dim=0)
            else:                                                       # denoiser.py:46-52
                ids = torch.tensor([speaker_id], device=dev, dtype=torch.int64)
                bias_audio = self._checked(waveglow.infer(mel_input, speaker_ids=ids, sigma=wg_sigma))
            bias_spec = self.mean_magnitude(bias_audio)                # [n, cutoff]
            assert torch.isfinite(bias_spec).all(), "Inf/NaN elements found in bias_spec"
        self.register_buffer("bias_spec", bias_spec.unsqueeze(2))      # [n_speakers, cutoff, 1]

    @staticmethod
    def _checked(audio):
        audio = audio.to(dtype=torch.float)
        assert not torch.isinf(audio).any(), "Inf elements found in Vocoder Output"
        assert not torch.isnan(audio).any(), "NaN elements found in Vocoder Output"
        return audio

    # ------------------------------------------------------------------ helpers
    def _ws(self, nbytes: int, dev) -> int:
        if self._workspace is None or self._workspace.numel() < nbytes + 1024 or self._workspace.device != dev:
            self._workspace = None
            self._workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        return (self._workspace.data_ptr() + 1023) // 1024 * 1024

    def _prep(self, audio):
        dev = self.forward_basis.device
        if dev.type != "cuda":
            raise RuntimeError("cookietts_b200.Denoiser needs a CUDA device (no CPU fallback)")
        if audio.dim() == 3 and audio.shape[1] == 1:
            audio = audio[:, 0]
        if audio.dim() != 2:
            raise ValueError(f"audio must be [B, T], got {tuple(audio.shape)}")
        return audio.to(device=dev, dtype=torch.float32).contiguous(), dev

    @torch.no_grad()
    def mean_magnitude(self, audio) -> torch.Tensor:
        """[B, T] -> mean over frames of the STFT magnitude, [B, cutoff] (denoiser.py:50-57)."""
        lib = _cabi.load()
        audio, dev = self._prep(audio)
        B, T = audio.shape
        with torch.cuda.device(dev):
            nbytes = lib.cwg_denoise_workspace_bytes(B, T, self.filter_length, self.hop_length)
            if nbytes == 0:
                raise _cabi.CwgError(lib.cwg_last_error().decode())
            ws = self._ws(nbytes, dev)
            out = torch.empty(B, self.cutoff, device=dev, dtype=torch.float32)
            _cabi.check(lib.cwg_stft_mean_magnitude(audio.data_ptr(), B, T, self.filter_length, self.hop_length,
                                                    self.forward_basis.data_ptr(), out.data_ptr(), ws, nbytes,
                                                    torch.cuda.current_stream(dev).cuda_stream))
        return out

    @torch.no_grad()
    def forward(self, wg_audio, speaker_ids=None, strength=0.1):
        """[B, T] vocoder output -> [B, 1, T_out] with the bias spectrum subtracted (denoiser.py:59-71)."""
        lib = _cabi.load()
        audio, dev = self._prep(wg_audio)
        B, T = audio.shape
        fl, hop = self.filter_length, self.hop_length
        with torch.cuda.device(dev):
            nbytes = lib.cwg_denoise_workspace_bytes(B, T, fl, hop)
            if nbytes == 0:
                raise _cabi.CwgError(lib.cwg_last_error().decode())
            t_out = lib.cwg_denoise_out_samples(T, fl, hop)
            nf = (T + 2 * (fl // 2) - fl) // hop + 1
            if nf not in self._window_sums:
                self._window_sums[nf] = torch.from_numpy(window_sumsquare(nf, fl, hop, self.win_length)).to(dev)
            wsum = self._window_sums[nf]
            idx_ptr = None
            if speaker_ids is not None and self.bias_spec.shape[0] > 1:  # denoiser.py:64-67
                idx = torch.as_tensor(speaker_ids, device=dev).to(torch.int32).contiguous()
                if idx.numel() != B:
                    raise ValueError("speaker_ids must have one entry per utterance")
                if int(idx.min()) < 0 or int(idx.max()) >= self.bias_spec.shape[0]:
                    raise IndexError(f"speaker id out of range for a bias spectrum of {self.bias_spec.shape[0]} rows")
                idx_ptr = idx.data_ptr()
            elif self.bias_spec.shape[0] not in (1, B):
                raise ValueError("bias_spec has one row per speaker: pass speaker_ids")
            elif self.bias_spec.shape[0] == B and B > 1:                 # broadcasting of the reference's subtraction
                idx = torch.arange(B, device=dev, dtype=torch.int32)
                idx_ptr = idx.data_ptr()
            ws = self._ws(nbytes, dev)
            out = torch.empty(B, 1, t_out, device=dev, dtype=torch.float32)
            bias = self.bias_spec.to(device=dev, dtype=torch.float32).contiguous()
            _cabi.check(lib.cwg_denoise(audio.data_ptr(), B, T, fl, hop, self.forward_basis.data_ptr(),
                                        self.inverse_basis_t.data_ptr(), wsum.data_ptr(), bias.data_ptr(),
                                        C.c_void_p(idx_ptr), float(strength), out.data_ptr(), ws, nbytes,
                                        torch.cuda.current_stream(dev).cuda_stream))
        return out
