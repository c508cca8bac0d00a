"""CPU oracle for the reference Denoiser (the WaveGlow call sites' post-filter).

TEST INFRASTRUCTURE ONLY (same rules as waveglow_oracle.py).  numpy restatement of
  CookieTTS/utils/audio/stft.py:44-151           STFT.__init__ / transform / inverse
  CookieTTS/utils/audio/audio_processing.py:7-57  window_sumsquare
  CookieTTS/_4_mtw/waveglow/denoiser.py:59-71     Denoiser.forward
Parity status: PINNED by oracle/make_golden_denoiser.py (reference STFT + Denoiser.forward run here
with librosa's two helper functions `pad_center` / `tiny` stubbed - librosa is not installed - and
the results stored in tests/golden/denoiser_*.npz).
"""
from __future__ import annotations

import numpy as np
from scipy.signal import get_window


def pad_center(data: np.ndarray, size: int) -> np.ndarray:
    lpad = (size - data.shape[-1]) // 2
    return np.pad(data, (lpad, size - data.shape[-1] - lpad))


class StftOracle:
    """stft.py:44-78: windowed Fourier bases (forward [2*cutoff, fl], inverse via pinv)."""

    def __init__(self, filter_length: int, hop_length: int, win_length: int, window: str = "hann"):
        self.fl, self.hop, self.win, self.window = filter_length, hop_length, win_length, window
        scale = filter_length / hop_length
        fb = np.fft.fft(np.eye(filter_length))
        self.cutoff = int(filter_length / 2 + 1)
        fb = np.vstack([np.real(fb[:self.cutoff]), np.imag(fb[:self.cutoff])])
        fwd = fb.astype(np.float32)                                   # torch.FloatTensor(...)
        inv = np.linalg.pinv(scale * fb).T.astype(np.float32)
        w = pad_center(get_window(window, win_length, fftbins=True), filter_length).astype(np.float32)
        self.forward_basis = fwd * w[None, :]                         # [2*cutoff, fl]
        self.inverse_basis = inv * w[None, :]                         # [2*cutoff, fl]

    def n_frames(self, n_samples: int) -> int:
        return (n_samples + 2 * (self.fl // 2) - self.fl) // self.hop + 1

    def transform_ri(self, x: np.ndarray, dtype=np.float64):
        """stft.py:79-111 up to the real/imag parts: x [B, T] -> (re, im) each [B, cutoff, frames]."""
        B, T = x.shape
        p = self.fl // 2
        xp = np.pad(x.astype(dtype), ((0, 0), (p, p)), mode="reflect")
        nf = self.n_frames(T)
        idx = np.arange(nf)[:, None] * self.hop + np.arange(self.fl)[None, :]
        frames = xp[:, idx]                                            # [B, nf, fl]
        out = np.einsum("bfk,nk->bnf", frames, self.forward_basis.astype(dtype), optimize=True)
        return out[:, :self.cutoff], out[:, self.cutoff:]

    def transform(self, x, dtype=np.float64):
        re, im = self.transform_ri(x, dtype)
        return np.sqrt(re ** 2 + im ** 2), np.arctan2(im, re)

    def window_sumsquare(self, n_frames: int) -> np.ndarray:
        """audio_processing.py:7-57 (norm=None)."""
        n = self.fl + self.hop * (n_frames - 1)
        x = np.zeros(n, dtype=np.float32)
        win_sq = pad_center(get_window(self.window, self.win, fftbins=True) ** 2, self.fl)
        for i in range(n_frames):
            s = i * self.hop
            x[s:min(n, s + self.fl)] += win_sq[:max(0, min(self.fl, n - s))]
        return x

    def inverse_ri(self, re: np.ndarray, im: np.ndarray, dtype=np.float64) -> np.ndarray:
        """stft.py:117-146 from the recombined real/imag parts: -> [B, 1, T_out]."""
        B, _, nf = re.shape
        rec = np.concatenate([re, im], axis=1).astype(dtype)          # [B, 2*cutoff, nf]
        n = self.fl + self.hop * (nf - 1)
        y = np.zeros((B, n), dtype)
        contrib = np.einsum("bnf,nk->bfk", rec, self.inverse_basis.astype(dtype), optimize=True)
        for f in range(nf):                                           # conv_transpose1d, stride hop
            y[:, f * self.hop:f * self.hop + self.fl] += contrib[:, f]
        ws = self.window_sumsquare(nf)
        nz = ws > np.finfo(np.float32).tiny
        y[:, nz] /= ws[nz].astype(dtype)
        y *= self.fl / self.hop
        p = self.fl // 2
        return y[:, None, p:n - p]


def denoise(stft: StftOracle, audio: np.ndarray, bias_spec: np.ndarray, strength: float, dtype=np.float64) -> np.ndarray:
    """Denoiser.forward (denoiser.py:59-71): audio [B, T], bias_spec [1 or B, cutoff, 1] -> [B, 1, T_out].
    mag*cos(phase) = re * (mag_new/mag), so the phase round trip is evaluated as a per-bin scale."""
    mag, ph = stft.transform(audio, dtype)
    new = np.maximum(mag - bias_spec.astype(dtype) * strength, 0.0)
    return stft.inverse_ri(new * np.cos(ph), new * np.sin(ph), dtype)


def bias_spectrum(stft: StftOracle, bias_audio: np.ndarray, dtype=np.float64) -> np.ndarray:
    """denoiser.py:50-57: mean magnitude over frames of the vocoder's output for a near-silent mel."""
    mag, _ = stft.transform(bias_audio, dtype)
    return mag.mean(axis=2, keepdims=True)
