"""T2S-server vocoder stage: WAV container on CPU; the trim / pad / int16 kernel on the GPU against
numpy's `(audio * 2**15).astype('int16')` (text2speech.py:694)."""
import io

import numpy as np
import pytest
import torch

from cookietts_b200 import serving


def test_wav_bytes_match_scipy():
    from scipy.io import wavfile
    rs = np.random.RandomState(0)
    for n in (0, 1, 7, 22050):
        pcm = rs.randint(-32768, 32768, size=n).astype(np.int16)
        buf = io.BytesIO()
        wavfile.write(buf, 22050, pcm)
        assert serving.wav_bytes(22050, pcm) == buf.getvalue()


def test_pcm16_refuses_cpu():
    with pytest.raises(RuntimeError):
        serving.pcm16(torch.zeros(1, 8))


def _ref_item(audio_row, n, pad):
    a = np.pad(audio_row[:n], (0, pad))                                # text2speech.py:675,689-691
    return (a.astype(np.float32) * 2 ** 15).astype("int16")            # text2speech.py:694


@pytest.mark.gpu
@pytest.mark.parametrize("T,pad", [(4096, 0), (4096, 800), (1003, 37), (8, 8)])
def test_pcm16_matches_numpy(T, pad):
    rs = np.random.RandomState(T + pad)
    B = 5
    audio = (rs.uniform(-0.999, 0.999, size=(B, T))).astype(np.float32)
    audio[0, :6] = [0.0, -0.0, 1e-9, -1e-9, 0.5 / 32768, -0.5 / 32768]
    audio[1, :3] = [32767.0 / 32768, -1.0, 0.99999]
    n_valid = [T, T // 2, 1, 0, max(T - 3, 0)]
    out = serving.pcm16(torch.from_numpy(audio).cuda(), n_valid, pad).cpu().numpy()
    assert out.dtype == np.int16 and out.shape == (B, T + pad)
    for b in range(B):
        np.testing.assert_array_equal(out[b, :n_valid[b] + pad], _ref_item(audio[b], n_valid[b], pad))
        assert not out[b, n_valid[b]:].any()


@pytest.mark.gpu
def test_pcm16_out_of_range_modes():
    v = np.array([[1.0, 1.5, -1.5, 3.0, np.nan, np.inf, -np.inf, 7e4]], np.float32)
    wrap = serving.pcm16(torch.from_numpy(v).cuda()).cpu().numpy()[0]
    sat = serving.pcm16(torch.from_numpy(v).cuda(), saturate=True).cpu().numpy()[0]
    s = (v[0].astype(np.float64) * 32768)
    expect_wrap = [int(np.int64(np.trunc(x)) & 0xffff) if np.isfinite(x) and abs(x) < 2 ** 31 else 0 for x in s]
    expect_wrap = np.array(expect_wrap, np.uint16).view(np.int16)
    np.testing.assert_array_equal(wrap, expect_wrap)
    np.testing.assert_array_equal(sat, [32767, 32767, -32768, 32767, 0, 32767, -32768, 32767])


class _ToneVocoder(torch.nn.Module):
    """Deterministic stand-in vocoder: [B, n_mel, T_mel] -> [B, T_mel*hop]."""

    def __init__(self, hop):
        super().__init__()
        self.hop = hop

    def infer(self, mel, sigma=1.0):
        B, _, Tm = mel.shape
        t = torch.arange(Tm * self.hop, device=mel.device, dtype=torch.float32)
        return 0.5 * torch.sin(t[None] * 0.01 * mel[:, 0, :1]) * mel[:, 1, :1] * sigma


@pytest.mark.gpu
def test_vocode_pcm16_slices_and_trims(tmp_path):
    hop, sr = 64, 8000
    mels = torch.ones(5, 4, 20, device="cuda")
    mels[:, 0] = torch.arange(1, 6, device="cuda")[:, None]          # per-utterance tone
    mels[:, 1] = torch.linspace(0.5, 1.5, 5, device="cuda")[:, None]
    lengths = [20, 3, 11, 0, 19]
    voc = _ToneVocoder(hop)
    got = serving.vocode_pcm16(voc, mels, lengths, hop, sr, vocoder_batch_size=2, cat_silence_s=0.01, sigma=0.5)
    full = voc.infer(mels, sigma=0.5).cpu().numpy()
    pad = int(0.01 * sr)
    assert len(got) == 5
    for j, n in enumerate(lengths):
        np.testing.assert_array_equal(got[j], _ref_item(full[j], n * hop, pad))
    serving.write_wav(str(tmp_path / "a.wav"), sr, got[0])
    from scipy.io import wavfile
    rate, data = wavfile.read(str(tmp_path / "a.wav"))
    assert rate == sr
    np.testing.assert_array_equal(data, got[0])
