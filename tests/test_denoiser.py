"""`cookietts_b200.Denoiser`: host-side bases against the oracle on CPU; the CUDA filter against the
golden vectors of the unmodified reference (tests/golden/denoiser_*.npz, oracle/make_golden_denoiser.py)."""
import os

import numpy as np
import pytest
import torch

from cookietts_b200 import _cabi
from cookietts_b200.denoiser import Denoiser, stft_bases, window_sumsquare
from oracle.denoiser_oracle import StftOracle, denoise
from tests.helpers import GOLDEN_DIR, max_abs

CASES = ["denoiser_22k", "denoiser_48k", "denoiser_small"]


@pytest.mark.parametrize("fl,hop,win", [(551, 55, 551), (1200, 120, 1200), (100, 10, 100), (64, 16, 48)])
def test_host_bases_match_oracle(fl, hop, win):
    st = StftOracle(fl, hop, win)
    fwd, inv = stft_bases(fl, hop, win)
    assert max_abs(fwd, st.forward_basis) < 1e-6
    assert max_abs(inv, st.inverse_basis) < 1e-6
    for nf in (1, 7, 40):
        assert max_abs(window_sumsquare(nf, fl, hop, win), st.window_sumsquare(nf)) < 1e-6


def test_denoise_shape_helpers_cpu():
    lib = _cabi.load()
    for T, fl, hop in [(6000, 551, 55), (9000, 1200, 120), (1234, 100, 10)]:
        nf = (T + 2 * (fl // 2) - fl) // hop + 1
        assert lib.cwg_denoise_out_samples(T, fl, hop) == fl + hop * (nf - 1) - 2 * (fl // 2)
        assert lib.cwg_denoise_workspace_bytes(2, T, fl, hop) >= 2 * nf * (2 * (fl // 2 + 1) + fl) * 4
    assert lib.cwg_denoise_workspace_bytes(1, 10, 800, 200) == 0        # reflect pad needs T > fl/2
    assert lib.cwg_denoise(None, 1, 4000, 800, 200, None, None, None, None, None, 0.1, None, None, 0, None) != 0
    assert b"NULL" in lib.cwg_last_error()


class _FixedVocoder(torch.nn.Module):
    """Stand-in with the `infer(mel, speaker_ids=, sigma=)` signature the Denoiser constructor calls (denoiser.py:39-50)."""

    def __init__(self, bias_audio, dev):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1, device=dev))
        self.bias_audio = bias_audio.to(dev)

    def infer(self, mel, speaker_ids=None, sigma=1.0):
        i = int(speaker_ids[0]) % self.bias_audio.shape[0]
        return self.bias_audio[i:i + 1].clone()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_denoiser_matches_reference(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    sr = int(g["sampling_rate"])
    dev = torch.device("cuda:0")
    den = Denoiser(_FixedVocoder(torch.from_numpy(g["bias_audio"]), dev), sampling_rate=sr, n_mel_channels=80)
    assert tuple(den.bias_spec.shape) == tuple(g["bias_spec"].shape)
    assert max_abs(den.bias_spec.cpu().numpy(), g["bias_spec"]) < 1e-5
    out = den(torch.from_numpy(g["audio"]).to(dev), strength=float(g["strength"]))
    assert tuple(out.shape) == tuple(g["denoised"].shape)
    assert max_abs(out.cpu().numpy(), g["denoised"]) < 5e-5                # fp32 vs the reference's fp32
    st = StftOracle(sr // 40, sr // 400, sr // 40)
    ref64 = denoise(st, g["audio"], g["bias_spec"], float(g["strength"]))
    assert max_abs(out.cpu().numpy(), ref64) < 5e-5


@pytest.mark.gpu
def test_denoiser_speaker_dependant_and_zero_strength():
    rs = np.random.RandomState(5)
    dev = torch.device("cuda:0")
    sr = 8000
    bias_audio = torch.from_numpy((rs.standard_normal((3, 4000)) * 0.02).astype(np.float32))

    class Spk(_FixedVocoder):
        def __init__(self, *a):
            super().__init__(*a)
            self.speaker_embed = torch.nn.Embedding(3, 4)

    den = Denoiser(Spk(bias_audio, dev), sampling_rate=sr, n_mel_channels=80, speaker_dependant=True)
    assert tuple(den.bias_spec.shape) == (3, sr // 40 // 2 + 1, 1)
    st = StftOracle(sr // 40, sr // 400, sr // 40)
    audio = (rs.standard_normal((4, 5000)) * 0.2).astype(np.float32)
    ids = np.array([2, 0, 1, 2])
    out = den(torch.from_numpy(audio).to(dev), speaker_ids=torch.from_numpy(ids).to(dev), strength=0.7).cpu().numpy()
    bias = den.bias_spec.cpu().numpy()
    ref = denoise(st, audio, bias[ids], 0.7)
    assert max_abs(out, ref) < 5e-5
    # strength 0: the filter is the (near-perfect) STFT round trip
    out0 = den(torch.from_numpy(audio).to(dev), speaker_ids=ids.tolist(), strength=0.0).cpu().numpy()
    with pytest.raises(ValueError):                                      # the reference fails to broadcast [4] - [3] too
        den(torch.from_numpy(audio).to(dev), strength=0.0)
    assert max_abs(out0[:, 0], audio[:, :out0.shape[2]]) < 1e-4
