"""The Denoiser oracle against the unmodified reference STFT / Denoiser.forward (tests/golden/denoiser_*.npz)."""
import os

import numpy as np
import pytest

from oracle.denoiser_oracle import StftOracle, denoise, bias_spectrum
from tests.helpers import GOLDEN_DIR, max_abs

CASES = ["denoiser_22k", "denoiser_48k", "denoiser_small"]


@pytest.mark.parametrize("name", CASES)
def test_denoiser_oracle(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    sr = int(g["sampling_rate"])
    st = StftOracle(sr // 40, sr // 400, sr // 40)
    mag, _ = st.transform(g["audio"])
    assert max_abs(mag, g["magnitude"]) < 2e-4 * max(1.0, np.abs(g["magnitude"]).max())     # reference is fp32
    bias = bias_spectrum(st, g["bias_audio"])
    assert max_abs(bias, g["bias_spec"]) < 1e-5
    out = denoise(st, g["audio"], g["bias_spec"], float(g["strength"]))
    assert out.shape == g["denoised"].shape
    assert max_abs(out, g["denoised"]) < 5e-5
