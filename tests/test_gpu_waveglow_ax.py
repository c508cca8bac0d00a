"""GPU parity of the ax model's 1-D configuration (waveflow=False) against golden vectors of the
unmodified reference (`efficient_model_ax.WaveGlow`)."""
import json
import os

import numpy as np
import pytest
import torch

from cookietts_b200 import WaveGlowAx
from oracle.make_golden_waveflow import reference_kwargs_ax1d
from oracle.waveglow_ax_oracle import AxConfig, synthetic_state_dict
from oracle.waveglow_oracle import snr_db
from tests.helpers import GOLDEN_DIR, max_abs

pytestmark = pytest.mark.gpu

TOL = {"ffma": dict(max_abs=1e-4, snr=100.0), "bf16x3": dict(max_abs=1e-3, snr=60.0), "bf16": dict(max_abs=5e-2, snr=40.0),
       "f16f8": dict(max_abs=1e-3, snr=60.0)}


def run(name, precision):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = AxConfig(**json.loads(str(g["config"])))
    sd = synthetic_state_dict(cfg, int(g["weight_seed"]))
    m = WaveGlowAx(precision=precision, **reference_kwargs_ax1d(cfg))
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    m = m.cuda().eval()
    sigma = float(g["sigma"])
    z, mel = torch.from_numpy(g["z"]).cuda(), torch.from_numpy(g["mel"]).cuda()
    inv, _ = m.inverse(z * sigma, mel)
    aud = m.infer(mel, sigma=sigma, z=z)
    return inv.numpy(), aud.numpy(), g


@pytest.mark.parametrize("name", ["waveglow_ax_tiny", "waveglow_ax_permute", "waveglow_ax_mixlast", "waveglow_ax_256"])
def test_ax_fp32_cuda_cores(name):
    inv, aud, g = run(name, "ffma")
    assert max_abs(inv, g["inverse_ref_fp64"]) <= TOL["ffma"]["max_abs"]
    assert snr_db(g["inverse_ref_fp64"], inv) >= TOL["ffma"]["snr"]
    assert aud.shape == g["infer_ref_fp64"].shape
    assert max_abs(aud, g["infer_ref_fp64"]) <= TOL["ffma"]["max_abs"]


@pytest.mark.parametrize("precision", ["bf16x3", "bf16", "f16f8"])
def test_ax_tensor_cores(precision):
    inv, aud, g = run("waveglow_ax_256", precision)
    assert np.isfinite(inv).all()
    assert max_abs(inv, g["inverse_ref_fp64"]) <= TOL[precision]["max_abs"]
    assert snr_db(g["inverse_ref_fp64"], inv) >= TOL[precision]["snr"]
    assert max_abs(aud, g["infer_ref_fp64"]) <= TOL[precision]["max_abs"]
