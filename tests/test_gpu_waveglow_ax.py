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
    ids = torch.from_numpy(g["speaker_ids"]).cuda() if "speaker_ids" in g.files and g["speaker_ids"].size else None
    aud = m.infer(mel, speaker_ids=ids, sigma=sigma, z=z)
    if name.endswith("_crop"):          # WN-level upsample net with a centre crop: `inverse` on an un-padded mel has nothing to
        with pytest.raises(RuntimeError, match="cannot be cropped"):      # crop (the reference returns an empty slice there)
            m.inverse(z * sigma, mel, speaker_ids=ids)
        inv = aud
    else:
        inv, _ = m.inverse(z * sigma, mel, speaker_ids=ids)
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


AXV_CASES = ["waveglow_axv_gsirru", "waveglow_axv_merge", "waveglow_axv_noskip", "waveglow_axv_cond", "waveglow_axv_256",
             "waveglow_axv_tconv_crop", "waveglow_axv_tconv_interp", "waveglow_axv_group40"] + [
    "waveglow_axv_unit_" + u for u in ("gtru", "ttu", "stu", "gtsu", "gsiu", "gsiru", "gtsru", "gsirlru", "gsirrlru")]


@pytest.mark.parametrize("name", AXV_CASES)
def test_ax_wn_variants_general_fp32_mode(name):
    """WN_config variants (the 14 gated units, listed dilations, merged / absent res_skip, multi-layer cond stacks;
    glow_ax.py:168-198,:297-335,:399-414) through cwg_axg_flow against the unmodified reference.  The SIREN units multiply
    rounding differences by 16 per layer, so the bar scales with the reference's own fp32-vs-fp64 gap."""
    with pytest.warns(UserWarning, match="general fp32"):
        inv, aud, g = run(name, "bf16x3")                 # a tensor-core precision request falls back to fp32, loudly
    bar = max(TOL["ffma"]["max_abs"], 20 * max_abs(g["inverse_ref_fp32"], g["inverse_ref_fp64"]))
    assert np.isfinite(inv).all()
    assert max_abs(inv, g["inverse_ref_fp64"]) <= bar
    assert aud.shape == g["infer_ref_fp64"].shape
    assert max_abs(aud, g["infer_ref_fp64"]) <= max(TOL["ffma"]["max_abs"], 20 * max_abs(g["infer_ref_fp32"], g["infer_ref_fp64"]))


def test_ax_general_mode_matches_packed_kernels_on_a_plain_model():
    """The general fp32 mode and the packed FFMA layer kernels are two implementations of the same WN: on a plain GTU model
    (forced through the general mode by spelling out its 2^i dilations as a shifted list and back) they agree to fp32 noise."""
    g = np.load(os.path.join(GOLDEN_DIR, "waveglow_ax_tiny.npz"))
    cfg = AxConfig(**json.loads(str(g["config"])))
    sd = synthetic_state_dict(cfg, int(g["weight_seed"]))
    kw = reference_kwargs_ax1d(cfg)
    outs = []
    for general in (False, True):
        m = WaveGlowAx(precision="ffma", **kw)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        m = m.cuda().eval()
        if general:
            m.general, m._dilations = True, [2 ** i for i in range(cfg.n_layers)]
        inv, _ = m.inverse(torch.from_numpy(g["z"]).cuda() * float(g["sigma"]), torch.from_numpy(g["mel"]).cuda())
        outs.append(inv.numpy())
    assert max_abs(outs[0], outs[1]) <= 2e-5
    assert max_abs(outs[1], g["inverse_ref_fp64"]) <= TOL["ffma"]["max_abs"]
