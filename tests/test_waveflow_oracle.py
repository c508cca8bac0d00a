"""The WaveFlow oracle is pinned against the unmodified reference ax model
(tests/golden/waveflow_*.npz, made by oracle/make_golden_waveflow.py)."""
import json
import os

import numpy as np
import pytest

from oracle.waveflow_oracle import WaveFlowConfig, synthetic_state_dict, inverse, infer_with_z, permute_height
from oracle.waveglow_oracle import snr_db
from tests.helpers import GOLDEN_DIR, max_abs

CASES = ["waveflow_tiny", "waveflow_nearest", "waveflow_small", "waveflow_config5",
         # general WN_2d shapes: dense 5x3, depthwise-separable 7x7 at squeeze height 20 (16 and 128 channels)
         "waveflow_5x3", "waveflow_sep7", "waveflow_sep7_128",
         # WN_config variants: gated units, width / height dilations, merged / absent res_skip, WN speaker embedding
         "waveflow_v_gate", "waveflow_v_merge", "waveflow_v_noskip", "waveflow_v_speaker", "waveflow_v_early", "waveflow_v_mixlast",
         "waveflow_v_conv", "waveflow_v_conv_mixlast", "waveflow_v_cond", "waveflow_v_tconv"]


def load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = WaveFlowConfig(**json.loads(str(g["config"])))
    return cfg, synthetic_state_dict(cfg, int(g["weight_seed"])), g


@pytest.mark.parametrize("name", CASES)
def test_inverse_fp64(name):
    cfg, sd, g = load(name)
    ids = g["speaker_ids"] if "speaker_ids" in g.files and g["speaker_ids"].size else None
    out = inverse(sd, cfg, g["z"].astype(np.float64) * float(g["sigma"]), g["mel"], np.float64, speaker_ids=ids)
    assert max_abs(out, g["inverse_ref_fp64"]) < 1e-9


@pytest.mark.parametrize("name", CASES[:3])
def test_infer_fp64_and_fp32(name):
    cfg, sd, g = load(name)
    out = infer_with_z(sd, cfg, g["mel"], g["z"], float(g["sigma"]), 1, np.float64)
    assert out.shape == g["infer_ref_fp64"].shape
    assert max_abs(out, g["infer_ref_fp64"]) < 1e-9
    out32 = infer_with_z(sd, cfg, g["mel"], g["z"], float(g["sigma"]), 1, np.float32)
    assert max_abs(out32, g["infer_ref_fp32"]) < 2e-5
    assert snr_db(g["infer_ref_fp64"], out32) > 100.0


def test_permute_height_is_an_involution():
    x = np.arange(2 * 16 * 3).reshape(2, 16, 3)
    for k in range(8):
        assert np.array_equal(permute_height(permute_height(x, k), k), x)
    assert list(permute_height(x, 0)[0, :, 0] // 3) == list(range(15, -1, -1))
    assert list(permute_height(x, 2)[0, :, 0] // 3) == [7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8]


AX_CASES = ["waveglow_ax_tiny", "waveglow_ax_permute", "waveglow_ax_mixlast", "waveglow_ax_256"]


@pytest.mark.parametrize("name", AX_CASES)
def test_ax_1d_oracle_matches_reference(name):
    """oracle/waveglow_ax_oracle.py (ax model, waveflow=False) against the unmodified reference."""
    from oracle.waveglow_ax_oracle import AxConfig, synthetic_state_dict as ax_sd, inverse as ax_inverse, infer_with_z as ax_infer
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = AxConfig(**json.loads(str(g["config"])))
    sd = ax_sd(cfg, int(g["weight_seed"]))
    out = ax_inverse(sd, cfg, g["z"].astype(np.float64) * float(g["sigma"]), g["mel"], np.float64)
    assert max_abs(out, g["inverse_ref_fp64"]) < 1e-9
    if name != "waveglow_ax_256":
        out2 = ax_infer(sd, cfg, g["mel"], g["z"], float(g["sigma"]), 1, np.float64)
        assert out2.shape == g["infer_ref_fp64"].shape
        assert max_abs(out2, g["infer_ref_fp64"]) < 1e-9


AXV_CASES = ["waveglow_axv_gsirru", "waveglow_axv_merge", "waveglow_axv_noskip", "waveglow_axv_cond", "waveglow_axv_256",
             "waveglow_axv_tconv_crop", "waveglow_axv_tconv_interp", "waveglow_axv_group40"] + [
    "waveglow_axv_unit_" + u for u in ("gtru", "ttu", "stu", "gtsu", "gsiu", "gsiru", "gtsru", "gsirlru", "gsirrlru")]


@pytest.mark.parametrize("name", AXV_CASES)
def test_ax_wn_variants_oracle_matches_reference(name):
    """The WN_config variants (gated units, listed dilations, merged / absent res_skip, multi-layer cond stacks, glow_ax.py:
    168-198,:297-335,:399-414) of oracle/waveglow_ax_oracle.py against the unmodified reference."""
    from oracle.waveglow_ax_oracle import AxConfig, synthetic_state_dict as ax_sd, inverse as ax_inverse, infer_with_z as ax_infer
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = AxConfig(**json.loads(str(g["config"])))
    assert cfg.is_variant()
    sd = ax_sd(cfg, int(g["weight_seed"]))
    ids = g["speaker_ids"] if g["speaker_ids"].size else None
    # the SIREN units multiply rounding differences by 16 per layer: scale the bar with the reference's own fp32-vs-fp64 gap
    gap = max_abs(g["infer_ref_fp32"], g["infer_ref_fp64"])
    if not name.endswith("_crop"):       # (the reference's centre crop is empty for `inverse` on an un-padded mel)
        out = ax_inverse(sd, cfg, g["z"].astype(np.float64) * float(g["sigma"]), g["mel"], np.float64, speaker_ids=ids)
        assert max_abs(out, g["inverse_ref_fp64"]) < max(1e-9, 1e-4 * max_abs(g["inverse_ref_fp32"], g["inverse_ref_fp64"]))
    if cfg.n_channels <= 16:
        out2 = ax_infer(sd, cfg, g["mel"], g["z"], float(g["sigma"]), 1, np.float64, speaker_ids=ids)
        assert out2.shape == g["infer_ref_fp64"].shape
        assert max_abs(out2, g["infer_ref_fp64"]) < max(1e-9, 1e-4 * gap)
