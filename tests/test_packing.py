"""Host logic: the packed/folded weight form reproduces the reference (golden vectors),
so a kernel that implements `tests/packed_eval.py` exactly is parity-correct."""
import numpy as np
import pytest

from cookietts_b200.packing import PackConfig, pack_state_dict, split_hi_lo, bf16_bits_to_f32
from oracle.waveglow_oracle import snr_db
from tests.helpers import load_golden, max_abs
from tests.packed_eval import packed_infer


def pack_cfg(cfg):
    return PackConfig(n_mel=cfg.n_mel_channels, n_flows=cfg.n_flows, n_group=cfg.n_group,
                      n_early_every=cfg.n_early_every, n_early_size=cfg.n_early_size,
                      win_length=cfg.win_length, hop_length=cfg.hop_length, n_layers=cfg.n_layers,
                      n_channels=cfg.n_channels, kernel_size=cfg.kernel_size, cond_hidden=cfg.cond_hidden,
                      speaker_embed_dim=cfg.speaker_embed_dim, rezero=cfg.rezero)


@pytest.mark.parametrize("name", ["tiny", "small", "rezero"])
def test_packed_form_matches_reference(name):
    cfg, sd, g = load_golden(name)
    pc = pack_cfg(cfg)
    pk = pack_state_dict(sd, pc, planes=("f32",))
    out = packed_infer(pk, pc, g["mel"], g["z"], float(g["sigma"]))
    ref = g["audio_ref_fp64"]
    assert max_abs(out, ref) < 5e-5          # fp32-rounded packed weights, fp64 arithmetic
    assert snr_db(ref, out) > 90.0


def test_split_hi_lo_is_16_bit_accurate():
    rs = np.random.RandomState(0)
    w = rs.standard_normal(10000) * np.exp(rs.uniform(-8, 8, 10000))
    hi, lo = split_hi_lo(w)
    rec = bf16_bits_to_f32(hi).astype(np.float64) + bf16_bits_to_f32(lo).astype(np.float64)
    assert np.all(np.abs(rec - w) <= np.abs(w) * 2.0 ** -16)


@pytest.mark.parametrize("mode,snr_min,abs_max", [("bf16x3", 85.0, 2e-4), ("bf16", 40.0, 5e-2)])
def test_emulated_tensor_modes_error_budget(mode, snr_min, abs_max):
    """Predicts the tensor-core modes' error (operand rounding only; fp64 accumulate)."""
    cfg, sd, g = load_golden("small")
    pc = pack_cfg(cfg)
    pk = pack_state_dict(sd, pc)
    out = packed_infer(pk, pc, g["mel"], g["z"], float(g["sigma"]), emulate=mode)
    ref = g["audio_ref_fp64"]
    assert snr_db(ref, out) > snr_min
    assert max_abs(out, ref) < abs_max
