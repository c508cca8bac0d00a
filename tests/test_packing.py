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


@pytest.mark.parametrize("name", ["tiny", "small", "rezero", "group24"])
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


def test_f16f8_planes_and_predicted_accuracy():
    """CWG_MODE_F16F8 packing: the planes decode to what include/cwg.h states, and the numpy evaluation of the kernel
    arithmetic (fp16 pass + two e5m2 correction passes in the in_layer GEMM, three fp16 passes elsewhere) stays far
    inside the fp32-path bar on a 256-channel golden case."""
    from cookietts_b200.packing import F8_P, F8_Q, e5m2_bits_to_f32
    cfg, sd, g = load_golden("mel20_256")
    pc = PackConfig(n_mel=cfg.n_mel_channels, n_flows=cfg.n_flows, n_group=cfg.n_group, n_early_every=cfg.n_early_every,
                    n_early_size=cfg.n_early_size, win_length=cfg.win_length, hop_length=cfg.hop_length,
                    n_layers=cfg.n_layers, n_channels=cfg.n_channels, kernel_size=cfg.kernel_size, cond_hidden=256)
    pk = pack_state_dict(sd, pc, planes=("f16f8",))
    ref = pack_state_dict(sd, pc, planes=("f32",))["w1_f32"].astype(np.float64)
    w16 = pk["w1_hi"].view(np.float16).astype(np.float64)
    assert np.abs(w16 - ref).max() <= np.abs(ref).max() * 2.0 ** -11
    h8 = e5m2_bits_to_f32(pk["w1_h8"]).astype(np.float64) * 2.0 ** F8_P
    l8 = e5m2_bits_to_f32(pk["w1_l8"]).astype(np.float64) * 2.0 ** -F8_Q
    big = np.abs(w16) > 2.0 ** -8                      # e5m2 normal range after the 2^-6 scaling
    assert np.abs(h8 - w16)[big].max() <= 0.126 * np.abs(w16)[big].max()
    assert (np.abs(h8 - w16)[big] <= 0.126 * np.abs(w16)[big]).all()
    d = ref - w16
    assert (np.abs(l8 - d) <= 0.126 * np.abs(d) + 2.0 ** -24).all()
    from tests.packed_eval import packed_infer
    out = packed_infer(pk, pc, g["mel"], g["z"], float(g["sigma"]), emulate="f16f8")
    assert max_abs(out, g["audio_ref_fp64"]) < 1e-4


@pytest.mark.parametrize("name", ["axfe_speaker_cond", "axfe_nb_small"])
def test_ax_packed_form_matches_reference(name):
    """The ax packing - including the wide group layout (n_group 24), the WN-level speaker embedding as a per-utterance gate
    bias and upsample_first=False run as interpolate-then-contract - reproduces the reference's own output."""
    from cookietts_b200.waveglow_ax import pack_ax_state_dict
    from oracle.ax_frontend_oracle import frontend
    from oracle.waveflow_oracle import upsample_cond
    from tests.ax_frontend_helpers import load_case
    from tests.packed_eval import packed_ax_inverse
    kind, cfg, fe, sd, g = load_case(name)
    mel = np.concatenate([g["mel"].astype(np.float64), np.zeros(g["mel"].shape[:2] + (1,))], axis=2)   # artifact_trimming
    samples = (mel.shape[2] - 1) * cfg.hop_length
    samples -= samples % cfg.n_group
    cond = frontend(sd, fe, mel, g["speaker_ids"], samples // cfg.n_group, np.float64)
    if cond.shape[2] != samples // cfg.n_group:            # frame-rate cond: what cwg_ax_infer's k_mel_up interpolates
        cond = upsample_cond(cond, samples // cfg.n_group, cfg.upsample_mode)
    pc = PackConfig(n_mel=cond.shape[1], n_flows=cfg.n_flows, n_group=cfg.n_group, n_early_every=cfg.n_early_every,
                    n_early_size=cfg.n_early_size, win_length=cfg.hop_length, hop_length=cfg.hop_length, n_layers=cfg.n_layers,
                    n_channels=cfg.n_channels, kernel_size=cfg.kernel_size, cond_hidden=cond.shape[1])
    pk = pack_ax_state_dict(sd, pc, cfg.channel_mixing, planes=("f32",), wn_speaker_dim=cfg.wn_speaker_embed_dim)
    mg = 16 if cfg.n_group <= 16 else 32
    assert pk["eo_b"].shape == (cfg.n_flows, mg) and pk["winv"].shape == (cfg.n_flows, mg, mg)
    assert pk["start_w"].shape[2] == mg // 2 and pk["w2_f32"].shape[2] == cfg.n_channels + mg
    z = g["z"].astype(np.float64)[:, :samples] * float(g["sigma"])
    out = packed_ax_inverse(pk, pc, cond, z, cfg.mix_first, g["speaker_ids"])[:, :-cfg.hop_length]
    ref = g["infer_ref_fp64"]
    assert out.shape == ref.shape
    assert max_abs(out, ref) < 5e-5 and snr_db(ref, out) > 90.0
