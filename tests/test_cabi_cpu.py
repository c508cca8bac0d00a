"""CPU checks of the drop-in boundary: libcwg.so loads, exports every symbol include/cwg.h
declares, validates arguments (no compute calls without a GPU), and the Python module keeps
the reference's state_dict layout."""
import os
import re

import numpy as np
import pytest
import torch

from cookietts_b200 import WaveGlow, _cabi
from cookietts_b200.packing import PackConfig
from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def module_kwargs(cfg):
    return dict(n_mel_channels=cfg.n_mel_channels, n_flows=cfg.n_flows, n_group=cfg.n_group,
                n_early_every=cfg.n_early_every, n_early_size=cfg.n_early_size,
                WN_config=dict(n_layers=cfg.n_layers, n_channels=cfg.n_channels, kernel_size=cfg.kernel_size,
                               speaker_embed_dim=cfg.speaker_embed_dim, rezero=cfg.rezero),
                win_length=cfg.win_length, hop_length=cfg.hop_length)


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "cwg.h")).read()
    declared = set(re.findall(r"\b(cwg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cwg_abi_version() == _cabi.ABI_VERSION == 5


def test_workspace_and_argument_validation():
    lib = _cabi.load()
    cfg = _cabi.make_config(PackConfig())
    n_ffma = lib.cwg_workspace_bytes(cfg, _cabi.MODE_FFMA, 2, 10)
    n_tc = lib.cwg_workspace_bytes(cfg, _cabi.MODE_BF16X3, 2, 10)
    assert n_ffma > 0 and n_tc > 0 and n_tc < n_ffma
    assert lib.cwg_workspace_bytes(cfg, 7, 2, 10) == 0
    assert b"mode" in lib.cwg_last_error()
    bad = _cabi.make_config(PackConfig(n_group=8, hop_length=252))
    assert lib.cwg_workspace_bytes(bad, _cabi.MODE_FFMA, 1, 1) == 0
    # tensor-core modes are specialised for 256 channels: other widths are refused loudly
    c512 = _cabi.make_config(PackConfig(n_channels=512))
    assert lib.cwg_workspace_bytes(c512, _cabi.MODE_FFMA, 1, 4) > 0
    assert lib.cwg_launch_count(cfg, _cabi.MODE_FFMA) == 1 + 12 * (1 + 3 * 8 + 1)
    # f16f8: 6-byte x planes instead of 4; only for the 256-channel layer kernel
    n_f8 = lib.cwg_workspace_bytes(cfg, _cabi.MODE_F16F8, 2, 10)
    assert n_tc < n_f8 < n_ffma
    assert lib.cwg_launch_count(cfg, _cabi.MODE_F16F8) == lib.cwg_launch_count(cfg, _cabi.MODE_BF16X3) + 1   # + the range guard scan
    assert lib.cwg_workspace_bytes(c512, _cabi.MODE_F16F8, 1, 4) == 0 and b"F16F8" in lib.cwg_last_error()
    # maximum size: batch * T' beyond what one call indexes is refused with a message, not wrapped around
    assert lib.cwg_workspace_bytes(cfg, _cabi.MODE_F16F8, 4096, 8192) == 0 and b"too large" in lib.cwg_last_error()
    # NULL weights -> error code, not a crash
    rc = lib.cwg_infer(cfg, None, _cabi.MODE_FFMA, None, None, None, 1.0, None, None, 0, 1, 1, None)
    assert rc != 0 and b"NULL" in lib.cwg_last_error()


@pytest.mark.parametrize("name", ["tiny", "rezero"])
def test_state_dict_layout_matches_reference(name):
    cfg, sd, _ = load_golden(name)
    model = WaveGlow(**module_kwargs(cfg))
    own = model.state_dict()
    assert set(own.keys()) == set(sd.keys())
    for k, v in sd.items():
        assert tuple(own[k].shape) == tuple(v.shape), k
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)


def test_plain_weight_checkpoint_is_accepted():
    cfg, sd, _ = load_golden("tiny")
    from oracle.waveglow_oracle import weight_norm_effective
    plain = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            plain[k[:-9] + ".weight"] = weight_norm_effective(v, sd[k[:-9] + ".weight_v"])
        elif not k.endswith(".weight_v"):
            plain[k] = v
    model = WaveGlow(**module_kwargs(cfg))
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in plain.items()}, strict=True)
    from cookietts_b200.packing import pack_state_dict
    a = pack_state_dict({k: v.detach().numpy() for k, v in model.state_dict().items()}, model.pack_config, planes=("f32",))
    b = pack_state_dict(sd, model.pack_config, planes=("f32",))
    for key in ("w1_f32", "w2_f32", "cond_w_f32", "start_w"):
        assert np.allclose(a[key], b[key], atol=1e-6), key


def test_unsupported_constructor_values_raise():
    with pytest.raises(ValueError):
        WaveGlow(spect_scaling=True)
    with pytest.raises(ValueError):
        WaveGlow(memory_efficient=True)
    with pytest.raises(ValueError):
        WaveGlow(upsample_mode="simple")


def test_infer_refuses_cpu():
    cfg, sd, g = load_golden("tiny")
    model = WaveGlow(**module_kwargs(cfg))
    with pytest.raises(RuntimeError):
        model.infer(torch.from_numpy(g["mel"]))


def test_waveflow_state_dict_layout_and_packing_cpu():
    """Host logic of the WaveFlow path: reference key layout; packed form reproduces the oracle's
    WN_2d step (fp64) on CPU."""
    from cookietts_b200 import WaveFlow
    from cookietts_b200.waveflow import pack_waveflow_state_dict
    from oracle.make_golden_waveflow import reference_kwargs
    from oracle.waveflow_oracle import WaveFlowConfig, synthetic_state_dict, wn2d_step
    cfg = WaveFlowConfig(n_flows=2, n_layers=2)
    sd = synthetic_state_dict(cfg, 5)
    m = WaveFlow(**reference_kwargs(cfg))
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    pk = pack_waveflow_state_dict(sd, m.pack_config)
    # evaluate one AR step (row 2: full 3-row queue) of flow 1 from the packed arrays
    rs = np.random.RandomState(0)
    B, T, Cc, L = 1, 40, cfg.n_channels, cfg.n_layers
    rows = rs.standard_normal((3, B, T))
    cond_up = rs.standard_normal((B, cfg.n_mel_channels, T))
    k = 1
    from oracle.waveflow_oracle import _w
    w_c = _w(sd, f"WN.{k}.WN.cond_layers.0", np.float64)[:, :, 0]
    spec_all = np.einsum("oc,bct->bot", w_c, cond_up) + np.asarray(sd[f"WN.{k}.WN.cond_layers.0.bias"], np.float64)[None, :, None]
    queues = [None] * L
    for r in range(3):
        log_s, t = wn2d_step(sd, k, cfg, rows[r], spec_all, queues, np.float64)
    # packed evaluation: x rings per layer hold rows 0..2
    x = [pk["start_w"][k].astype(np.float64)[None, None, :] * rows[r][:, :, None] + pk["start_b"][k].astype(np.float64) for r in range(3)]
    eo = np.tile(pk["eo_b"][k].astype(np.float64), (B, T, 1))
    hist = [list(x)]            # hist[l][r] = layer-l input at row r
    for l in range(L):
        d = 2 ** l
        nxt = []
        for r in range(3):
            cols = []
            for a in range(3):
                rr = r - 2 + a
                src = hist[l][rr] if rr >= 0 else np.zeros((B, T, Cc))
                xp = np.zeros((B, T + 2 * d, Cc)); xp[:, d:d + T] = src
                cols += [xp[:, b * d:b * d + T] for b in range(3)]
            mel_pad = np.zeros((B, T, 128)); mel_pad[:, :, :cfg.n_mel_channels] = cond_up.transpose(0, 2, 1)
            a_mat = np.concatenate(cols + [mel_pad], axis=2)
            pre = a_mat @ pk["w1_f64"][k, l].T + pk["b1"][k, l].astype(np.float64)
            acts = np.tanh(pre[..., :Cc]) / (1 + np.exp(-pre[..., Cc:]))
            rsk = acts @ pk["w2_f64"][k, l].T
            nxt.append(hist[l][r] + rsk[..., :Cc] + pk["b2"][k, l].astype(np.float64))
            if r == 2:
                eo = eo + rsk[..., Cc:]
        hist.append(nxt)
    assert np.abs(eo[..., 0] - log_s).max() < 1e-6
    assert np.abs(eo[..., 1] - t).max() < 1e-6


def test_compat_import_paths():
    """`compat.install()` serves the reference import paths (run in a subprocess: it edits sys.modules)."""
    import subprocess, sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import cookietts_b200.compat as c; c.install()\n"
        "from CookieTTS._4_mtw.waveglow.glow import WaveGlow\n"
        "from CookieTTS._4_mtw.waveglow.efficient_model_ax import WaveGlow as Ax\n"
        "import cookietts_b200 as p\n"
        "assert WaveGlow is p.WaveGlow\n"
        "from oracle.make_golden_waveflow import reference_kwargs, reference_kwargs_ax1d\n"
        "from oracle.waveflow_oracle import WaveFlowConfig\n"
        "from oracle.waveglow_ax_oracle import AxConfig\n"
        "assert isinstance(Ax(**reference_kwargs(WaveFlowConfig(n_flows=2, n_layers=1))), p.WaveFlow)\n"
        "assert isinstance(Ax(**reference_kwargs_ax1d(AxConfig(n_flows=2, n_layers=1, n_channels=8))), p.WaveGlowAx)\n"
        "assert c.is_ax({'upsample_first': True}) and not c.is_ax({'n_flows': 12})\n"
        "print('ok')\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_checkpoint_key_remap():
    """train.py:98-99,122 key renames (NVIDIA checkpoint -> ax layout, legacy names)."""
    from cookietts_b200.compat import remap_checkpoint_keys, load_checkpoint
    nv = {"WN.0.in_layers.1.weight_v": 1, "WN.3.res_skip_layers.0.bias": 2, "WN.2.start.weight_g": 3, "WN.2.end.weight": 4,
          "convinv.5.conv.weight": 5, "upsample.weight": 6}
    out = remap_checkpoint_keys(nv, nvidia_checkpoint=True)
    assert set(out) == {"WN.0.WN.in_layers.1.weight_v", "WN.3.WN.res_skip_layers.0.bias", "WN.2.WN.start.weight_g",
                        "WN.2.WN.end.weight", "convinv.5.weight", "upsample.weight"}
    legacy = remap_checkpoint_keys({"invconv1x1.0.weight": 1, "WNs.1.F.start.bias": 2})
    assert set(legacy) == {"convinv.0.weight", "WN.1.WN.start.bias"}
    cfg, sd, _ = load_golden("tiny")
    model = WaveGlow(**module_kwargs(cfg))
    it = load_checkpoint(model, {"model": {k: torch.from_numpy(v) for k, v in sd.items()}, "iteration": 7})
    assert it == 7


@pytest.mark.parametrize("variant", [dict(merge_res_skip=True, gated_unit="GTRU"), dict(res_skip=False, merge_res_skip=True, gated_unit="SPTU"),
                                     dict(dilations_w=[3, 1], dilations_h=2, gated_unit="GSIU")])
def test_waveflow_variant_packing_cpu(variant):
    """WN_2d config variants in the packed fp32 layout (merged / absent res_skip = zero res rows, height dilations = deeper conv
    queues, gated units): replaying cwg_wf_ffma.cu's per-row layer loop in numpy on the packed arrays reproduces the oracle's
    WN_2d steps."""
    import warnings
    from cookietts_b200 import WaveFlow
    from cookietts_b200.waveflow import pack_waveflow_state_dict
    from oracle.make_golden_waveflow import reference_kwargs
    from oracle.waveflow_oracle import WaveFlowConfig, synthetic_state_dict, wn2d_step, _w
    from oracle.waveglow_ax_oracle import GATED_UNITS
    cfg = WaveFlowConfig(n_mel_channels=6, n_flows=2, n_group=8, n_layers=2, n_channels=8, win_length=32, hop_length=8, **variant)
    sd = synthetic_state_dict(cfg, 6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = WaveFlow(**reference_kwargs(cfg))
    assert m.precision == "ffma"
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    pk = pack_waveflow_state_dict(sd, m.pack_config)
    rs = np.random.RandomState(1)
    B, T, Cc, L, kh, kw = 2, 23, cfg.n_channels, cfg.n_layers, cfg.kernel_size_h, cfg.kernel_size_w
    n_rows = 6
    rows = rs.standard_normal((n_rows, B, T))
    cond_up = rs.standard_normal((B, cfg.n_mel_channels, T))
    k = 1
    w_c = _w(sd, f"WN.{k}.WN.cond_layers.0", np.float64)[:, :, 0]
    spec_all = np.einsum("oc,bct->bot", w_c, cond_up) + np.asarray(sd[f"WN.{k}.WN.cond_layers.0.bias"], np.float64)[None, :, None]
    queues = [None] * L
    ua, ub = GATED_UNITS[cfg.gated_unit.upper()]
    hist = [[] for _ in range(L + 1)]                 # hist[l][r]: layer-l input at row r, [B, T, C] (the conv queues of the kernel)
    for r in range(n_rows):
        log_s, t = wn2d_step(sd, k, cfg, rows[r], spec_all, queues, np.float64)
        hist[0].append(pk["start_w"][k].astype(np.float64)[None, None, :] * rows[r][:, :, None] + pk["start_b"][k].astype(np.float64))
        eo = np.tile(pk["eo_b"][k].astype(np.float64), (B, T, 1))
        for l in range(L):
            d, dh = cfg.dilation_w(l), cfg.dilation_h(l)
            pad = ((kw - 1) * d) // 2
            cols = []
            for a in range(kh):
                rr = r - (kh - 1 - a) * dh            # wff_a: src_row = row - (KH - 1 - a) * dil_h
                src = hist[l][rr] if rr >= 0 else np.zeros((B, T, Cc))
                xp = np.zeros((B, T + 2 * pad, Cc)); xp[:, pad:pad + T] = src
                cols += [xp[:, b * d:b * d + T] for b in range(kw)]
            a_mat = np.concatenate(cols + [cond_up.transpose(0, 2, 1)], axis=2)
            pre = a_mat @ pk["w1_f64"][k, l].T + pk["b1"][k, l].astype(np.float64)
            acts = ua(pre[..., :Cc]) * ub(pre[..., Cc:])
            rsk = acts @ pk["w2_f64"][k, l].T
            hist[l + 1].append(hist[l][r] + rsk[..., :Cc] + pk["b2"][k, l].astype(np.float64))
            eo = eo + rsk[..., Cc:]
        assert np.abs(eo[..., 0] - log_s).max() < 1e-9 and np.abs(eo[..., 1] - t).max() < 1e-9, r
        if cfg.merge_res_skip:                        # the hidden tensor is never updated
            assert all(np.array_equal(hist[l][r], hist[0][r]) for l in range(1, L))
