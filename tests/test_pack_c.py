"""cwg_pack_weights (csrc/cwg_pack.cu, fp64 on the device) against the numpy statement of the same algebra
(cookietts_b200/packing.py), and torch.ops.cookietts_b200.* end to end against the reference's golden vectors."""
import numpy as np
import pytest
import torch

from cookietts_b200 import WaveGlow, _cabi
from cookietts_b200.packing import pack_state_dict, bf16_bits_to_f32, e5m2_bits_to_f32
from oracle.waveglow_oracle import snr_db
from tests.helpers import load_golden, max_abs
from tests.test_cabi_cpu import module_kwargs


def test_ops_library_registers_schemas():
    """CPU: _cwg_torch.so loads, links libcwg.so and registers the ops the module calls (no compute)."""
    from cookietts_b200 import _torch_ops
    ops = _torch_ops.load()
    assert int(ops.abi_version()) == _cabi.ABI_VERSION
    for name in ("waveglow_pack", "waveglow_infer", "waveglow_launch_count"):
        assert hasattr(ops, name)
    cfg = [80, 12, 8, 4, 2, 1024, 256, 8, 256, 3, 256]
    assert int(ops.waveglow_launch_count(cfg, _cabi.MODE_F16F8)) > 100
    with pytest.raises((RuntimeError, NotImplementedError)):          # no CPU kernel is registered: CPU tensors must not run
        ops.waveglow_pack([torch.zeros(1)], "x", cfg, 0)


def _planes(name, planes):
    cfg, sd, _ = load_golden(name)
    m = WaveGlow(precision={"f32": "ffma", "hi": "bf16x3", "f16f8": "f16f8"}[planes[0]], **module_kwargs(cfg))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.cuda().eval()
    got = {k: v.cpu() for k, v in m.packed_views().items()}
    ref = pack_state_dict(sd, m.pack_config, planes=planes)
    return got, ref, m


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny", "rezero", "speaker", "config1", "c512", "group24"])
def test_pack_fp32_planes(name):
    got, ref, m = _planes(name, ("f32",))
    for k in ("cond_w_f32", "w1_f32", "w2_f32", "b1", "b2", "eo_b", "start_w", "start_b", "winv", "cond_b_base"):
        a, b = got[k].numpy().astype(np.float64), ref[k].astype(np.float64)
        assert a.shape == b.shape, k
        scale = max(np.abs(b).max(), 1e-30)
        assert np.abs(a - b).max() <= 2e-7 * scale, (k, np.abs(a - b).max(), scale)       # fp32 rounding of fp64 sums
    if m.multispeaker:
        assert np.abs(got["cond_w_spk"].numpy() - ref["cond_w_spk"]).max() <= 2e-7 * np.abs(ref["cond_w_spk"]).max()
        for k, wn in enumerate(m.WN):
            assert torch.equal(got["spk_embed"][k], wn.speaker_embed.weight.detach().cpu())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["config1", "c512", "group24_256"])
def test_pack_bf16_planes(name):
    got, ref, _ = _planes(name, ("hi", "lo"))
    for k in ("cond_w", "w1", "w2"):
        a = got[k + "_hi"].float().double() + got[k + "_lo"].float().double()
        b = bf16_bits_to_f32(ref[k + "_hi"]).astype(np.float64) + bf16_bits_to_f32(ref[k + "_lo"]).astype(np.float64)
        scale = np.abs(b).max()
        assert float((a - torch.from_numpy(b)).abs().max()) <= 2e-7 * scale, k
        same = (got[k + "_hi"].view(torch.int16).numpy().view(np.uint16) == ref[k + "_hi"]).mean()
        assert same > 0.9999, (k, same)          # identical rounding except where the fp64 sums differ in the last bits
    if name == "config1":                        # layer-0 fold planes (256-channel models)
        a = got["w0_hi"].float().double() + got["w0_lo"].float().double()
        b = bf16_bits_to_f32(ref["w0_hi"]).astype(np.float64) + bf16_bits_to_f32(ref["w0_lo"]).astype(np.float64)
        assert float((a - torch.from_numpy(b)).abs().max()) <= 2e-7 * np.abs(b).max()
        assert np.abs(b).max() > 0


@pytest.mark.gpu
def test_pack_f16f8_planes():
    got, ref, _ = _planes("config1", ("f16f8",))
    for k in ("cond_w", "w1", "w2"):
        a = got[k + "_hi"].float().double() + got[k + "_lo"].float().double()
        b = ref[k + "_hi"].view(np.float16).astype(np.float64) + ref[k + "_lo"].view(np.float16).astype(np.float64)
        assert float((a - torch.from_numpy(b)).abs().max()) <= 2e-7 * np.abs(b).max(), k
    for k in ("w1_h8", "w1_l8", "w2_h8", "w2_l8"):
        a = got[k].view(torch.uint8).numpy()
        same = (a == ref[k]).mean()
        assert same > 0.999, (k, same)
        assert np.abs(e5m2_bits_to_f32(a) - e5m2_bits_to_f32(ref[k])).max() <= 0.26 * np.abs(e5m2_bits_to_f32(ref[k])).max()


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol,snr", [("ffma", 1e-4, 100.0), ("f16f8", 5e-4, 80.0)])
def test_torch_ops_end_to_end(precision, tol, snr):
    """torch.ops.cookietts_b200.waveglow_pack / waveglow_infer driven directly (no Python module in between) against the
    reference's config-1 golden."""
    from cookietts_b200 import _torch_ops
    ops = _torch_ops.load()
    cfg, sd, g = load_golden("config1")
    m = WaveGlow(precision=precision, **module_kwargs(cfg))
    names = list(sd.keys())
    cfgl = _torch_ops.config_list(m.pack_config)
    mode = _cabi.MODES[precision]
    blob = ops.waveglow_pack([torch.from_numpy(sd[k]).cuda() for k in names], "\n".join(names), cfgl, mode)
    out, status = ops.waveglow_infer(blob, cfgl, mode, 0, 0, torch.from_numpy(g["mel"]).cuda(), None,
                                     torch.from_numpy(g["z"]).cuda(), float(g["sigma"]), [], [])
    torch.cuda.synchronize()
    assert int(status.item()) == 0                 # cwg_infer_status: finite waveform, fp16 planes in range
    out = out.cpu().numpy()
    ref = g["audio_ref_fp64"]
    assert max_abs(out, ref) <= tol and snr_db(ref, out) >= snr


@pytest.mark.gpu
def test_pack_rejects_bad_state_dict():
    from cookietts_b200 import _torch_ops
    ops = _torch_ops.load()
    cfg, sd, _ = load_golden("tiny")
    m = WaveGlow(precision="ffma", **module_kwargs(cfg))
    names = [k for k in sd.keys() if k != "WN.0.end.bias"]
    with pytest.raises(RuntimeError, match="WN.0.end.bias"):
        ops.waveglow_pack([torch.from_numpy(sd[k]).cuda() for k in names], "\n".join(names), _torch_ops.config_list(m.pack_config), 0)
