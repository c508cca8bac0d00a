"""The ax front-end oracle against the unmodified reference (tests/golden/axfe_*.npz, oracle/make_golden_ax_frontend.py),
and the drop-in modules' state_dict layout with the front-end switched on."""
import numpy as np
import pytest
import torch

from tests.ax_frontend_helpers import CASES, load_case, oracle_infer, module_kwargs
from tests.helpers import max_abs


@pytest.mark.parametrize("name", CASES)
def test_frontend_oracle_matches_reference(name):
    kind, cfg, fe, sd, g = load_case(name)
    if name in ("axfe_256", "axfe_waveflow"):
        pytest.importorskip("scipy")
    out = oracle_infer(kind, cfg, fe, sd, g["mel"], g["z"], g["speaker_ids"], float(g["sigma"]), np.float64)
    ref = g["infer_ref_fp64"]
    assert out.shape == ref.shape
    assert max_abs(out, ref) < 1e-8 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name", CASES)
def test_dropin_state_dict_layout(name):
    from cookietts_b200 import WaveGlowAx, WaveFlow
    kind, cfg, fe, sd, g = load_case(name)
    kw = module_kwargs(kind, cfg, fe)
    model = (WaveGlowAx if kind == "ax" else WaveFlow)(precision="ffma" if cfg.n_channels not in (128, 256) else "bf16x3", **kw)
    own = model.state_dict()
    assert set(own.keys()) == set(sd.keys())
    for k, v in sd.items():
        assert tuple(own[k].shape) == tuple(v.shape), k
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)


def test_notebook_model_constructs_with_the_reference_layout():
    """`notebook_ax_kwargs()` is cell 2 of the reference's speed-test notebook; the module built from it has exactly the
    parameter names / shapes of the checkpoint the reference loaded with strict=True when `axfe_notebook` was generated,
    and the case's own kwargs describe the same model."""
    from cookietts_b200 import WaveGlowAx
    from cookietts_b200.synthetic import notebook_ax_kwargs
    kind, cfg, fe, sd, g = load_case("axfe_notebook")
    model = WaveGlowAx(precision="f16f8", **notebook_ax_kwargs())
    own = model.state_dict()
    assert set(own.keys()) == set(sd.keys())
    for k, v in sd.items():
        assert tuple(own[k].shape) == tuple(v.shape), k
    other = WaveGlowAx(precision="f16f8", **module_kwargs(kind, cfg, fe))
    assert {k: tuple(v.shape) for k, v in other.state_dict().items()} == {k: tuple(v.shape) for k, v in own.items()}
    assert (model.n_group, model.n_flows, model.mix_first, model.channel_mixing, model.wn_speaker_embed_dim) == (24, 48, False, "permuteheight", 96)
    # variants that do not commute with the interpolation leave the packed kernels: general fp32 mode, with a warning
    kw = notebook_ax_kwargs()
    kw["WN_config"] = dict(kw["WN_config"], cond_layers=2, n_layers=2)
    kw["n_flows"] = 2
    with pytest.warns(UserWarning, match="general fp32"):
        m2 = WaveGlowAx(**kw)
    assert m2.general and m2.precision == "ffma"
    assert "WN.0.WN.cond_layers.1.weight_v" in m2.state_dict()
    # what is still not built raises: a WN-level upsample net with upsample_first=True (the reference mis-sizes it)
    kw["WN_config"] = dict(kw["WN_config"], transposed_conv_scales=[2])
    kw["upsample_first"] = True
    with pytest.raises(NotImplementedError):
        WaveGlowAx(**kw)
