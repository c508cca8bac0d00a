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
