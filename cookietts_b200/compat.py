"""Import-path compatibility: make a CookieTTS checkout pick up the B200 classes unchanged.

    import cookietts_b200.compat as compat; compat.install()
    from CookieTTS._4_mtw.waveglow.glow import WaveGlow                 # -> cookietts_b200.WaveGlow
    from CookieTTS._4_mtw.waveglow.efficient_model_ax import WaveGlow   # -> ax factory below

`install()` registers stand-in modules for the two reference module paths the call sites import
(scripts/*.ipynb, train.py:385-397, denoiser callers) in `sys.modules`; nothing else of CookieTTS
is shadowed.  The ax entry is a factory because one reference class covers two of ours: it returns
`WaveFlow` for `waveflow=True` and `WaveGlowAx` for `waveflow=False` (efficient_model_ax.py:19-20,57-62).
"""
from __future__ import annotations

import sys
import types

from .waveglow import WaveGlow
from .waveflow import WaveFlow
from .waveglow_ax import WaveGlowAx


def AxWaveGlow(*args, **kwargs):
    """Drop-in for `efficient_model_ax.WaveGlow(**waveglow_config)`."""
    import inspect
    names = list(inspect.signature(WaveFlow.__init__).parameters)[1:]
    bound = dict(zip(names, args))
    bound.update(kwargs)
    return WaveFlow(**bound) if bound.get("waveflow", True) else WaveGlowAx(**bound)


def is_ax(config: dict) -> bool:
    """The reference's own selector (scripts/inference.ipynb): ax configs carry `upsample_first`."""
    return "upsample_first" in config


def build_vocoder(waveglow_config: dict, precision: str = "bf16x3"):
    """Instantiate the right class for a reference `waveglow_config` dict."""
    if is_ax(waveglow_config):
        return AxWaveGlow(precision=precision, **waveglow_config)
    return WaveGlow(precision=precision, **waveglow_config)


def remap_checkpoint_keys(model_dict: dict, nvidia_checkpoint: bool = False) -> dict:
    """Key renames `load_checkpoint` applies before `load_state_dict` (train.py:98-99,122): NVIDIA/waveglow
    checkpoints into the ax layout (`.in_layers` -> `.WN.in_layers`, ..., `.conv.weight` -> `.weight`), and the
    legacy names `invconv1x1` -> `convinv`, `.F.` -> `.WN.`, `WNs.` -> `WN.`."""
    out = dict(model_dict)
    if nvidia_checkpoint:
        out = {k.replace(".in_layers", ".WN.in_layers").replace(".res_skip_layers", ".WN.res_skip_layers")
                .replace(".start", ".WN.start").replace(".end", ".WN.end").replace(".conv.weight", ".weight"): v
               for k, v in out.items()}
    return {k.replace("invconv1x1", "convinv").replace(".F.", ".WN.").replace("WNs.", "WN."): v for k, v in out.items()}


def load_checkpoint(model, checkpoint_dict: dict, nvidia_checkpoint: bool = False, strict: bool = True):
    """`model.load_state_dict` of a reference checkpoint dict (`{'model': state_dict | nn.Module, ...}`,
    train.py:93-99,135-143); returns the stored iteration (0 if absent)."""
    sd = checkpoint_dict["model"]
    if hasattr(sd, "state_dict"):                       # checkpoints that pickled the whole module (train.py:94-95)
        sd = sd.state_dict()
    model.load_state_dict(remap_checkpoint_keys(sd, nvidia_checkpoint), strict=strict)
    return int(checkpoint_dict.get("iteration", 0))


def install() -> None:
    glow = types.ModuleType("CookieTTS._4_mtw.waveglow.glow")
    glow.WaveGlow = WaveGlow
    ax = types.ModuleType("CookieTTS._4_mtw.waveglow.efficient_model_ax")
    ax.WaveGlow = AxWaveGlow
    for name in ("CookieTTS", "CookieTTS._4_mtw", "CookieTTS._4_mtw.waveglow"):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = []
            sys.modules[name] = pkg
    sys.modules["CookieTTS._4_mtw.waveglow.glow"] = glow
    sys.modules["CookieTTS._4_mtw.waveglow.efficient_model_ax"] = ax
    sys.modules["CookieTTS._4_mtw.waveglow"].glow = glow
    sys.modules["CookieTTS._4_mtw.waveglow"].efficient_model_ax = ax
