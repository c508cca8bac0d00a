"""cookietts_b200 - B200 (sm_100a) implementation of CookieTTS's WaveGlow inverse pass.

Public surface: `WaveGlow` (drop-in for `CookieTTS/_4_mtw/waveglow/glow.py::WaveGlow`,
inverse pass only) and the C ABI in `include/cwg.h` (`cookietts_b200/libcwg.so`)."""
from .waveglow import WaveGlow  # noqa: F401
from .packing import PackConfig, pack_state_dict  # noqa: F401
from .waveflow import WaveFlow  # noqa: F401
from .waveglow_ax import WaveGlowAx  # noqa: F401
from .denoiser import Denoiser  # noqa: F401
from . import serving  # noqa: F401
from .flow_decoder import FlowDecoder  # noqa: F401
