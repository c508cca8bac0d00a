"""Loader of the torch.ops.cookietts_b200.* binding (csrc/torch_ops.cpp -> _cwg_torch.so, which links libcwg.so).

    torch.ops.cookietts_b200.waveglow_pack(tensors, names, cfg, mode)  -> packed blob (uint8, on the device)
    torch.ops.cookietts_b200.waveglow_infer(packed, cfg, mode, embed_dim, n_speakers, mel, speaker_ids, z, sigma,
                                            ev_begin, ev_end)          -> audio [B, T]

No fallback: a missing library raises."""
from __future__ import annotations

import os

import torch

from . import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
OPS_PATH = os.path.join(_HERE, "_cwg_torch.so")
_loaded = False


def load():
    global _loaded
    if _loaded:
        return torch.ops.cookietts_b200
    if not os.path.exists(OPS_PATH):
        raise _cabi.CwgError(f"{OPS_PATH} not found - build the CUDA extension first "
                             f"(python -c 'import __graft_entry__ as g; g.build()')")
    _cabi.load()                     # libcwg.so first: same file the shim's DT_NEEDED resolves to
    torch.ops.load_library(OPS_PATH)
    if int(torch.ops.cookietts_b200.abi_version()) != _cabi.ABI_VERSION:
        raise _cabi.CwgError("_cwg_torch.so was built against another libcwg ABI; rebuild")
    _loaded = True
    return torch.ops.cookietts_b200


def config_list(pc) -> list:
    """The 11 fields of cwg_config in declaration order (include/cwg.h)."""
    return [pc.n_mel, pc.n_flows, pc.n_group, pc.n_early_every, pc.n_early_size, pc.win_length, pc.hop_length,
            pc.n_layers, pc.n_channels, pc.kernel_size, pc.cond_hidden]
