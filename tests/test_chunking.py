"""Host logic of the time-chunked (long-form) path: with the halo of
`cookietts_b200.parallel.plan_chunks`, chunked inference equals the un-chunked function on every
core sample.  The compute function here is the oracle (CPU); the GPU version of this test is in
tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from cookietts_b200.packing import PackConfig
from cookietts_b200.parallel import plan_chunks, halo_frames, infer_long, shard_indices, unshard_order
from oracle.waveglow_oracle import OracleConfig, synthetic_state_dict, synthetic_inputs, infer_with_z


class OracleModel:
    """`infer`-compatible adapter around the oracle (tests only)."""

    def __init__(self, cfg, sd):
        self.cfg, self.sd = cfg, sd
        self.pack_config = PackConfig(n_mel=cfg.n_mel_channels, n_flows=cfg.n_flows, n_group=cfg.n_group,
                                      n_early_every=cfg.n_early_every, n_early_size=cfg.n_early_size,
                                      win_length=cfg.win_length, hop_length=cfg.hop_length,
                                      n_layers=cfg.n_layers, n_channels=cfg.n_channels, kernel_size=cfg.kernel_size)

    def infer(self, spect, sigma=1.0, z=None):
        out = infer_with_z(self.sd, self.cfg, spect.numpy(), z.numpy(), sigma, np.float64)
        return torch.from_numpy(out)


def test_plan_covers_every_frame_once():
    pc = PackConfig()
    assert halo_frames(pc) == 96            # 12 flows x 255 steps / 32 steps per frame
    for t_mel, n in [(5168, 8), (861, 3), (10, 4), (7, 16)]:
        plan = plan_chunks(t_mel, n, pc)
        assert plan[0].core0 == 0 and plan[-1].core1 == t_mel
        for a, b in zip(plan, plan[1:]):
            assert a.core1 == b.core0
        for ch in plan:
            assert 0 <= ch.lo <= ch.core0 < ch.core1 <= ch.hi <= t_mel
            assert ch.lo == max(0, ch.core0 - 96 - 3) and ch.hi == min(t_mel, ch.core1 + 96)


def test_chunked_equals_unchunked_oracle():
    cfg = OracleConfig(n_mel_channels=8, n_flows=4, n_group=8, n_early_every=2, n_early_size=2,
                       win_length=32, hop_length=8, n_layers=3, n_channels=16)
    sd = synthetic_state_dict(cfg, 3)
    mel, z = synthetic_inputs(cfg, 2, 150, 4)
    model = OracleModel(cfg, sd)
    assert halo_frames(model.pack_config) == 28
    full = model.infer(torch.from_numpy(mel).double(), 0.8, torch.from_numpy(z).double())
    for n_chunks in (2, 3, 5):
        got = infer_long(model, torch.from_numpy(mel).double(), 0.8, torch.from_numpy(z).double(), n_chunks=n_chunks)
        assert got.shape == full.shape
        assert float((got - full).abs().max()) < 1e-12, n_chunks


def test_too_small_halo_is_detected():
    """Sanity of the test itself: cutting the halo makes the chunked result differ."""
    cfg = OracleConfig(n_mel_channels=8, n_flows=4, n_group=8, n_early_every=2, n_early_size=2,
                       win_length=32, hop_length=8, n_layers=3, n_channels=16)
    sd = synthetic_state_dict(cfg, 3)
    mel, z = synthetic_inputs(cfg, 1, 150, 4)
    model = OracleModel(cfg, sd)
    full = model.infer(torch.from_numpy(mel).double(), 0.8, torch.from_numpy(z).double())
    from cookietts_b200.parallel import Chunk
    bad = [Chunk(0, 75, 0, 80), Chunk(75, 150, 70, 150)]
    got = infer_long(model, torch.from_numpy(mel).double(), 0.8, torch.from_numpy(z).double(), chunks=bad)
    assert float((got - full).abs().max()) > 1e-6


def test_shard_indices_roundtrip():
    for n, world in [(256, 8), (5, 2), (3, 4), (16, 1)]:
        seen = sorted(i for r in range(world) for i in shard_indices(n, world, r))
        assert seen == list(range(n))
        assert sorted(unshard_order(n, world)) == list(range(n))
