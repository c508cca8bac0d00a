"""world_size-2 gloo test (CPU) of the utterance-sharded path's host logic: round-robin shards,
the final waveform gather and the restoration of utterance order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cookietts_b200.packing import PackConfig
from cookietts_b200.parallel import infer_sharded


class FakeModel(torch.nn.Module):
    """Stands in for the GPU module: a deterministic per-utterance 'waveform' on CPU."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.pack_config = PackConfig()

    def infer(self, spect, sigma=1.0, z=None):
        hop = self.pack_config.hop_length
        base = spect.sum(dim=(1, 2)).unsqueeze(1)
        out = base + torch.arange(spect.shape[2] * hop, dtype=spect.dtype).unsqueeze(0) * sigma
        return out + (z if z is not None else 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_items, result_file):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    spect = torch.randn(n_items, 80, 3)
    z = torch.randn(n_items, 3 * 256)
    model = FakeModel()
    got = infer_sharded(model, spect, sigma=0.5, z=z, dst=0)
    if rank == 0:
        want = model.infer(spect, sigma=0.5, z=z)
        torch.save({"ok": bool(torch.equal(got, want)), "shape": tuple(got.shape)}, result_file)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def _run(n_items, world, tmp_path):
    result_file = str(tmp_path / f"res_{n_items}_{world}.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_items, result_file), nprocs=world, join=True)
    res = torch.load(result_file)
    assert res["ok"] and res["shape"] == (n_items, 768)


def _worker_long(rank, world, port, result_file):
    """Chunk-sharded long-form path with the oracle as the compute function (tiny model, CPU)."""
    import numpy as np
    from cookietts_b200.parallel import infer_long_sharded
    from oracle.waveglow_oracle import OracleConfig, synthetic_state_dict, synthetic_inputs
    from tests.test_chunking import OracleModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = OracleConfig(n_mel_channels=8, n_flows=4, n_group=8, n_early_every=2, n_early_size=2,
                       win_length=32, hop_length=8, n_layers=3, n_channels=16)
    sd = synthetic_state_dict(cfg, 3)
    mel, z = synthetic_inputs(cfg, 1, 130, 4)

    class M(OracleModel):
        def parameters(self):
            return iter([torch.zeros(1)])

    model = M(cfg, sd)
    got = infer_long_sharded(model, torch.from_numpy(mel).double(), 0.8, torch.from_numpy(z).double(), n_chunks=3)
    if rank == 0:
        full = model.infer(torch.from_numpy(mel).double(), 0.8, torch.from_numpy(z).double())
        torch.save({"err": float((got.double() - full).abs().max())}, result_file)
    dist.barrier()
    dist.destroy_process_group()


def test_long_form_chunks_sharded_over_ranks_gloo(tmp_path):
    result_file = str(tmp_path / "long.pt")
    mp.spawn(_worker_long, args=(2, _free_port(), result_file), nprocs=2, join=True)
    assert torch.load(result_file)["err"] < 1e-6       # gathered through fp32 buffers


def test_sharded_infer_gloo_even(tmp_path):
    _run(4, 2, tmp_path)


def test_sharded_infer_gloo_uneven(tmp_path):
    _run(5, 2, tmp_path)


def _worker_async(rank, world, port, result_file):
    from cookietts_b200.parallel import WaveformGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = WaveformGather((3, 16), torch.device("cpu"), dst=0, depth=2)
    ok = True
    slots = []
    for step in range(5):                       # more calls than slots: slot re-use waits for the older collective
        audio = torch.full((3, 16), float(10 * step + rank))
        slots.append(g.submit(audio))
        if rank == 0 and step >= 1:             # the previous call's slot is complete once it has been waited for
            prev = slots[step - 1]
            g.pending[prev][0].wait()
            ok &= all(bool((g.result(prev)[r] == 10 * (step - 1) + r).all()) for r in range(world))
    g.wait_all()
    if rank == 0:
        ok &= all(bool((g.result(slots[-1])[r] == 40 + r).all()) for r in range(world))
        ok &= slots == [0, 1, 0, 1, 0]
        torch.save({"ok": ok}, result_file)
    else:
        assert g.result(0) is None
    dist.barrier()
    dist.destroy_process_group()


def test_async_waveform_gather_gloo(tmp_path):
    result_file = str(tmp_path / "async.pt")
    mp.spawn(_worker_async, args=(2, _free_port(), result_file), nprocs=2, join=True)
    assert torch.load(result_file)["ok"]
