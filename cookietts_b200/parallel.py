"""Sharding of the WaveGlow inverse pass: by utterance across GPUs and by halo-overlapped
chunks along time (long-form audio).

The reference `infer` is single-process and processes a whole utterance in one tensor
(SURVEY 2.2, 5 "long-context").  The path shards naturally (glow.py:314-350 has no reduction
over the batch, and every output sample depends on a bounded window of mel frames and latent),
so this layer is net-new host logic:

* utterance sharding: rank r takes utterances r, r+world, ...; weights are replicated; the only
  collective is the final gather of waveforms (never inside a flow step);
* time chunking: a chunk `[c0, c1)` of mel frames is computed from frames
  `[c0 - halo - (J-1), c1 + halo)` with the latent sliced by position, and only the core is kept.
  `halo = ceil(n_flows * (2^L - 1) * (k-1)/2 / P)` frames covers the receptive field of the
  dilated convs of all flows (glow.py:167-171), `J-1` more frames on the left feed the
  ConvTranspose1d taps (glow.py:238-241).  With that halo the chunked result equals the
  un-chunked function on every core sample (tests/test_chunking.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from .packing import PackConfig


def shard_indices(n_items: int, world: int, rank: int) -> List[int]:
    """Utterance indices owned by `rank` (round-robin, like the reference's per-GPU process
    layout in distributed.py:146-171)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world))


def unshard_order(n_items: int, world: int) -> List[int]:
    """Position in the rank-major concatenation of every utterance index: gathered[j] is
    utterance order[j]."""
    order = []
    for r in range(world):
        order.extend(shard_indices(n_items, world, r))
    return order


def halo_frames(pc: PackConfig) -> int:
    steps = pc.n_flows * (2 ** pc.n_layers - 1) * (pc.kernel_size - 1) // 2
    return -(-steps // pc.phases)


@dataclass(frozen=True)
class Chunk:
    core0: int      # first mel frame of the core
    core1: int      # one past the last core frame
    lo: int         # first frame actually computed (core0 - halo - (J-1), clamped)
    hi: int         # one past the last frame computed (core1 + halo, clamped)


def plan_chunks(t_mel: int, n_chunks: int, pc: PackConfig) -> List[Chunk]:
    """Split `t_mel` frames into `n_chunks` contiguous cores (sizes differ by at most 1)."""
    if n_chunks < 1 or t_mel < 1:
        raise ValueError("need t_mel >= 1 and n_chunks >= 1")
    n_chunks = min(n_chunks, t_mel)
    h, warm = halo_frames(pc), pc.taps - 1
    base, rem = divmod(t_mel, n_chunks)
    chunks, c0 = [], 0
    for i in range(n_chunks):
        c1 = c0 + base + (1 if i < rem else 0)
        chunks.append(Chunk(c0, c1, max(0, c0 - h - warm), min(t_mel, c1 + h)))
        c0 = c1
    return chunks


def infer_chunk(model, spect: torch.Tensor, z: torch.Tensor, sigma: float, ch: Chunk, **kw) -> torch.Tensor:
    """Audio of the core of one chunk: [B, (core1-core0)*hop]."""
    hop = model.pack_config.hop_length
    out = model.infer(spect[:, :, ch.lo:ch.hi], sigma=sigma, z=z[:, ch.lo * hop:ch.hi * hop], **kw)
    return out[:, (ch.core0 - ch.lo) * hop:(ch.core1 - ch.lo) * hop]


def infer_long(model, spect: torch.Tensor, sigma: float = 1.0, z: Optional[torch.Tensor] = None,
               n_chunks: int = 8, chunks: Optional[Sequence[Chunk]] = None, **kw) -> torch.Tensor:
    """Long-form inference by halo-overlapped chunks on ONE device (the multi-GPU variant hands
    `plan_chunks(...)[rank::world]` to each rank and gathers)."""
    pc = model.pack_config
    B, _, t_mel = spect.shape
    if z is None:
        z = model.draw_z(B, t_mel)
    plan = list(chunks) if chunks is not None else plan_chunks(t_mel, n_chunks, pc)
    parts = [infer_chunk(model, spect, z, sigma, ch, **kw) for ch in plan]
    return torch.cat(parts, dim=1)


def infer_long_sharded(model, spect: torch.Tensor, sigma: float = 1.0, z: Optional[torch.Tensor] = None,
                       n_chunks: Optional[int] = None, dst: int = 0, **kw) -> Optional[torch.Tensor]:
    """Long-form inference of the SAME utterance(s) across ranks (BASELINE config 4): the mel is cut
    into `n_chunks` (default: world size) halo-overlapped chunks, rank r computes chunks r, r+world, ...
    on its own GPU with the latent sliced by position, and rank `dst` receives the stitched waveform
    `[B, T_mel*hop]` (None elsewhere).  `spect` and `z` must be identical on every rank (pass an
    explicit z; ranks cannot draw the same latent independently).  Needs an initialised process group."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    pc = model.pack_config
    B, _, t_mel = spect.shape
    if z is None:
        raise ValueError("infer_long_sharded needs an explicit latent z shared by all ranks")
    plan = plan_chunks(t_mel, n_chunks or world, pc)
    hop = pc.hop_length
    dev = next(model.parameters()).device
    # every rank sends only the CORES of its chunks, packed back to back into a buffer sized for the largest per-rank
    # total (cores differ by at most one frame, so the padding is at most one frame per chunk) - not a full-length,
    # mostly-zero waveform
    per_rank = [sum(ch.core1 - ch.core0 for ch in plan[r::world]) for r in range(world)]
    width = max(per_rank) * hop
    mine = torch.zeros(B, width, device=dev)
    off = 0
    for ch in plan[rank::world]:
        n = (ch.core1 - ch.core0) * hop
        mine[:, off:off + n] = infer_chunk(model, spect, z, sigma, ch, **kw)
        off += n
    bufs = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty(B, t_mel * hop, device=dev)
    offs = [0] * world
    for i, ch in enumerate(plan):
        r, n = i % world, (ch.core1 - ch.core0) * hop
        out[:, ch.core0 * hop:ch.core1 * hop] = bufs[r][:, offs[r]:offs[r] + n]
        offs[r] += n
    return out


class WaveformGather:
    """The path's only collective - the final gather of `[B_local, T]` waveforms to rank `dst` - taken off the critical
    path: `submit(audio)` issues it asynchronously (`async_op=True`: NCCL runs it on its own stream once `audio` is
    complete) into one of `depth` receive slots, so the gather of call i overlaps the kernels of call i+1 instead of
    sitting between them.  A slot is only waited for when it is about to be re-used (`depth` calls later) or in
    `wait_all()`; `result(slot)` is the list of per-rank tensors on `dst` (valid after the wait).  Works on NCCL and gloo."""

    def __init__(self, shape, device, dst: int = 0, depth: int = 2, dtype=torch.float32):
        import torch.distributed as dist
        self.dist, self.dst, self.depth = dist, dst, depth
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.bufs = ([[torch.empty(shape, device=device, dtype=dtype) for _ in range(self.world)] for _ in range(depth)]
                     if self.rank == dst else None)
        self.pending = [None] * depth
        self.calls = 0

    def submit(self, audio: torch.Tensor) -> int:
        slot = self.calls % self.depth
        if self.pending[slot] is not None:
            self.pending[slot][0].wait()
        work = self.dist.gather(audio, self.bufs[slot] if self.bufs is not None else None, dst=self.dst, async_op=True)
        self.pending[slot] = (work, audio)        # keeps `audio` alive until the collective has consumed it
        self.calls += 1
        return slot

    def wait_all(self):
        for i, p in enumerate(self.pending):
            if p is not None:
                p[0].wait()
                self.pending[i] = None

    def result(self, slot: int):
        return self.bufs[slot] if self.bufs is not None else None


def slice_per_utterance(kw: dict, idx, n_items: int) -> dict:
    """Per-utterance keyword arguments (speaker_id / speaker_ids: one entry per utterance) follow the shard: entries
    `idx` of every tensor / sequence of length `n_items`; scalars and one-element tensors pass through."""
    out = {}
    for k, v in kw.items():
        if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == n_items and n_items > 1:
            out[k] = v[idx] if not isinstance(idx, slice) else v[idx]
        elif isinstance(v, (list, tuple)) and len(v) == n_items and n_items > 1:
            out[k] = [v[i] for i in (range(*idx.indices(n_items)) if isinstance(idx, slice) else idx)]
        else:
            out[k] = v
    return out


def gather_waveforms(audio_local: torch.Tensor, n_items: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Final collective of the sharded path: gathers every rank's `[B_local, T]` waveforms to
    `dst` and restores utterance order.  Ranks may own different counts (round-robin shards differ
    by at most one), so shards are padded to the largest count for the fixed-size gather.
    Returns the `[n_items, T]` tensor on `dst`, None elsewhere.  Works on NCCL (GPU) and gloo (CPU)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    T = audio_local.shape[1]
    max_local = -(-n_items // world)
    padded = audio_local.new_zeros(max_local, T)
    padded[:audio_local.shape[0]] = audio_local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    out = audio_local.new_empty(n_items, T)
    for r in range(world):
        idx = shard_indices(n_items, world, r)
        if idx:
            out[idx] = bufs[r][:len(idx)]
    return out


def infer_sharded(model, spect: torch.Tensor, sigma: float = 1.0, z: Optional[torch.Tensor] = None,
                  dst: int = 0, **kw) -> Optional[torch.Tensor]:
    """Data-parallel `infer` over the utterances of `spect` ([N, n_mel, T_mel], same on every
    rank): each rank computes its round-robin shard on its own GPU, rank `dst` receives all
    waveforms in utterance order.  Needs an initialised process group."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    n = spect.shape[0]
    idx = shard_indices(n, world, rank)
    hop = model.pack_config.hop_length
    if idx:
        z_local = z[idx] if z is not None else None
        audio = model.infer(spect[idx], sigma=sigma, z=z_local, **slice_per_utterance(kw, idx, n))
    else:
        audio = torch.zeros(0, spect.shape[2] * hop, device=next(model.parameters()).device)
    return gather_waveforms(audio, n, dst)
