"""Slice sharding across ranks: sampling has NO data-path collective.

The unit of work is one 2D slice (own x_T, noise, conditioning; GroupNorm / attention are per
sample), so the slice list — a batch, or the 155 slices of a BraTS volume from the reference's
``get_test_data_dict`` (trainers/trainer_ds_diff.py:256-265) — is partitioned over ranks, the weights
are replicated, and the only communication is the final gather of finished slices to rank 0, which
assembles the 2D->3D volume like ``on_predict_end`` (trainer_ds_diff.py:854-875).  Per-SLICE seeds make
the result independent of the world size.  One process per GPU; ``torch.distributed`` with NCCL over
NVLink on GPUs (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of slice indices for ``rank``; sizes differ by at most one and the first
    ``n_total % world`` ranks get the extra slice (155 slices on 8 ranks -> 20,20,20,19,...)."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError(f"bad shard request: n_total={n_total} rank={rank} world={world}")
    base, extra = divmod(n_total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def slice_noise(slice_ids: Sequence[int], shape: Sequence[int], base_seed: int, device, stream_id: int = 0) -> Tensor:
    """N(0,1) tensor [len(slice_ids), *shape] where slice i is drawn from its OWN generator
    (seed = base_seed + 1000003 * stream_id + slice id): identical for any world size / batch split."""
    out = torch.empty((len(slice_ids), *shape), dtype=torch.float32, device=device)
    for j, sid in enumerate(slice_ids):
        g = torch.Generator(device=device)
        g.manual_seed(int(base_seed) + 1000003 * int(stream_id) + int(sid))
        out[j] = torch.randn(tuple(shape), generator=g, device=device, dtype=torch.float32)
    return out


def gather_slices(local: Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> Optional[Tensor]:
    """Gather every rank's finished slices ``[n_r, ...]`` (ragged n_r allowed) to rank 0 in global slice
    order.  Returns the full ``[n_total, ...]`` tensor on rank 0, None elsewhere.  Single-process: identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local.shape[0] != n_total:
            raise ValueError("gather_slices: local slice count does not match n_total")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nmax = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    b, e = shard_range(n_total, rank, world)
    if local.shape[0] != e - b:
        raise ValueError(f"rank {rank}: expected {e - b} slices, got {local.shape[0]}")
    pad = torch.zeros((nmax, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: e - b] = local
    bufs: Optional[List[Tensor]] = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dst = dist.get_global_rank(group, 0) if group is not None else 0      # dst is a GLOBAL rank
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        rb, re = shard_range(n_total, r, world)
        parts.append(bufs[r][: re - rb])
    return torch.cat(parts, dim=0)


def pad_to_multiple(x: Tensor, multiple: int = 32) -> Tuple[Tensor, Tuple[int, int, int, int]]:
    """Zero-pad the last two dims symmetrically up to a multiple of ``multiple`` (what MONAI's ``DivisiblePadd(k=32)``
    does in the reference's train / val transforms, training_transform.py:261,404; the TEST transform leaves BraTS
    slices at 240x240, which a 6-level UNet cannot take — SURVEY §8 gotcha 8).  Returns (padded, (top, bottom, left,
    right)) for ``crop_pad``."""
    h, w = x.shape[-2], x.shape[-1]
    ph, pw = (-h) % multiple, (-w) % multiple
    pads = (ph // 2, ph - ph // 2, pw // 2, pw - pw // 2)
    if ph == 0 and pw == 0:
        return x, pads
    return torch.nn.functional.pad(x, (pads[2], pads[3], pads[0], pads[1])), pads


def crop_pad(x: Tensor, pads: Tuple[int, int, int, int]) -> Tensor:
    t, b, l, r = pads
    return x[..., t:x.shape[-2] - b, l:x.shape[-1] - r]


def sample_volume(sample_fn, cond: Tensor, out_channels: int = 1, batch_size: int = 8, base_seed: int = 2024,
                  multiple: int = 32, device=None, group: Optional[dist.ProcessGroup] = None) -> Optional[Tensor]:
    """Sample every 2D slice of one volume, sharded over the ranks of ``group`` (a 155-slice BraTS
    volume over 8 GPUs), and return the assembled ``[n_slices, out_channels, H, W]`` stack on rank 0 (None elsewhere) —
    the tensor ``on_predict_end`` turns into a NIfTI volume (trainers/trainer_ds_diff.py:854-875).

    ``cond``      [n_slices, C, H, W] conditioning slices of the volume (any device);
    ``sample_fn`` (cond_batch [b, C, H', W'], x_T [b, out_channels, H', W']) -> samples [b, out_channels, H', W'], e.g.
                  ``lambda c, x: DDIMSampler(model).sample(50, len(c), x.shape[1:], c, x_T=x, verbose=False)[0]``.
    Slices are padded to a multiple of ``multiple`` and cropped back; x_T of slice i comes from its own generator
    (``slice_noise``), so for DETERMINISTIC samplers (DDIM eta = 0, DPM-Solver++) the result does not depend on the world
    size or on ``batch_size``.  Samplers that draw per-step noise (ancestral, DDIM eta > 0) take it from the rank's global
    CUDA generator inside the step engine, batch-shaped: reproducible for a fixed (world size, batch_size, seed) only.
    No collective runs inside the sampling loop; the only communication is the final gather."""
    n_total = cond.shape[0]
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    device = device if device is not None else cond.device
    b, e = shard_range(n_total, rank, world)
    outs = []
    for s in range(b, e, batch_size):
        ids = list(range(s, min(e, s + batch_size)))
        c, pads = pad_to_multiple(cond[ids[0]:ids[-1] + 1].to(device).float(), multiple)
        x_T = slice_noise(ids, (out_channels, c.shape[-2], c.shape[-1]), base_seed, device)
        outs.append(crop_pad(sample_fn(c.contiguous(), x_T), pads))
    shape = (0, out_channels, cond.shape[-2], cond.shape[-1])
    local = torch.cat(outs, 0) if outs else torch.empty(shape, dtype=torch.float32, device=device)
    return gather_slices(local.contiguous(), n_total, group)
