"""Multi-GPU plumbing for the extractor: frames are independent units, so ranks take contiguous
frame ranges and the only data-path collective is one all-gather of the fixed-stride result
records.  One process per GPU, torch.distributed (NCCL over NVLink on the
GPU box; gloo in the CPU tests).  BA does not shard: replicas only."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_frames(n_frames: int, rank: int, world: int):
    """Contiguous, equal-sized shards (the all-gather needs equal send counts): returns
    (first, count, padded_count).  The last ranks may own fewer real frames than padded_count."""
    per = (n_frames + world - 1) // world
    first = min(rank * per, n_frames)
    return first, max(0, min(per, n_frames - first)), per


class DevicePtr:
    """Expose a raw device allocation of the C-ABI as a torch tensor (zero copy) through
    __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def as_tensor(ptr: int, shape, typestr: str, device) -> torch.Tensor:
    return torch.as_tensor(DevicePtr(ptr, shape, typestr), device=device)


def all_gather_records(kps: torch.Tensor, desc: torch.Tensor, counts: torch.Tensor, group=None):
    """kps [B, cap, 24] u8, desc [B, cap, 32] u8, counts [B] i32 on every rank ->
    ([W*B, cap, 24], [W*B, cap, 32], [W*B]) on every rank, rank-major (= global frame order)."""
    world = dist.get_world_size(group)
    outs = []
    for t in (kps, desc, counts):
        t = t.contiguous()
        o = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(o, t, group=group)
        outs.append(o)
    return tuple(outs)


def gathered_frame(kps_all: torch.Tensor, desc_all: torch.Tensor, counts_all: torch.Tensor, frame: int):
    """Key-points / descriptors of global frame `frame` out of the gathered buffers (host copies)."""
    n = int(counts_all[frame])
    k = kps_all[frame, :n].cpu().numpy().copy().view(np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                                                              ("response", "<f4"), ("octave", "<i4")])).reshape(n)
    return k, desc_all[frame, :n].cpu().numpy()
