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


class SymmetricGather:
    """Gathered result buffers in symmetric memory (torch.distributed._symmetric_memory): every rank owns
    [world][P][cap] key-point / descriptor / count arrays and maps every peer's copy over NVLink, so that the
    extractor's descriptor kernel can store each record straight into all of them (adb_orb_set_gather)."""

    def __init__(self, world: int, rank: int, frames: int, cap: int, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank, self.P, self.cap = world, rank, frames, cap
        self.kp_bytes = world * frames * cap * 24
        self.desc_bytes = world * frames * cap * 32
        self.cnt_bytes = world * frames * 4
        total = self.kp_bytes + self.desc_bytes + self.cnt_bytes
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.peers = [self.hdl.get_buffer(r, (total,), torch.uint8) for r in range(world)]
        dist.barrier()

    def multicast_ptr(self) -> int:
        """NVLS multicast address of the symmetric buffer (0 when the fabric / driver has no multicast support)."""
        try:
            return int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        except Exception:   # noqa: BLE001
            return 0

    def targets(self, prefer_multicast: bool = True):
        """Pointer triples for ORBextractor.set_gather, offset to this rank's slot: ONE triple on the NVLS multicast mapping when the
        NVSwitch can replicate stores (a record leaves the GPU once), else one triple per peer.  Returns (kps, desc, counts, multicast)."""
        slot_kp, slot_desc, slot_cnt = self.P * self.cap * 24, self.P * self.cap * 32, self.P * 4
        mc = self.multicast_ptr() if prefer_multicast else 0
        if mc:
            return ([mc + self.rank * slot_kp], [mc + self.kp_bytes + self.rank * slot_desc],
                    [mc + self.kp_bytes + self.desc_bytes + self.rank * slot_cnt], True)
        kps = [p.data_ptr() + self.rank * slot_kp for p in self.peers]
        desc = [p.data_ptr() + self.kp_bytes + self.rank * slot_desc for p in self.peers]
        cnts = [p.data_ptr() + self.kp_bytes + self.desc_bytes + self.rank * slot_cnt for p in self.peers]
        return kps, desc, cnts, False

    def kps_view(self):
        return self.buf[:self.kp_bytes].view(self.world, self.P, self.cap, 24)

    def desc_view(self):
        return self.buf[self.kp_bytes:self.kp_bytes + self.desc_bytes].view(self.world, self.P, self.cap, 32)

    def counts_view(self):
        return self.buf[self.kp_bytes + self.desc_bytes:].view(torch.int32).view(self.world, self.P)
