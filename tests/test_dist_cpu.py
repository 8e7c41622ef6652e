"""world_size-2 gloo test of the sharding + all-gather plumbing (host logic only, CPU tensors)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_frames, cap, q):
    sys.path.insert(0, ROOT)
    from airdos_b200 import dist as adist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count, per = adist.shard_frames(n_frames, rank, world)
    # deterministic fake records: frame g has (g % 7) + 1 key-points whose bytes encode (g, i)
    kps = torch.zeros(per, cap, 24, dtype=torch.uint8); desc = torch.zeros(per, cap, 32, dtype=torch.uint8)
    counts = torch.zeros(per, dtype=torch.int32)
    for j in range(count):
        g = first + j
        n = g % 7 + 1
        counts[j] = n
        for i in range(n):
            kps[j, i] = (g * 31 + i) % 251
            desc[j, i] = (g * 17 + i * 3) % 253
    K, D, C = adist.all_gather_records(kps, desc, counts)
    ok = K.shape == (world * per, cap, 24) and C.shape == (world * per,)
    for g in range(n_frames):
        r, j = divmod(g, per)
        slot = r * per + j
        n = g % 7 + 1
        ok &= int(C[slot]) == n
        ok &= bool((K[slot, :n, 0] == torch.tensor([(g * 31 + i) % 251 for i in range(n)], dtype=torch.uint8)).all())
        ok &= bool((D[slot, :n, 5] == torch.tensor([(g * 17 + i * 3) % 253 for i in range(n)], dtype=torch.uint8)).all())
    ok &= int(C.sum()) == sum(g % 7 + 1 for g in range(n_frames))   # padding frames carry count 0
    q.put((rank, bool(ok), first, count, per))
    dist.destroy_process_group()


def test_shard_and_all_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, cap = 11, 9      # ragged: 6 + 5 frames
    procs = [ctx.Process(target=_worker, args=(r, 2, 29517, n_frames, cap, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert res[0][1] and res[1][1]
    assert (res[0][2], res[0][3], res[0][4]) == (0, 6, 6) and (res[1][2], res[1][3], res[1][4]) == (6, 5, 6)


def test_shard_frames_covers_everything():
    from airdos_b200.dist import shard_frames
    for n in (0, 1, 7, 128, 1024, 1025):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                f, c, per = shard_frames(n, r, w)
                seen += list(range(f, f + c))
                assert c <= per
            assert seen == list(range(n))
    assert shard_frames(1024, 3, 8) == (384, 128, 128)
