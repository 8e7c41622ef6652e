"""Writes tests/golden/quadtree_ref.npz: results of the REFERENCE's own ORBextractor::DistributeOctTree / ExtractorNode::DivideNode
(src/ORBextractor.cc:497-765; the two function definitions compiled from /root/reference by `make -C oracle ref`, oracle/ref_orb.cpp)
on seeded candidate sets.  The reference orders nodes of equal size by heap ADDRESS (:686 sorts (size, ExtractorNode*) pairs), so its
result depends on the allocator; the fixture is taken with an allocator that hands out increasing addresses and reuses nothing, where
address order = creation order = convention D.1 of the oracle and the CUDA kernel.  The generator also counts how many cases the
stock malloc run reproduces (the tie-break is reached in most dense cases).  Run in the build container (needs /root/reference):

    python oracle/gen_ref_orb_golden.py
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_orb.so")
N_CASES = 120


def ref_distribute(L, cand, min_x, max_x, min_y, max_y, n, monotonic=1):
    cand = np.ascontiguousarray(cand, np.float32)
    out = np.zeros((len(cand) + 8, 3), np.float32)
    L.ref_distribute.restype = C.c_int
    k = L.ref_distribute(cand.ctypes.data_as(C.c_void_p), len(cand), min_x, max_x, min_y, max_y, int(n), out.ctypes.data_as(C.c_void_p), len(out), monotonic)
    return out[:k].copy()


def make_case(i):
    """Distinct integer pixels of a w x h region (one or two quad-tree roots), random responses, a quota from 1 to well above the count."""
    rng = np.random.default_rng(90000 + i)
    w = int(rng.integers(100, 1300))
    h = int(rng.integers(w // 2 + 1, w + 1)) if rng.random() < 0.7 else int(rng.integers(max(40, w // 5), w // 2 + 2))
    h = min(h, 2 * w - 1)                       # round(w / h) >= 1: the shapes the extractor accepts
    m = int(rng.integers(1, 7000)); n = int(rng.integers(1, 1200))
    pos = rng.choice(w * h, size=min(m, w * h), replace=False)
    cand = np.stack([pos % w, pos // w, rng.integers(1, 255, len(pos))], 1).astype(np.float32)
    return cand, w, h, n


def main():
    import oracle
    oracle.build()
    L = C.CDLL(LIB)
    crcs, counts, same_malloc, same_oracle = [], [], 0, 0
    keep = {}
    for i in range(N_CASES):
        cand, w, h, n = make_case(i)
        a = ref_distribute(L, cand, 0, w, 0, h, n, 1)
        crcs.append(zlib.crc32(a.tobytes())); counts.append(len(a))
        b = ref_distribute(L, cand, 0, w, 0, h, n, 0)
        same_malloc += int(a.shape == b.shape and (a == b).all())
        o = oracle.distribute(cand, 0, w, 0, h, n)
        same_oracle += int(a.shape == o.shape and (a == o).all())
        if i < 4:
            keep[f"out{i}"] = a
    print(f"{N_CASES} cases: the stock-malloc run equals the monotonic-allocator run in {same_malloc}; the oracle equals the latter in {same_oracle}")
    path = os.path.join(ROOT, "tests", "golden", "quadtree_ref.npz")
    np.savez_compressed(path, crc=np.array(crcs, np.int64), count=np.array(counts, np.int32), same_with_stock_malloc=np.int32(same_malloc), **keep)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
