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


TABLE_CONFIGS = [(2000, 1.2, 8), (1000, 1.2, 8), (1500, 1.2, 8), (500, 1.5, 5), (3000, 1.1, 12), (1200, 2.0, 4), (700, 1.3, 1)]
N_DESC_CASES = 6


def make_desc_case(i):
    """A level image (noise, smooth gradients + noise, or blocks: flat regions exercise the t0 < t1 ties and the zero-moment angle),
    and key-point positions kept 19 pixels inside like the extractor's."""
    rng = np.random.default_rng(91000 + i)
    w, h = int(rng.integers(80, 400)), int(rng.integers(60, 300))
    if i % 3 == 0:
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
    elif i % 3 == 1:
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.clip(128 + 90 * np.sin(xx / 9.0 + i) * np.cos(yy / 7.0) + rng.normal(0, 6, (h, w)), 0, 255).astype(np.uint8)
    else:
        img = np.kron(rng.integers(0, 4, ((h + 15) // 16, (w + 15) // 16)) * 80, np.ones((16, 16), int))[:h, :w].astype(np.uint8)
    n = int(rng.integers(50, 600))
    xy = np.stack([rng.integers(19, w - 19, n), rng.integers(19, h - 19, n)], 1).astype(np.float32)
    return img, xy


def tables_main(L):
    """tests/golden/extractor_tables_ref.npz: the reference's own constructor (src/ORBextractor.cc:411-472: scale / sigma^2 tables, quotas,
    umax, and the pattern it copies from bit_pattern_31_ :151-409) and its IC_Angle / computeOrbDescriptor (:78-148) on seeded images."""
    import oracle
    out = {}
    same = 0
    for c, (nf, sf, nl) in enumerate(TABLE_CONFIGS):
        r = oracle.orb_tables(nf, sf, nl, lib=L); o = oracle.orb_tables(nf, sf, nl)
        same += int(all((r[k].view(np.int32) == o[k].view(np.int32)).all() for k in r))
        for k, v in r.items():
            if k != "pattern" or c == 0:
                out[f"t{c}_{k}"] = v
    dsame = 0
    for i in range(N_DESC_CASES):
        img, xy = make_desc_case(i)
        bl = oracle.blur7(img)
        ra, rd = oracle.orient_describe(img, bl, xy, lib=L); oa, od = oracle.orient_describe(img, bl, xy)
        dsame += int((ra.view(np.int32) == oa.view(np.int32)).all() and (rd == od).all())
        out[f"d{i}_angle"] = ra; out[f"d{i}_desc"] = rd
    print(f"tables: oracle equals the reference constructor in {same} of {len(TABLE_CONFIGS)} configurations; "
          f"IC_Angle / computeOrbDescriptor bit-identical in {dsame} of {N_DESC_CASES} images")
    path = os.path.join(ROOT, "tests", "golden", "extractor_tables_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


# ORBextractor::operator() as a whole: (width, height, nfeatures, iniThFAST, minThFAST, person mask, image seed)
EXTRACT_CASES = [(640, 480, 1000, 20, 7, False, 1), (640, 480, 2000, 20, 7, True, 2), (320, 240, 500, 12, 7, True, 3), (752, 480, 1500, 20, 7, False, 4),
                 (200, 150, 300, 20, 7, True, 5), (640, 360, 1500, 30, 3, False, 6), (1241, 376, 2000, 20, 7, True, 7)]


def make_extract_case(i):
    from airdos_b200 import synth
    w, h, nf, ini, mn, masked, seed = EXTRACT_CASES[i]
    img = synth.make_stereo_pair(seed, w, h)[0]
    msk = synth.make_human_mask(seed, w, h, 2) if masked else None
    return img, msk, nf, ini, mn


def extract_main(L):
    """tests/golden/extractor_ref.npz: the reference's own ORBextractor::operator() / ComputePyramid / ComputeKeyPointsOctTree /
    computeOrientation / computeDescriptors (src/ORBextractor.cc:474-481, 767-864, 1040-1156; whole definitions compiled from /root/reference)
    on seeded images, their OpenCV calls landing in the oracle's primitives (oracle/ref_orb.cpp)."""
    import oracle
    out = {}
    same = 0
    for i in range(len(EXTRACT_CASES)):
        img, msk, nf, ini, mn = make_extract_case(i)
        k, d, pyr = oracle.ref_orb_extract(L, img, msk, nf, 1.2, 8, ini, mn)
        o = oracle.orb_extract(img, msk, nf, 1.2, 8, ini, mn, want_pyramid=True)
        ok = len(k) == len(o["kps"]) and k.tobytes() == o["kps"].tobytes() and bool((d == o["desc"]).all()) and \
            bool((pyr == np.concatenate([l.ravel() for l in o["pyramid"]])).all())
        same += int(ok)
        out[f"e{i}_n"] = np.int32(len(k)); out[f"e{i}_kps_crc"] = np.int64(zlib.crc32(k.tobytes())); out[f"e{i}_desc_crc"] = np.int64(zlib.crc32(d.tobytes()))
        out[f"e{i}_pyr_crc"] = np.int64(zlib.crc32(pyr.tobytes()))
        out[f"e{i}_per_level"] = np.bincount(k["octave"], minlength=8).astype(np.int32)
        if i in (2, 4):
            out[f"e{i}_kps"] = k; out[f"e{i}_desc"] = d
        print(f"extract {i}: {EXTRACT_CASES[i]} -> {len(k)} key-points, per level {list(out[f'e{i}_per_level'])}: oracle {'identical' if ok else 'DIFFERENT'}")
    print(f"operator(): the oracle equals the reference function in {same} of {len(EXTRACT_CASES)} images (key-points, descriptors, pyramid: bit for bit)")
    path = os.path.join(ROOT, "tests", "golden", "extractor_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


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
    tables_main(L)
    extract_main(L)


if __name__ == "__main__":
    main()
