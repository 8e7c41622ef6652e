"""Writes tests/golden/stereo_ref.npz: seeded stereo pairs and the outputs of the REFERENCE's own Frame::ComputeStereoMatches /
ORBmatcher::DescriptorDistance (oracle/_ref/libref_match.so = those two function bodies compiled from /root/reference by
`make -C oracle ref`, see oracle/ref_match.cpp).  The matcher's inputs (key-points, descriptors, pyramids of both images) are
produced by the extractor oracle, which has its own cv2-derived pins; the fixture stores the level-0 images, the extraction
parameters and the reference's mvuRight / mvDepth.  Run in the build container (needs /root/reference):

    python oracle/gen_ref_match_golden.py
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_match.so")


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_stereo(L, kl, dl, kr, dr, pyr_l, pyr_r, scale, mb, mbf):
    """Same packing as oracle.stereo_match (flattened pyramids + offsets) into the reference's function."""
    nl = len(pyr_l)
    lw = np.array([p.shape[1] for p in pyr_l], np.int32); lh = np.array([p.shape[0] for p in pyr_l], np.int32)
    off = np.zeros(nl, np.int64); off[1:] = np.cumsum(lw.astype(np.int64) * lh)[:-1]
    pl = np.concatenate([np.ascontiguousarray(p, np.uint8).ravel() for p in pyr_l])
    pr = np.concatenate([np.ascontiguousarray(p, np.uint8).ravel() for p in pyr_r])
    sc = np.ascontiguousarray(scale, np.float32); inv = (np.float32(1.0) / sc).astype(np.float32)
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr); dl = np.ascontiguousarray(dl); dr = np.ascontiguousarray(dr)
    ur = np.zeros(len(kl), np.float32); dp = np.zeros(len(kl), np.float32)
    L.ref_stereo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.ref_stereo_match(P(kl), P(dl), len(kl), P(kr), P(dr), len(kr), P(pl), P(pr), P(off), P(lw), P(lh), nl, P(sc), P(inv), mb, mbf, P(ur), P(dp))
    return ur, dp


def distinct_case(rng, n_points=300):
    """Observation sets of n_points map points: CSR (descriptors [total, 32], point_ptr)."""
    lens = rng.integers(0, 40, n_points); lens[:4] = [0, 1, 2, 128]
    ptr = np.zeros(n_points + 1, np.int32); ptr[1:] = np.cumsum(lens)
    base = rng.integers(0, 256, (n_points, 32), dtype=np.uint8)
    desc = np.repeat(base, lens, axis=0)
    flip = rng.random(desc.shape) < 0.04
    desc = (desc ^ (flip * rng.integers(1, 256, desc.shape)).astype(np.uint8)).astype(np.uint8)
    desc[ptr[5]:ptr[6]] = desc[ptr[5]]                            # all identical: every median is 0, the first row wins
    return np.ascontiguousarray(desc), ptr


CASES = [  # (seed, width, height, nfeatures, iniTh, minTh)
    (0, 640, 480, 1000, 12, 7), (7, 640, 480, 2000, 20, 7), (11, 320, 240, 500, 20, 7), (23, 752, 480, 1200, 12, 7)]


def main():
    import oracle
    from airdos_b200 import synth
    oracle.build()
    L = C.CDLL(LIB)
    out = {}
    for ci, (seed, w, h, nf, ini, mn) in enumerate(CASES):
        il, ir = synth.make_stereo_pair(seed, w, h)
        a = oracle.orb_extract(il, None, nf, 1.2, 8, ini, mn, want_pyramid=True)
        b = oracle.orb_extract(ir, None, nf, 1.2, 8, ini, mn, want_pyramid=True)
        sc = oracle.orb_params(nf, 1.2, 8, w, h)["scale"]
        mbf = synth.BF; mb = mbf / synth.FX
        ur, dp = ref_stereo(L, a["kps"], a["desc"], b["kps"], b["desc"], a["pyramid"], b["pyramid"], sc, mb, mbf)
        out[f"c{ci}_params"] = np.array([seed, w, h, nf, ini, mn], np.int32)
        if ci in (0, 2):                                   # two cases carry their images; the others are regenerated from the seed
            out[f"c{ci}_left"] = il; out[f"c{ci}_right"] = ir   # (airdos_b200/synth.py) and checked by CRC
        out[f"c{ci}_crc"] = np.array([zlib.crc32(il.tobytes()), zlib.crc32(ir.tobytes())], np.int64)
        out[f"c{ci}_u_right"] = ur; out[f"c{ci}_depth"] = dp
        print(f"case {ci}: {len(ur)} left key-points, {(dp > 0).sum()} stereo matches")
    # DescriptorDistance on random descriptor pairs
    rng = np.random.default_rng(20261017)
    d = rng.integers(0, 256, (512, 2, 32), dtype=np.uint8)
    d[:8, 1] = d[:8, 0]; d[8:16, 1] = ~d[8:16, 0]
    L.ref_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    out["dd_pairs"] = d
    out["dd_dist"] = np.array([L.ref_descriptor_distance(P(np.ascontiguousarray(x[0])), P(np.ascontiguousarray(x[1]))) for x in d], np.int32)
    # MapPoint::ComputeDistinctiveDescriptors on seeded observation sets (0, 1, 2, 128 observations, all-identical rows, noisy copies)
    desc, ptr = distinct_case(np.random.default_rng(20261018))
    chosen = np.full((len(ptr) - 1, 32), 0xAB, np.uint8)          # untouched where the point has no observation
    L.ref_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ref_distinctive(P(desc), P(ptr), len(ptr) - 1, P(chosen))
    out["distinct_crc"] = np.array([zlib.crc32(desc.tobytes()), zlib.crc32(ptr.tobytes())], np.int64)
    out["distinct_chosen"] = chosen
    path = os.path.join(ROOT, "tests", "golden", "stereo_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
