"""Writes tests/golden/bow_ref.npz: outputs of the REFERENCE's own vocabulary-bucket searches -- ORBmatcher::SearchByBoW(KeyFrame*,
Frame&, ...) (src/ORBmatcher.cc:159-288) and ORBmatcher::SearchForTriangulation with CheckDistEpipolarLine (:657-823, 140-157) --
compiled from /root/reference by `make -C oracle ref` (oracle/ref_match.cpp) and run on the seeded problems of
airdos_b200/synth.py::make_bow_problem.  The reference walks two DBoW2::FeatureVectors (std::map node id -> feature indices); the
problems carry the common nodes only, so the generator gives both sides extra one-sided nodes (holding the features that sit in no
common node) between the common ones, which exercises the lower_bound branches of the walk.  Run in the build container:

    python oracle/gen_ref_bow_golden.py
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_match.so")

CASES = [  # (seed, mode, n1, n2, n_nodes, check_orientation)
    (0, 0, 1800, 2000, 400, 1), (1, 0, 1500, 1500, 60, 1), (2, 0, 2500, 3000, 1500, 0),
    (3, 1, 1800, 2000, 400, 1), (4, 1, 1200, 1500, 60, 1), (5, 1, 2500, 2500, 1500, 0)]


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def problem(case):
    from airdos_b200 import synth
    seed, mode, n1, n2, nn, chk = case
    pr = synth.make_bow_problem(seed, mode=mode, n1=n1, n2=n2, n_nodes=nn)
    pr["check_orientation"] = chk
    return pr


def feature_vector(ptr, idx, n, side):
    """(node ids ascending, CSR) of one side: common bucket j -> node 3 j + 1; the features outside every common bucket go, a few at a
    time, into nodes 3 j + (0 for side 1, 2 for side 2) that the other side does not have."""
    ptr = np.asarray(ptr, np.int64); idx = np.asarray(idx, np.int64)
    stray = np.setdiff1d(np.arange(n), idx)
    nodes, lists = [], []
    nb = len(ptr) - 1
    chunks = np.array_split(stray, max(1, min(len(stray), nb + 1))) if len(stray) else []
    for j in range(nb + 1):
        if j < len(chunks) and len(chunks[j]):
            nodes.append(3 * j + (0 if side == 1 else 2)); lists.append(chunks[j])
        if j < nb:
            nodes.append(3 * j + 1); lists.append(idx[ptr[j]:ptr[j + 1]])
    order = np.argsort(nodes)
    nodes = np.array(nodes, np.int32)[order]; lists = [lists[k] for k in order]
    p = np.zeros(len(lists) + 1, np.int32); p[1:] = np.cumsum([len(l) for l in lists])
    return nodes, p, (np.concatenate(lists).astype(np.int32) if lists else np.zeros(0, np.int32))


def run_ref(L, pr):
    from airdos_b200.capi import KP_DTYPE
    k1 = np.ascontiguousarray(pr["kps1"], KP_DTYPE); k2 = np.ascontiguousarray(pr["kps2"], KP_DTYPE)
    d1 = np.ascontiguousarray(pr["desc1"], np.uint8); d2 = np.ascontiguousarray(pr["desc2"], np.uint8)
    n1, n2 = len(k1), len(k2)
    nd1, p1, i1 = feature_vector(pr["b_ptr1"], pr["b_idx1"], n1, 1)
    nd2, p2, i2 = feature_vector(pr["b_ptr2"], pr["b_idx2"], n2, 2)
    f1 = np.ascontiguousarray(pr["flags1"], np.uint8)
    if pr["mode"] == 0:
        m = np.zeros(n2, np.int32)
        L.ref_search_by_bow.restype = C.c_int
        L.ref_search_by_bow.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [
            C.c_int, C.c_float, C.c_int, C.c_void_p]
        n = L.ref_search_by_bow(P(k1), P(d1), P(f1), n1, P(k2), P(d2), n2, P(nd1), P(p1), P(i1), len(nd1), P(nd2), P(p2), P(i2), len(nd2),
                                float(pr["nn_ratio"]), int(pr["check_orientation"]), P(m))
        return int(n), m
    f2 = np.ascontiguousarray(pr["flags2"], np.uint8)
    has1 = (1 - f1).astype(np.uint8); has2 = (1 - f2).astype(np.uint8)          # flags = "not triangulated yet"
    ur1 = np.ascontiguousarray(pr["u_right1"], np.float32); ur2 = np.ascontiguousarray(pr["u_right2"], np.float32)
    F = np.ascontiguousarray(pr["f12"], np.float32)
    ex, ey = pr["epipole"]
    # key-frame 2 at the origin with unit intrinsics and key-frame 1's centre at (ex, ey, 1): the reference computes exactly this epipole
    ow1 = np.array([ex, ey, 1.0], np.float32); r2w = np.eye(3, dtype=np.float32).ravel(); t2w = np.zeros(3, np.float32)
    cam2 = np.array([1.0, 1.0, 0.0, 0.0], np.float32)
    sf = np.ascontiguousarray(pr["scale_factors2"], np.float32); sg = np.ascontiguousarray(pr["level_sigma2_2"], np.float32)
    m = np.zeros(n1, np.int32)
    L.ref_search_for_triangulation.restype = C.c_int
    L.ref_search_for_triangulation.argtypes = ([C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 +
                                               [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p])
    n = L.ref_search_for_triangulation(P(k1), P(ur1), P(d1), P(has1), n1, P(k2), P(ur2), P(d2), P(has2), n2, P(nd1), P(p1), P(i1), len(nd1), P(nd2), P(p2),
                                       P(i2), len(nd2), P(F), P(ow1), P(r2w), P(t2w), P(cam2), P(sf), P(sg), len(sf), float(pr["nn_ratio"]),
                                       int(pr["check_orientation"]), 0, P(m))
    return int(n), m


def problem_crc(pr):
    c = 0
    for k in sorted(pr):
        if isinstance(pr[k], np.ndarray):
            c = zlib.crc32(np.ascontiguousarray(pr[k]).tobytes(), c)
    return c


def main():
    L = C.CDLL(LIB)
    out = {"cases": np.array(CASES, np.int32)}
    for i, case in enumerate(CASES):
        pr = problem(case)
        n, m = run_ref(L, pr)
        out[f"c{i}_n"] = np.int32(n); out[f"c{i}_match"] = m; out[f"c{i}_crc"] = np.int64(problem_crc(pr))
        print(f"case {i} (mode {case[1]}): {n} matches")
    path = os.path.join(ROOT, "tests", "golden", "bow_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
