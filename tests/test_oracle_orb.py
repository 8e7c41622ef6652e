"""CPU tests of the ORB oracle: golden fixtures (generated from cv2 4.13 by oracle/crosscheck_cv2.py),
primitive known answers, and the quad-tree against an independently written level-synchronous
formulation (the one the CUDA kernel uses)."""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "orb_*.npz"))), ids=os.path.basename)
def test_oracle_matches_golden(oracle_mod, path):
    g = np.load(path, allow_pickle=False)
    nf, nl, ini, mn = [int(v) for v in g["params"]]
    mask = g["mask"] if g["mask"].size else None
    o = oracle_mod.orb_extract(g["image"], mask, nf, float(g["scale"]), nl, ini, mn, want_pyramid=True)
    assert len(o["kps"]) == len(g["kps"])
    assert o["kps"].tobytes() == g["kps"].tobytes()          # bit-exact incl. float angle patterns
    assert (o["desc"] == g["desc"]).all()
    assert [p.shape for p in o["pyramid"]] == [tuple(s) for s in g["pyramid_sizes"]]
    assert [int(p.astype(np.int64).sum()) for p in o["pyramid"]] == [int(v) for v in g["pyramid_sums"]]


def test_params_match_survey_table(oracle_mod):
    # SURVEY.md appendix A.5 (computed with the reference's float32 rules, src/ORBextractor.cc:411-445)
    p = oracle_mod.orb_params(1000, 1.2, 8, 640, 480)
    assert list(p["w"]) == [640, 533, 444, 370, 309, 257, 214, 179]
    assert list(p["h"]) == [480, 400, 333, 278, 231, 193, 161, 134]
    assert list(p["quota"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(oracle_mod.orb_params(2000, 1.2, 8, 640, 480)["quota"]) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(p["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(oracle_mod.orb_params(1500, 1.2, 8, 640, 360)["quota"]) == [326, 271, 226, 189, 157, 131, 109, 91]


def test_primitive_known_answers(oracle_mod):
    o = oracle_mod
    # REFLECT_101: gfedcb|abcdefgh|gfedcba
    row = np.arange(8, dtype=np.uint8)[None].repeat(3, 0)
    b = o.border101(row, 2)
    assert list(b[2]) == [2, 1, 0, 1, 2, 3, 4, 5, 6, 7, 6, 5]
    # blur of a constant image is the constant; kernel sums to 256
    assert (o.blur7(np.full((20, 20), 93, np.uint8)) == 93).all()
    # resize to the same size is the identity
    img = np.random.default_rng(0).integers(0, 256, (30, 40), dtype=np.uint8)
    assert (o.resize(img, 40, 30) == img).all()
    # fastAtan2 special values (cv::fastAtan2 returns exactly these)
    assert o.fast_atan2(0, 0) == 0.0
    assert abs(o.fast_atan2(1, 1) - 45.0) < 0.02 and abs(o.fast_atan2(-1, 0) - 270.0) < 0.02
    # erode: a single zero pixel grows to a 10x10 block anchored at (5,5)
    m = np.full((30, 30), 255, np.uint8); m[15, 15] = 0
    e = o.erode10(m)
    ys, xs = np.nonzero(e == 0)
    assert (ys.min(), ys.max(), xs.min(), xs.max()) == (11, 20, 11, 20)
    # FAST: flat image has no corners, an isolated bright dot is not a FAST-9 corner, a bright square corner is
    assert len(o.fast(np.full((20, 20), 50, np.uint8), 7)) == 0
    noise = np.random.default_rng(2).integers(0, 256, (40, 40), dtype=np.uint8)
    k = o.fast(noise, 20)
    assert len(k) >= 1 and all(s >= 20 for s in k[:, 2])          # response = best - 1 >= threshold
    assert k[:, 0].min() >= 3 and k[:, 0].max() <= 36 and k[:, 1].min() >= 3 and k[:, 1].max() <= 36
    assert (np.lexsort((k[:, 0], k[:, 1])) == np.arange(len(k))).all()   # row-major output order


def _quadtree_level_sync(cand, min_x, max_x, min_y, max_y, N):
    """Independent formulation of DistributeOctTree (src/ORBextractor.cc:541-765): nodes are numbered
    in creation order; the reference's list is always 'newest first' (roots, ascending, at the
    tail); each pass splits the splittable nodes in list order, or by (count, seq) descending
    with an early stop once the next pass could overshoot N."""
    n_ini = int(np.round(np.float32(max_x - min_x) / np.float32(max_y - min_y)))
    hx = np.float32(max_x - min_x) / np.float32(n_ini)
    xs, ys, rs = cand[:, 0], cand[:, 1], cand[:, 2]
    root = (xs / hx).astype(np.int32)
    nodes = {}   # seq -> (ulx, uly, brx, bry, key indices)
    for i in range(n_ini):
        idx = np.nonzero(root == i)[0]
        if len(idx):
            nodes[i] = (int(np.float32(hx) * np.float32(i)), 0, int(np.float32(hx) * np.float32(i + 1)), max_y - min_y, idx)
    next_seq = n_ini
    active = [s for s in sorted(nodes) if len(nodes[s][4]) > 1]
    mode, first = 0, True

    def split(s):
        nonlocal next_seq
        ulx, uly, brx, bry, idx = nodes.pop(s)
        mx, my = ulx + (brx - ulx + 1) // 2, uly + (bry - uly + 1) // 2
        left, top = xs[idx] < mx, ys[idx] < my
        kids = []
        for box, sel in (((ulx, uly, mx, my), left & top), ((mx, uly, brx, my), ~left & top),
                         ((ulx, my, mx, bry), left & ~top), ((mx, my, brx, bry), ~left & ~top)):
            if sel.any():
                nodes[next_seq] = (*box, idx[sel]); kids.append(next_seq); next_seq += 1
        return kids

    while active:
        size = len(nodes)
        order = (active if first else active[::-1]) if mode == 0 else \
            sorted(active, key=lambda s: (len(nodes[s][4]), s), reverse=True)
        new_active = []
        for s in order:
            new_active += [k for k in split(s) if len(nodes[k][4]) > 1]
            if mode == 1 and len(nodes) >= N:
                break
        first = False
        if len(nodes) >= N or len(nodes) == size:
            break
        if mode == 0 and len(nodes) + 3 * len(new_active) > N:
            mode = 1
        active = sorted(new_active)
    out = []
    for s in sorted(nodes, key=lambda q: q if q >= n_ini else n_ini - 1 - q, reverse=True):
        idx = nodes[s][4]
        out.append(cand[idx[np.argmax(rs[idx])]])     # first maximum wins
    return np.array(out, np.float32).reshape(-1, 3)


@pytest.mark.parametrize("seed,w,h,n,N", [(0, 608, 448, 6000, 434), (1, 608, 448, 900, 217), (2, 608, 328, 3000, 326),
                                            (3, 1200, 300, 5000, 500), (4, 147, 102, 300, 60), (5, 608, 448, 50, 400),
                                            (6, 608, 448, 3, 5), (7, 300, 300, 4000, 3000)])
def test_quadtree_vs_level_synchronous_formulation(oracle_mod, seed, w, h, n, N):
    rng = np.random.default_rng(seed)
    pts = rng.permutation(w * h)[:n]
    cand = np.stack([pts % w, pts // w, rng.integers(7, 60, n)], 1).astype(np.float32)
    cand = cand[np.lexsort((cand[:, 0], cand[:, 1]))]   # any order works; ties in response exercise 'first wins'
    ref = oracle_mod.distribute(cand, 16, 16 + w, 16, 16 + h, N)
    got = _quadtree_level_sync(cand, 16, 16 + w, 16, 16 + h, N)
    assert ref.shape == got.shape and (ref == got).all()
    assert len(ref) >= min(N, 1)


def test_extract_edge_cases(oracle_mod):
    o = oracle_mod
    flat = o.orb_extract(np.full((480, 640), 128, np.uint8), None, 1000, 1.2, 8, 12, 7)
    assert len(flat["kps"]) == 0
    # all-zero mask rejects everything; all-255 mask equals no mask
    from airdos_b200 import synth
    img = synth.make_stereo_pair(7, 320, 240)[0]
    assert len(o.orb_extract(img, np.zeros_like(img), 500, 1.2, 8, 20, 7)["kps"]) == 0
    a = o.orb_extract(img, None, 500, 1.2, 8, 20, 7)
    b = o.orb_extract(img, np.full_like(img, 255), 500, 1.2, 8, 20, 7)
    assert a["kps"].tobytes() == b["kps"].tobytes() and (a["desc"] == b["desc"]).all()
    # key-points respect the 19-px edge threshold on their level and the per-level quota (+2 overshoot)
    p = o.orb_params(500, 1.2, 8, 320, 240)
    for l in range(8):
        k = a["kps"][a["kps"]["octave"] == l]
        assert len(k) <= p["quota"][l] + 2
        if len(k):
            x, y = k["x"] / p["scale"][l], k["y"] / p["scale"][l]
            assert x.min() >= 18.99 and y.min() >= 18.99 and x.max() <= p["w"][l] - 19.99 and y.max() <= p["h"][l] - 19.99


def test_portrait_shapes_with_zero_roots_are_refused(oracle_mod):
    """DistributeOctTree takes nIni = round(width / height) root nodes (src/ORBextractor.cc:545-549); for a region taller than
    twice its width that is 0 and the reference divides by it -- undefined.  Oracle and library refuse the shape (ADB_ERR_INVALID)."""
    import numpy as np
    import pytest
    img = np.random.default_rng(0).integers(0, 256, (640, 200), dtype=np.uint8)
    with pytest.raises(ValueError):
        oracle_mod.orb_extract(img, None, 500, 1.2, 8, 20, 7)
    ok = np.random.default_rng(0).integers(0, 256, (400, 300), dtype=np.uint8)     # 268 / 368 rounds to 1: fine
    assert len(oracle_mod.orb_extract(ok, None, 500, 1.2, 8, 20, 7)["kps"]) > 0
