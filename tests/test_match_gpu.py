"""GPU parity tests of the Hamming kernels and the stereo matcher against the oracle (bit-exact,
float outputs compared by bit pattern)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def adb():
    import airdos_b200
    return airdos_b200


def test_best2_all_pairs_lists_ties_and_empty(adb, oracle_mod):
    rng = np.random.default_rng(0)
    m = adb.ORBmatcher()
    q = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (2111, 32), dtype=np.uint8)
    t[100:140] = t[7]            # duplicates: first in list order must win
    q[:20] = t[7]
    for a, b in zip(m.best2(q, t), oracle_mod.best2(q, t)):
        assert (a == b).all()
    lens = rng.integers(0, 60, len(q)); lens[::7] = 0
    off = np.zeros(len(q) + 1, np.int32); off[1:] = np.cumsum(lens)
    idx = rng.integers(0, len(t), off[-1]).astype(np.int32)
    idx[rng.integers(0, len(idx), 500)] = 7
    got, ref = m.best2(q, t, off, idx), oracle_mod.best2(q, t, off, idx)
    for a, b in zip(got, ref):
        assert (a == b).all()
    assert (got[0][::7] == -1).all() and (got[1][::7] == 256).all() and (got[2][::7] == 256).all()
    # no targets at all
    bi, bd, sd = m.best2(q[:5], np.zeros((0, 32), np.uint8))
    assert (bi == -1).all() and (bd == 256).all() and (sd == 256).all()
    # linearity-style property at size: distance to self is 0 and symmetric argmin of a permutation
    perm = rng.permutation(len(t))
    bi, bd, _ = m.best2(t[perm], t)
    assert (bd == 0).all() and (t[bi] == t[perm]).all()
    # candidate lists index the train descriptors on the device: an index outside [0, nt) or a non-monotone offset is refused
    bad_idx = idx.copy(); bad_idx[11] = len(t)
    bad_off = off.copy(); bad_off[5] = bad_off[6] + 1
    for o, i in ((off, bad_idx), (bad_off, idx)):
        with pytest.raises(adb.AdbError) as e:
            m.best2(q, t, o, i)
        assert e.value.status == 1
    m.close()


@pytest.mark.parametrize("nf,w,h", [(2000, 640, 480), (1000, 640, 480), (1500, 640, 360)])
def test_stereo_matches_oracle(adb, oracle_mod, nf, w, h):
    from airdos_b200 import synth
    F = 3
    pairs = synth.make_stereo_batch(F, w, h, start=40)
    exL = adb.ORBextractor(nf, 1.2, 8, 12, 7, w, h, max_batch=F)
    exR = adb.ORBextractor(nf, 1.2, 8, 12, 7, w, h, max_batch=F)
    kl, dl, cl = exL.extract_batch(pairs[:, 0])
    kr, dr, cr = exR.extract_batch(pairs[:, 1])
    mbf = synth.BF; mb = mbf / synth.FX
    ur, dp, bi, bd = adb.compute_stereo_matches(exL, exR, F, mb, mbf)
    sc = np.array(exL.GetScaleFactors(), np.float32)
    for f in range(F):
        nl, nr = cl[f], cr[f]
        o = oracle_mod.stereo_match(kl[f, :nl], dl[f, :nl], kr[f, :nr], dr[f, :nr], exL.pyramid(f), exR.pyramid(f), sc, mb, mbf)
        assert (o[1] > 0).sum() > 100
        for got, ref in zip((ur, dp, bi, bd), o):
            assert (got[f, :nl].view(np.uint32) == ref.view(np.uint32)).all()
    exL.close(); exR.close()


def test_stereo_no_matches_is_a_noop(adb):
    """Unrelated left / right images: few or no matches; the empty median cut must not crash (D.8)."""
    from airdos_b200 import synth
    L = synth.make_stereo_pair(1)[0]
    R = np.full_like(L, 90)
    exL = adb.ORBextractor(1000, 1.2, 8, 12, 7); exR = adb.ORBextractor(1000, 1.2, 8, 12, 7)
    exL.extract_batch(L[None]); exR.extract_batch(R[None])
    ur, dp, bi, bd = adb.compute_stereo_matches(exL, exR, 1, 0.25, 193.137)
    assert (ur == -1).all() and (dp == -1).all() and (bi == -1).all()
    exL.close(); exR.close()


def test_distinctive_descriptors_match_oracle(adb, oracle_mod):
    """MapPoint::ComputeDistinctiveDescriptors batched: identical winner index and descriptor."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from test_oracle_match import _distinct_case
    m = adb.ORBmatcher()
    for seed in (2, 3):
        desc, ptr = _distinct_case(np.random.default_rng(seed), 500)
        bi, bd = adb.compute_distinctive_descriptors(m, desc, ptr)
        ref = oracle_mod.distinctive(desc, ptr)
        assert (bi == ref).all()
        ok = ref >= 0
        assert (bd[ok] == desc[ptr[:-1][ok] + ref[ok]]).all()
    with pytest.raises(adb.AdbError):                       # more observations than ADB_MAX_OBSERVATIONS: refused, never truncated
        adb.compute_distinctive_descriptors(m, np.zeros((129, 32), np.uint8), np.array([0, 129], np.int32))
    m.close()


@pytest.mark.parametrize("n,masked", [(40, False), (40, True), (3, False)])
def test_stereo_frames_batch_equals_the_three_calls(n, masked):
    """adb_stereo_frames_batch (the stereo Frame constructor's hot part as one chunk pipeline, src/Frame.cc:80-100) returns exactly what
    adb_orb_extract_batch x 2 + adb_stereo_match return: key-points, descriptors, counts, uRight / depth bit patterns, best index /
    distance -- chunked (40 frames) and below the chunking threshold (3 frames)."""
    import airdos_b200 as adb
    from airdos_b200 import synth
    w, h = 640, 480
    pairs = [synth.make_stereo_pair(300 + i, w, h) for i in range(4)]
    L = np.stack([pairs[i % 4][0] for i in range(n)]); R = np.stack([pairs[i % 4][1] for i in range(n)])
    mL = np.stack([synth.make_mask(310 + i % 3, w, h, 3) for i in range(n)]) if masked else None
    mR = np.stack([synth.make_mask(320 + i % 3, w, h, 3) for i in range(n)]) if masked else None
    mbf = synth.BF; mb = mbf / synth.FX
    exL = adb.ORBextractor(1500, 1.2, 8, 12, 7, w, h, max_batch=n); exR = adb.ORBextractor(1500, 1.2, 8, 12, 7, w, h, max_batch=n)
    (kl, dl, cl), (kr, dr, cr), (ur, dp, bi, bd) = adb.stereo_frames_batch(exL, exR, L, R, mb, mbf, mL, mR)
    kl, dl, cl, kr, dr, cr, ur, dp, bi, bd = [a.copy() for a in (kl, dl, cl, kr, dr, cr, ur, dp, bi, bd)]
    a = adb.ORBextractor(1500, 1.2, 8, 12, 7, w, h, max_batch=n); b = adb.ORBextractor(1500, 1.2, 8, 12, 7, w, h, max_batch=n)
    kl2, dl2, cl2 = a.extract_batch(L, mL); kr2, dr2, cr2 = b.extract_batch(R, mR)
    ur2, dp2, bi2, bd2 = adb.compute_stereo_matches(a, b, n, mb, mbf)
    assert (cl == cl2).all() and (cr == cr2).all() and cl.min() > 0
    for f in range(n):
        assert kl[f, :cl[f]].tobytes() == kl2[f, :cl[f]].tobytes() and (dl[f, :cl[f]] == dl2[f, :cl[f]]).all(), f
        assert kr[f, :cr[f]].tobytes() == kr2[f, :cr[f]].tobytes() and (dr[f, :cr[f]] == dr2[f, :cr[f]]).all(), f
        for x, y in ((ur, ur2), (dp, dp2), (bi, bi2), (bd, bd2)):
            assert (x[f, :cl[f]].view(np.uint32) == y[f, :cl[f]].view(np.uint32)).all(), f
    assert (dp > 0).sum() > 100 * n
    for e in (exL, exR, a, b):
        e.close()


@pytest.mark.parametrize("ci", [0, 2])
def test_cuda_stereo_equals_the_reference_function(ci):
    """CUDA extraction + stereo matcher on the images of tests/golden/stereo_ref.npz against mvuRight / mvDepth computed by the
    reference's own Frame::ComputeStereoMatches (src/Frame.cc:829-1003, compiled from /root/reference: oracle/ref_match.cpp)."""
    import os
    import airdos_b200 as adb
    from airdos_b200 import synth
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stereo_ref.npz"))
    seed, w, h, nf, ini, mn = [int(v) for v in g[f"c{ci}_params"]]
    il, ir = g[f"c{ci}_left"], g[f"c{ci}_right"]
    exL = adb.ORBextractor(nf, 1.2, 8, ini, mn, w, h); exR = adb.ORBextractor(nf, 1.2, 8, ini, mn, w, h)
    _, _, cl = exL.extract_batch(il[None]); exR.extract_batch(ir[None])
    mbf = synth.BF; mb = mbf / synth.FX
    ur, dp, _, _ = adb.compute_stereo_matches(exL, exR, 1, mb, mbf)
    n = int(cl[0])
    assert n == len(g[f"c{ci}_u_right"])
    assert (ur[0, :n].view(np.uint32) == g[f"c{ci}_u_right"].view(np.uint32)).all()
    assert (dp[0, :n].view(np.uint32) == g[f"c{ci}_depth"].view(np.uint32)).all()
    exL.close(); exR.close()


def test_cuda_distinctive_descriptor_equals_the_reference_function(adb):
    """adb_distinctive_descriptors against the descriptors the reference's own MapPoint::ComputeDistinctiveDescriptors
    (src/MapPoint.cc:245-310, compiled from /root/reference) left in mDescriptor: tests/golden/stereo_ref.npz."""
    import importlib.util
    import os
    import zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ref_match_golden", os.path.join(root, "oracle", "gen_ref_match_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    gold = np.load(os.path.join(root, "tests", "golden", "stereo_ref.npz"))
    desc, ptr = g.distinct_case(np.random.default_rng(20261018))
    assert [zlib.crc32(desc.tobytes()), zlib.crc32(ptr.tobytes())] == [int(v) for v in gold["distinct_crc"]]
    m = adb.ORBmatcher()
    bi, bd = adb.compute_distinctive_descriptors(m, desc, ptr)
    has = ptr[1:] > ptr[:-1]
    assert (bd[has] == gold["distinct_chosen"][has]).all() and (bi[~has] == -1).all()
    m.close()
