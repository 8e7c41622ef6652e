"""GPU parity tests of the guided window searches (airdos_b200/csrc/search.cu) against the oracle: final
CurrentFrame.mvpMapPoints, nmatches, bestIdx / bestDist of every query -- all bit-exact, through the C-ABI."""
import numpy as np
import pytest

from airdos_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def adb():
    import airdos_b200
    return airdos_b200


def _same(got, ref):
    n, km, bi, bd = got
    rn, rkm, rbi, rbd = ref[:4]
    assert n == rn, (n, rn)
    assert (km == rkm).all(), np.nonzero(km != rkm)[0][:10]
    assert (bi == rbi).all() and (bd == rbd).all()


@pytest.mark.parametrize("seed,n_kp,n_q,dz", [(0, 2000, 1500, 0.02), (1, 2000, 2000, 0.6), (2, 1200, 1800, -0.6), (3, 4000, 3000, 0.02),
                                               (4, 300, 50, 0.02)])
def test_last_frame_search_matches_oracle(adb, oracle_mod, seed, n_kp, n_q, dz):
    pr = synth.make_tracking_problem(seed, n_kp=n_kp, n_q=n_q, last_dz=dz, dup_frac=0.2)
    m = adb.ORBmatcher(0.9, True)
    ref = oracle_mod.search_by_projection(pr)
    _same(m.SearchByProjection(pr), ref)
    assert ref[0] > 20
    # mono flag (no forward / backward rule) and the wider second pass of TrackWithMotionModel (2 * th)
    pr2 = dict(pr); pr2["mono"] = 1; pr2["th"] = 2 * pr["th"]
    _same(m.SearchByProjection(pr2), oracle_mod.search_by_projection(pr2))
    m.close()


def test_map_point_search_matches_oracle(adb, oracle_mod):
    m = adb.ORBmatcher(0.8, True)
    for seed in (10, 11):
        pr = synth.make_tracking_problem(seed, n_kp=2000, n_q=2500, dup_frac=0.3)
        proj = oracle_mod.search_by_projection(pr)[4]
        pm = synth.tracking_problem_as_map_points(pr, proj, nn_ratio=0.8)
        ref = oracle_mod.search_by_projection(pm)
        _same(m.SearchByProjection(pm), ref)
        assert ref[0] > 100
    m.close()


def test_search_batch_and_edge_cases(adb, oracle_mod):
    m = adb.ORBmatcher(0.9, True)
    probs = [synth.make_tracking_problem(20 + i, n_kp=500 + 300 * i, n_q=400 + 200 * i) for i in range(5)]
    # everything closed on entry; no valid query; no orientation check
    p_closed = dict(probs[0]); p_closed["taken"] = np.ones(len(p_closed["kps"]), np.uint8)
    p_none = dict(probs[1]); p_none["q_flags"] = np.zeros_like(p_none["q_flags"])
    p_noori = dict(probs[2]); p_noori["check_orientation"] = 0
    p_notaken = dict(probs[3]); p_notaken["taken"] = None
    probs += [p_closed, p_none, p_noori, p_notaken]
    got = m.search_by_projection(probs)
    for g, pr in zip(got, probs):
        _same(g, oracle_mod.search_by_projection(pr))
    assert got[5][0] == 0 and got[6][0] == 0
    assert m.search_last_ms() > 0
    # single calls give the same as the batch
    _same(m.SearchByProjection(probs[4]), got[4])
    m.close()


@pytest.mark.parametrize("seed,th", [(30, 1.0), (31, 5.0)])
def test_local_map_search_with_frustum_on_device(adb, oracle_mod, seed, th):
    """Tracking::SearchLocalPoints: isInFrustum + PredictScale + SearchByProjection(F, vpMapPoints, th) behind one call."""
    pr = synth.make_tracking_problem(seed, n_kp=2000, n_q=3000, dup_frac=0.25)
    pm = synth.tracking_problem_as_local_map(pr, seed=seed, th=th)
    m = adb.ORBmatcher(0.8, True)
    ref = oracle_mod.search_by_projection(pm)
    got = m.SearchByProjection(pm)
    _same(got[:4], ref)
    assert (got[5] == ref[4]["q_level"]).all()
    assert (got[4].view(np.uint32) == ref[4]["q_track"].view(np.uint32)).all()     # float outputs by bit pattern
    assert ref[0] > 100 and (ref[4]["q_level"] >= 0).sum() > 1000
    m.close()


def test_fuse_candidate_search_on_device(adb, oracle_mod):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th): projection, visibility tests, window search with the chi2 gate, TH_LOW."""
    m = adb.ORBmatcher(0.6, True)
    for seed, th in ((40, 3.0), (41, 4.0)):
        pr = synth.make_tracking_problem(seed, n_kp=2000, n_q=3000, dup_frac=0.3)
        pf = synth.tracking_problem_as_fuse(pr, seed=seed, th=th)
        ref = oracle_mod.search_by_projection(pf)
        got = m.SearchByProjection(pf)
        _same(got[:4], ref)
        assert ref[0] > 200
    m.close()


def test_bad_level_numbers_and_bucket_indices_are_refused(adb):
    """Numbers from the shim that index device tables are range-checked on the host (ADB_ERR_INVALID, nothing launched): the last frame's
    octaves and the KeyFrame's octaves of Fuse index the per-level tables; the vocabulary bucket lists index the key-point arrays."""
    m = adb.ORBmatcher(0.9, True)
    pr = synth.make_tracking_problem(60, n_kp=600, n_q=500)
    for bad in (-1, len(pr["scale_factors"])):
        q = dict(pr); q["last_octave"] = np.array(pr["last_octave"], np.int32).copy(); q["last_octave"][7] = bad
        with pytest.raises(adb.AdbError) as e:
            m.SearchByProjection(q)
        assert e.value.status == 1
    pf = synth.tracking_problem_as_fuse(synth.make_tracking_problem(61, n_kp=600, n_q=800), seed=61, th=3.0)
    pf["kps"] = pf["kps"].copy(); pf["kps"]["octave"][3] = len(pf["scale_factors"])
    with pytest.raises(adb.AdbError) as e:
        m.search_by_projection([pf])
    assert e.value.status == 1
    for mode, key, val in ((0, "b_idx1", 10 ** 6), (0, "b_idx2", -1), (1, "b_idx2", 10 ** 6)):
        pb = synth.make_bow_problem(62, mode, n1=700, n2=800, n_nodes=80)
        pb[key] = np.array(pb[key], np.int32).copy(); pb[key][5] = val
        with pytest.raises(adb.AdbError) as e:
            m.search_by_bow([pb])
        assert e.value.status == 1
    pb = synth.make_bow_problem(63, 0, n1=700, n2=800, n_nodes=80)
    pb["b_ptr1"] = np.array(pb["b_ptr1"], np.int32).copy(); pb["b_ptr1"][3] = pb["b_ptr1"][4] + 1      # not monotone
    with pytest.raises(adb.AdbError) as e:
        m.search_by_bow([pb])
    assert e.value.status == 1
    # the matcher is still usable afterwards
    assert m.SearchByProjection(pr)[0] > 0
    m.close()


@pytest.mark.parametrize("mode", [0, 1], ids=["SearchByBoW", "SearchForTriangulation"])
def test_bow_searches_on_device(adb, oracle_mod, mode):
    m = adb.ORBmatcher(0.7, True)
    probs = [synth.make_bow_problem(50 + s, mode, n1=1500 + 300 * s, n2=2000, n_nodes=nn) for s, nn in ((0, 400), (1, 60), (2, 1500))]
    p_noori = dict(probs[0]); p_noori["check_orientation"] = 0
    p_none = dict(probs[1]); p_none["flags1"] = np.zeros_like(p_none["flags1"])
    probs += [p_noori, p_none]
    got = m.search_by_bow(probs)
    for g, pr in zip(got, probs):
        n, match = oracle_mod.search_by_bow(pr)
        assert g[0] == n and (g[1] == match).all(), (g[0], n)
    assert got[0][0] > 100 and got[4][0] == 0
    # one problem per call gives the same
    one = m.search_by_bow([probs[2]])[0]
    assert one[0] == got[2][0] and (one[1] == got[2][1]).all()
    m.close()


@pytest.mark.parametrize("kind,i", [("last", 0), ("last", 1), ("last", 2), ("last", 3), ("last", 4), ("map", 0), ("map", 1), ("map", 2)])
def test_cuda_search_equals_the_reference_function(adb, kind, i):
    """CUDA searches against tests/golden/search_ref.npz: nmatches and final mvpMapPoints computed by the reference's own
    ORBmatcher::SearchByProjection functions (src/ORBmatcher.cc:45-129, 1328-1470; compiled from /root/reference, oracle/ref_match.cpp)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ref_search_golden", os.path.join(root, "oracle", "gen_ref_search_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    gold = np.load(os.path.join(root, "tests", "golden", "search_ref.npz"))
    pr = g.last_problem(g.LAST_CASES[i]) if kind == "last" else g.map_problem(g.MAP_CASES[i])
    assert g.problem_crc(pr) == int(gold[f"{kind}{i}_crc"])
    m = adb.ORBmatcher(0.9 if kind == "last" else float(pr["nn_ratio"]), True)
    n, km, _, _ = m.SearchByProjection(pr)
    assert n == int(gold[f"{kind}{i}_n"])
    assert (np.where(km >= 0, km, -1) == gold[f"{kind}{i}_kp_match"]).all()
    m.close()


def test_cuda_bow_searches_equal_the_reference_functions(adb):
    """CUDA vocabulary-bucket searches against tests/golden/bow_ref.npz: nmatches and the match tables computed by the reference's
    own ORBmatcher::SearchByBoW / SearchForTriangulation (src/ORBmatcher.cc:159-288, 657-823; compiled from /root/reference)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ref_bow_golden", os.path.join(root, "oracle", "gen_ref_bow_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    gold = np.load(os.path.join(root, "tests", "golden", "bow_ref.npz"))
    probs = [g.problem(c) for c in g.CASES]
    for i, pr in enumerate(probs):
        assert g.problem_crc(pr) == int(gold[f"c{i}_crc"])
    m = adb.ORBmatcher(0.7, True)
    got = m.search_by_bow(probs)
    for i, (n, match) in enumerate(got):
        assert n == int(gold[f"c{i}_n"]) and (match == gold[f"c{i}_match"]).all(), i
    m.close()


@pytest.mark.parametrize("i", [0, 1, 2])
def test_cuda_local_map_search_equals_the_reference_functions(adb, i):
    """Tracking::SearchLocalPoints' hot part on the device against tests/golden/search_ref.npz: mbTrackInView, mTrackProjX / Y / XR /
    mTrackViewCos (bit patterns), mnTrackScaleLevel, final mvpMapPoints and nmatches computed by the reference's own Frame::isInFrustum
    (src/Frame.cc:587-643), MapPoint::PredictScale (src/MapPoint.cc:405-420) and SearchByProjection(F, vpMapPoints, th)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ref_search_golden", os.path.join(root, "oracle", "gen_ref_search_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    gold = np.load(os.path.join(root, "tests", "golden", "search_ref.npz"))
    pm = g.local_problem(g.LOCAL_CASES[i], gold[f"local{i}_ow"])
    assert g.problem_crc(pm) == int(gold[f"local{i}_crc"])
    m = adb.ORBmatcher(float(pm["nn_ratio"]), True)
    got = m.SearchByProjection(pm)
    n, km, track, level = got[0], got[1], got[4], got[5]
    assert n == int(gold[f"local{i}_n"])
    assert (np.where(km >= 0, km, -1) == gold[f"local{i}_kp_match"]).all()
    assert ((level >= 0) == (gold[f"local{i}_in_view"] > 0)).all()
    assert (level == gold[f"local{i}_level"]).all()
    assert (track.view(np.uint32) == gold[f"local{i}_track"].view(np.uint32)).all()
    m.close()


@pytest.mark.parametrize("i", [0, 1, 2])
def test_cuda_fuse_search_equals_the_reference_function(adb, i):
    """The fuse mode of adb_search_by_projection against tests/golden/search_ref.npz: nFused and the key-point every map point was
    fused with by the reference's own ORBmatcher::Fuse (src/ORBmatcher.cc:825-975, compiled from /root/reference)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ref_search_golden", os.path.join(root, "oracle", "gen_ref_search_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    gold = np.load(os.path.join(root, "tests", "golden", "search_ref.npz"))
    pf = g.fuse_problem(g.FUSE_CASES[i], gold[f"fuse{i}_ow"])
    assert g.problem_crc(pf) == int(gold[f"fuse{i}_crc"])
    m = adb.ORBmatcher(0.6, True)
    got = m.SearchByProjection(pf)
    n, bi, bd = got[0], got[2], got[3]
    assert n == int(gold[f"fuse{i}_n"])
    assert (np.where(bd <= 50, bi, -1) == gold[f"fuse{i}_fused_with"]).all()
    m.close()
