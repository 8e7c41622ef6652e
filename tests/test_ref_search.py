"""Pins the tracking searches of oracle/match_oracle.cpp to the LITERAL reference: tests/golden/search_ref.npz holds nmatches and
the final CurrentFrame.mvpMapPoints of the reference's own ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)
(src/ORBmatcher.cc:1328-1470) and SearchByProjection(Frame&, vector<MapPoint*>&, th) (:45-129) -- with Frame::AssignFeaturesToGrid /
GetFeaturesInArea / PosInGrid (src/Frame.cc:534-549, 645-712), ComputeThreeMaxima and DescriptorDistance -- the function bodies
compiled from /root/reference by `make -C oracle ref` (oracle/ref_match.cpp) and run by oracle/gen_ref_search_golden.py on seeded
problems (forward / backward / lateral motion, mono flag, with and without the rotation check; three radii / ratios of the local-map
variant).  Bar: identical."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "search_ref.npz")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_match.so")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_ref_search_golden", os.path.join(ROOT, "oracle", "gen_ref_search_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    return g


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _problem(g, gold, kind, i):
    pr = g.last_problem(g.LAST_CASES[i]) if kind == "last" else g.map_problem(g.MAP_CASES[i])
    assert g.problem_crc(pr) == int(gold[f"{kind}{i}_crc"]), "synthetic generator drifted: regenerate the fixture"
    return pr


@pytest.mark.parametrize("kind,i", [("last", i) for i in range(5)] + [("map", i) for i in range(3)])
def test_oracle_search_equals_the_reference_function(gold, oracle_mod, kind, i):
    g = _gen()
    pr = _problem(g, gold, kind, i)
    n, km, _, _, _ = oracle_mod.search_by_projection(pr)
    assert n == int(gold[f"{kind}{i}_n"]) and n > 400
    assert (np.where(km >= 0, km, -1) == gold[f"{kind}{i}_kp_match"]).all()     # the oracle's -2 (cleared by the rotation check) is NULL there


@pytest.mark.parametrize("i", [0, 1, 2])
def test_oracle_frustum_and_local_map_search_equal_the_reference_functions(gold, oracle_mod, i):
    """Frame::isInFrustum (src/Frame.cc:587-643) with Frame::UpdatePoseMatrices (:578-584), MapPoint::PredictScale /
    Get{Min,Max}DistanceInvariance (src/MapPoint.cc:376-386, 405-420), then SearchByProjection(F, vpMapPoints, th): in-view flags,
    mTrackProj* / mTrackViewCos bit patterns, predicted levels, final matches."""
    g = _gen()
    pm = g.local_problem(g.LOCAL_CASES[i], gold[f"local{i}_ow"])
    assert g.problem_crc(pm) == int(gold[f"local{i}_crc"]), "synthetic generator drifted: regenerate the fixture"
    n, km, _, _, ex = oracle_mod.search_by_projection(pm)
    in_view = (ex["q_flags"] & 1).astype(np.uint8)
    assert (in_view == gold[f"local{i}_in_view"]).all() and in_view.sum() > 1000
    assert (ex["q_track"].view(np.uint32) == gold[f"local{i}_track"].view(np.uint32)).all()
    assert (np.where(in_view > 0, ex["q_level"], -1) == gold[f"local{i}_level"]).all()
    assert n == int(gold[f"local{i}_n"]) and (np.where(km >= 0, km, -1) == gold[f"local{i}_kp_match"]).all()


@pytest.mark.parametrize("i", [0, 1, 2])
def test_oracle_fuse_search_equals_the_reference_function(gold, oracle_mod, i):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:825-975) with KeyFrame::GetFeaturesInArea / IsInImage
    (src/KeyFrame.cc:589-633) and MapPoint::PredictScale(dist, KeyFrame*) (src/MapPoint.cc:388-403) of the reference itself; its map
    surgery (Replace / AddObservation / AddMapPoint) is recorded by the stand-ins: which key-point every map point was fused with."""
    g = _gen()
    pf = g.fuse_problem(g.FUSE_CASES[i], gold[f"fuse{i}_ow"])
    assert g.problem_crc(pf) == int(gold[f"fuse{i}_crc"]), "synthetic generator drifted: regenerate the fixture"
    n, km, bi, bd, _ = oracle_mod.search_by_projection(pf)
    assert n == int(gold[f"fuse{i}_n"]) and n > 500
    assert (np.where(bd <= 50, bi, -1) == gold[f"fuse{i}_fused_with"]).all()


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference")), reason="reference tree / oracle/_ref not present (GPU box)")
def test_fixture_is_what_the_reference_library_computes_now(gold, oracle_mod):
    import ctypes as C
    g = _gen()
    L = C.CDLL(REF_LIB)
    n, km = g.ref_last(L, _problem(g, gold, "last", 1))
    assert n == int(gold["last1_n"]) and (km == gold["last1_kp_match"]).all()
    n, km = g.ref_map(L, _problem(g, gold, "map", 2))
    assert n == int(gold["map2_n"]) and (km == gold["map2_kp_match"]).all()
    n, km, inview, track, level, ow = g.ref_local(L, g.local_problem(g.LOCAL_CASES[0]))
    assert n == int(gold["local0_n"]) and (km == gold["local0_kp_match"]).all() and (ow == gold["local0_ow"]).all()


BOW_GOLD = os.path.join(ROOT, "tests", "golden", "bow_ref.npz")


def _gen_bow():
    spec = importlib.util.spec_from_file_location("gen_ref_bow_golden", os.path.join(ROOT, "oracle", "gen_ref_bow_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    return g


@pytest.mark.parametrize("i", range(6))
def test_oracle_bow_searches_equal_the_reference_functions(oracle_mod, i):
    """tests/golden/bow_ref.npz: ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:159-288; cases 0-2) and
    SearchForTriangulation + CheckDistEpipolarLine (:657-823, 140-157; cases 3-5) of the reference itself, walking FeatureVectors
    with one-sided nodes between the common ones (oracle/gen_ref_bow_golden.py)."""
    g = _gen_bow()
    gold = np.load(BOW_GOLD)
    pr = g.problem(g.CASES[i])
    assert g.problem_crc(pr) == int(gold[f"c{i}_crc"]), "synthetic generator drifted: regenerate the fixture"
    n, match = oracle_mod.search_by_bow(pr)
    assert n == int(gold[f"c{i}_n"]) and n > 200
    assert (match == gold[f"c{i}_match"]).all()


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference")), reason="reference tree / oracle/_ref not present (GPU box)")
def test_bow_fixture_is_what_the_reference_library_computes_now():
    import ctypes as C
    g = _gen_bow()
    gold = np.load(BOW_GOLD)
    L = C.CDLL(REF_LIB)
    for i in (1, 4):
        n, m = g.run_ref(L, g.problem(g.CASES[i]))
        assert n == int(gold[f"c{i}_n"]) and (m == gold[f"c{i}_match"]).all()
