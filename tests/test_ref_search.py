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


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference")), reason="reference tree / oracle/_ref not present (GPU box)")
def test_fixture_is_what_the_reference_library_computes_now(gold, oracle_mod):
    import ctypes as C
    g = _gen()
    L = C.CDLL(REF_LIB)
    n, km = g.ref_last(L, _problem(g, gold, "last", 1))
    assert n == int(gold["last1_n"]) and (km == gold["last1_kp_match"]).all()
    n, km = g.ref_map(L, _problem(g, gold, "map", 2))
    assert n == int(gold["map2_n"]) and (km == gold["map2_kp_match"]).all()
