"""Pins the stereo matcher of oracle/match_oracle.cpp to the LITERAL reference: tests/golden/stereo_ref.npz holds mvuRight / mvDepth
of the reference's own Frame::ComputeStereoMatches (src/Frame.cc:829-1003) and distances of its ORBmatcher::DescriptorDistance
(src/ORBmatcher.cc:1647-1663) -- the two function bodies compiled from /root/reference by `make -C oracle ref` between stand-in
class declarations (oracle/ref_match.cpp, oracle/ref_shim/cv_shim.h) and run by oracle/gen_ref_match_golden.py on seeded stereo
pairs.  Bar: bit patterns equal.  Where the reference tree and the built library exist (the build container) the fixture is also
re-derived live."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "stereo_ref.npz")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_match.so")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _case_inputs(gold, ci, oracle_mod):
    from airdos_b200 import synth
    seed, w, h, nf, ini, mn = [int(v) for v in gold[f"c{ci}_params"]]
    if f"c{ci}_left" in gold:
        il, ir = gold[f"c{ci}_left"], gold[f"c{ci}_right"]
    else:
        il, ir = synth.make_stereo_pair(seed, w, h)
    assert [zlib.crc32(il.tobytes()), zlib.crc32(ir.tobytes())] == [int(v) for v in gold[f"c{ci}_crc"]], "synthetic generator drifted: regenerate the fixture"
    a = oracle_mod.orb_extract(il, None, nf, 1.2, 8, ini, mn, want_pyramid=True)
    b = oracle_mod.orb_extract(ir, None, nf, 1.2, 8, ini, mn, want_pyramid=True)
    sc = oracle_mod.orb_params(nf, 1.2, 8, w, h)["scale"]
    return a, b, sc, synth.BF / synth.FX, synth.BF


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_oracle_stereo_equals_the_reference_function(gold, oracle_mod, ci):
    a, b, sc, mb, mbf = _case_inputs(gold, ci, oracle_mod)
    ur, dp, _, _ = oracle_mod.stereo_match(a["kps"], a["desc"], b["kps"], b["desc"], a["pyramid"], b["pyramid"], sc, mb, mbf)
    assert len(ur) == len(gold[f"c{ci}_u_right"])
    assert (ur.view(np.uint32) == gold[f"c{ci}_u_right"].view(np.uint32)).all()
    assert (dp.view(np.uint32) == gold[f"c{ci}_depth"].view(np.uint32)).all()
    assert (dp > 0).sum() > 200


def test_oracle_hamming_equals_the_reference_function(gold, oracle_mod):
    d = gold["dd_pairs"]
    got = np.array([oracle_mod.hamming(x[0], x[1]) for x in d], np.int32)
    assert (got == gold["dd_dist"]).all()
    assert (gold["dd_dist"][:8] == 0).all() and (gold["dd_dist"][8:16] == 256).all()


def _gen():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_ref_match_golden", os.path.join(ROOT, "oracle", "gen_ref_match_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    return g


def test_oracle_distinctive_descriptor_equals_the_reference_function(gold, oracle_mod):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:245-310) of the reference itself on 300 observation sets: the
    descriptor it leaves in mDescriptor is the one the oracle's index points at."""
    desc, ptr = _gen().distinct_case(np.random.default_rng(20261018))
    assert [zlib.crc32(desc.tobytes()), zlib.crc32(ptr.tobytes())] == [int(v) for v in gold["distinct_crc"]]
    best = oracle_mod.distinctive(desc, ptr)
    for p in range(len(ptr) - 1):
        if ptr[p + 1] == ptr[p]:
            assert best[p] == -1 and (gold["distinct_chosen"][p] == 0xAB).all()
        else:
            assert (desc[ptr[p] + best[p]] == gold["distinct_chosen"][p]).all(), p


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference")), reason="reference tree / oracle/_ref not present (GPU box)")
def test_fixture_is_what_the_reference_library_computes_now(gold, oracle_mod):
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_ref_match_golden", os.path.join(ROOT, "oracle", "gen_ref_match_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    L = C.CDLL(REF_LIB)
    a, b, sc, mb, mbf = _case_inputs(gold, 2, oracle_mod)
    ur, dp = g.ref_stereo(L, a["kps"], a["desc"], b["kps"], b["desc"], a["pyramid"], b["pyramid"], sc, mb, mbf)
    assert (ur.view(np.uint32) == gold["c2_u_right"].view(np.uint32)).all() and (dp.view(np.uint32) == gold["c2_depth"].view(np.uint32)).all()
