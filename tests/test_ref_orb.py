"""Pins the quad-tree of oracle/orb_oracle.cpp to the LITERAL reference: tests/golden/quadtree_ref.npz holds the results of the
reference's own ORBextractor::DistributeOctTree / ExtractorNode::DivideNode (src/ORBextractor.cc:497-765, compiled from /root/reference:
oracle/ref_orb.cpp) on 120 seeded candidate sets -- 1 to 7000 distinct pixels, one or two roots, quotas from 1 to above the count.
The reference breaks ties between nodes of equal size by heap ADDRESS; the fixture is taken with an allocator that hands out
increasing addresses (address order = creation order = DESIGN.md convention D.1).  With the stock malloc the reference itself
reproduces only a fraction of these results (recorded in the fixture): its output order is a property of the allocator there."""
import importlib.util
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "quadtree_ref.npz")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_orb.so")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_ref_orb_golden", os.path.join(ROOT, "oracle", "gen_ref_orb_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    return g


def test_oracle_quadtree_equals_the_reference_function(oracle_mod):
    g = _gen()
    gold = np.load(GOLD)
    for i in range(g.N_CASES):
        cand, w, h, n = g.make_case(i)
        o = oracle_mod.distribute(cand, 0, w, 0, h, n)
        assert len(o) == int(gold["count"][i]), i
        assert zlib.crc32(o.tobytes()) == int(gold["crc"][i]), i
        if i < 4:
            assert (o == gold[f"out{i}"]).all()
    assert int(gold["same_with_stock_malloc"]) < g.N_CASES      # the address tie-break is real: the reference disagrees with itself


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference")), reason="reference tree / oracle/_ref not present (GPU box)")
def test_fixture_is_what_the_reference_library_computes_now():
    import ctypes as C
    g = _gen()
    gold = np.load(GOLD)
    L = C.CDLL(REF_LIB)
    for i in (0, 7, 33, 119):
        cand, w, h, n = g.make_case(i)
        a = g.ref_distribute(L, cand, 0, w, 0, h, n, 1)
        assert len(a) == int(gold["count"][i]) and zlib.crc32(a.tobytes()) == int(gold["crc"][i])


TABLES = os.path.join(ROOT, "tests", "golden", "extractor_tables_ref.npz")


def test_oracle_constructor_tables_equal_the_reference_constructor(oracle_mod):
    """tests/golden/extractor_tables_ref.npz holds what ORBextractor::ORBextractor (src/ORBextractor.cc:411-472, compiled from /root/reference)
    leaves in mvScaleFactor / mvInvScaleFactor / mvLevelSigma2 / mvInvLevelSigma2 / mnFeaturesPerLevel / umax / pattern (the last copied
    from bit_pattern_31_, :151-409) for seven configurations; the oracle's tables are the same bit patterns."""
    g = _gen()
    gold = np.load(TABLES)
    for c, (nf, sf, nl) in enumerate(g.TABLE_CONFIGS):
        o = oracle_mod.orb_tables(nf, sf, nl)
        for k, v in o.items():
            if k == "pattern" and c > 0:
                continue
            assert v.tobytes() == gold[f"t{c}_{k}"].tobytes(), (nf, sf, nl, k)


def test_oracle_orientation_and_descriptor_equal_the_reference_functions(oracle_mod):
    """IC_Angle and computeOrbDescriptor (src/ORBextractor.cc:78-148, compiled from /root/reference with cv::fastAtan2 / cvRound restated
    in oracle/ref_shim/cv_shim.h) on noise, smooth and block images: angles equal as bit patterns, descriptors byte for byte."""
    g = _gen()
    gold = np.load(TABLES)
    for i in range(g.N_DESC_CASES):
        img, xy = g.make_desc_case(i)
        a, d = oracle_mod.orient_describe(img, oracle_mod.blur7(img), xy)
        assert a.tobytes() == gold[f"d{i}_angle"].tobytes(), i
        assert (d == gold[f"d{i}_desc"]).all(), i


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference")), reason="reference tree / oracle/_ref not present (GPU box)")
def test_tables_fixture_is_what_the_reference_library_computes_now(oracle_mod):
    import ctypes as C
    g = _gen()
    gold = np.load(TABLES)
    L = C.CDLL(REF_LIB)
    r = oracle_mod.orb_tables(*g.TABLE_CONFIGS[4], lib=L)
    assert all(r[k].tobytes() == gold[f"t4_{k}"].tobytes() for k in r if k != "pattern")
    img, xy = g.make_desc_case(2)
    a, d = oracle_mod.orient_describe(img, oracle_mod.blur7(img), xy, lib=L)
    assert a.tobytes() == gold["d2_angle"].tobytes() and (d == gold["d2_desc"]).all()


EXTRACT = os.path.join(ROOT, "tests", "golden", "extractor_ref.npz")


def test_oracle_extractor_equals_the_reference_operator(oracle_mod):
    """tests/golden/extractor_ref.npz: the reference's own ORBextractor::operator(), ComputePyramid, ComputeKeyPointsOctTree, computeOrientation
    and computeDescriptors (src/ORBextractor.cc:474-481, 767-864, 1040-1156: whole definitions compiled from /root/reference, oracle/ref_orb.cpp)
    on seven seeded images (VGA, QVGA, 752 x 480, 640 x 360, a 200 x 150 image whose top levels are empty, a 1241 x 376 image with three
    quad-tree roots; with and without a person mask; three threshold pairs).  Their five OpenCV calls land in the oracle's primitives,
    everything else -- the in-place ROI / border handling of the pyramid, the resized mask pyramid, the cell grid and the ini / min
    threshold rule, the quad-tree, the orientation, the per-level blur and descriptors, the final scaling and the output order -- is the
    reference's.  The oracle's extract() gives the same key-points, descriptors and pyramid bit for bit."""
    g = _gen()
    gold = np.load(EXTRACT)
    for i in range(len(g.EXTRACT_CASES)):
        img, msk, nf, ini, mn = g.make_extract_case(i)
        o = oracle_mod.orb_extract(img, msk, nf, 1.2, 8, ini, mn, want_pyramid=True)
        assert len(o["kps"]) == int(gold[f"e{i}_n"]), i
        assert zlib.crc32(o["kps"].tobytes()) == int(gold[f"e{i}_kps_crc"]) and zlib.crc32(o["desc"].tobytes()) == int(gold[f"e{i}_desc_crc"]), i
        assert zlib.crc32(np.concatenate([l.ravel() for l in o["pyramid"]]).tobytes()) == int(gold[f"e{i}_pyr_crc"]), i
        assert (np.bincount(o["kps"]["octave"], minlength=8) == gold[f"e{i}_per_level"]).all()
        if f"e{i}_kps" in gold.files:
            assert o["kps"].tobytes() == gold[f"e{i}_kps"].tobytes() and (o["desc"] == gold[f"e{i}_desc"]).all()


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.isdir("/root/reference")), reason="reference tree / oracle/_ref not present (GPU box)")
def test_extractor_fixture_is_what_the_reference_library_computes_now(oracle_mod):
    import ctypes as C
    g = _gen()
    gold = np.load(EXTRACT)
    L = C.CDLL(REF_LIB)
    for i in (2, 4):
        img, msk, nf, ini, mn = g.make_extract_case(i)
        k, d, pyr = oracle_mod.ref_orb_extract(L, img, msk, nf, 1.2, 8, ini, mn)
        assert k.tobytes() == gold[f"e{i}_kps"].tobytes() and (d == gold[f"e{i}_desc"]).all() and zlib.crc32(pyr.tobytes()) == int(gold[f"e{i}_pyr_crc"])
