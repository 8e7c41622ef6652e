"""GPU parity tests of the CUDA extractor (through the C-ABI) against the CPU oracle and the
committed cv2-derived golden fixtures.  Bar: bit-exact key-points (incl. float angle patterns),
descriptors, pyramids and FAST candidate lists."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def adb():
    import airdos_b200
    return airdos_b200


def _same(kps, desc, ref):
    return len(kps) == len(ref["kps"]) and kps.tobytes() == ref["kps"].tobytes() and bool((desc == ref["desc"]).all())


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "orb_*.npz"))), ids=os.path.basename)
def test_cuda_matches_golden_fixture(adb, path):
    g = np.load(path, allow_pickle=False)
    nf, nl, ini, mn = [int(v) for v in g["params"]]
    img = g["image"]
    mask = g["mask"] if g["mask"].size else None
    ex = adb.ORBextractor(nf, float(g["scale"]), nl, ini, mn, img.shape[1], img.shape[0])
    kps, desc = ex(img, mask)
    assert len(kps) == len(g["kps"])
    assert kps.tobytes() == g["kps"].tobytes()
    assert (desc == g["desc"]).all()
    pyr = ex.pyramid(0)
    assert [p.shape for p in pyr] == [tuple(s) for s in g["pyramid_sizes"]]
    assert [int(p.astype(np.int64).sum()) for p in pyr] == [int(v) for v in g["pyramid_sums"]]
    ex.close()


def test_cuda_matches_the_reference_operator():
    """adb_orb_extract against tests/golden/extractor_ref.npz = the reference's own ORBextractor::operator() (with ComputePyramid,
    ComputeKeyPointsOctTree, DistributeOctTree, IC_Angle, computeOrbDescriptor: whole definitions compiled from /root/reference,
    oracle/ref_orb.cpp; OpenCV calls = the cv2-pinned primitives) on seven seeded images: key-points (all six fields, angle bit patterns
    included), descriptors and the pyramid bit for bit."""
    import importlib.util, zlib
    import airdos_b200 as adb
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ref_orb_golden", os.path.join(root, "oracle", "gen_ref_orb_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    gold = np.load(os.path.join(GOLDEN, "extractor_ref.npz"))
    for i in range(len(g.EXTRACT_CASES)):
        img, msk, nf, ini, mn = g.make_extract_case(i)
        ex = adb.ORBextractor(nf, 1.2, 8, ini, mn, img.shape[1], img.shape[0])
        kps, desc = ex(img, msk)
        assert len(kps) == int(gold[f"e{i}_n"]), i
        assert zlib.crc32(kps.tobytes()) == int(gold[f"e{i}_kps_crc"]) and zlib.crc32(np.ascontiguousarray(desc).tobytes()) == int(gold[f"e{i}_desc_crc"]), i
        assert zlib.crc32(np.concatenate([p.ravel() for p in ex.pyramid(0)]).tobytes()) == int(gold[f"e{i}_pyr_crc"]), i
        if f"e{i}_kps" in gold.files:
            assert kps.tobytes() == gold[f"e{i}_kps"].tobytes() and (desc == gold[f"e{i}_desc"]).all()
        ex.close()


@pytest.mark.parametrize("w,h,nf,ini,mn,masked", [(640, 480, 1000, 12, 7, False), (640, 480, 2000, 12, 7, False),
                                                  (640, 480, 1000, 12, 7, True), (640, 360, 1500, 12, 7, False),
                                                  (320, 240, 500, 20, 7, True), (333, 251, 700, 20, 7, False),
                                                  (752, 480, 1200, 20, 7, False), (1241, 376, 2000, 20, 7, False)])
def test_cuda_matches_oracle_batch(adb, oracle_mod, w, h, nf, ini, mn, masked):
    """Seeded frames, batched (ragged sizes, odd pitches, two quad-tree roots for wide images)."""
    from airdos_b200 import synth
    F = 3
    imgs = np.stack([synth.make_stereo_pair(10 + f, w, h)[f % 2] for f in range(F)])
    masks = np.stack([synth.make_mask(20 + f, w, h, 3) for f in range(F)]) if masked else None
    ex = adb.ORBextractor(nf, 1.2, 8, ini, mn, w, h, max_batch=F)
    kps, desc, cnt = ex.extract_batch(imgs, masks)
    for f in range(F):
        o = oracle_mod.orb_extract(imgs[f], None if masks is None else masks[f], nf, 1.2, 8, ini, mn, want_pyramid=True)
        for a, b in zip(ex.pyramid(f), o["pyramid"]):
            assert (a == b).all()
        assert [len(ex.debug_candidates(f, l)) for l in range(8)] == list(o["cand_counts"])
        assert _same(kps[f, :cnt[f]], desc[f, :cnt[f]], o), f"frame {f}"
    # single-frame call on the same handle gives the same answer as the batch (idempotence)
    k1, d1 = ex(imgs[1], None if masks is None else masks[1])
    assert k1.tobytes() == kps[1, :cnt[1]].tobytes() and (d1 == desc[1, :cnt[1]]).all()
    ex.close()


@pytest.mark.parametrize("ini,mn", [(90, 7), (40, 7), (7, 7), (5, 9), (250, 200)])
def test_ini_min_rule_both_ends_of_the_slot(adb, oracle_mod, ini, mn):
    """The warp-per-cell FAST kernel writes iniTh corners from the front of a cell's slot and minTh-only corners from its back; the
    cell count says which end holds the answer (src/ORBextractor.cc:812-824).  High iniTh: most cells fall back to minTh; iniTh ==
    minTh and iniTh < minTh: one list only; thresholds nothing passes: empty cells."""
    from airdos_b200 import synth
    w, h = 640, 480
    img = synth.make_stereo_pair(31, w, h)[0]
    img[:, : w // 3] = (img[:, : w // 3] // 8 + 100).astype(np.uint8)          # a low-contrast third: no corner at a high iniTh
    msk = synth.make_mask(32, w, h, 2)
    ex = adb.ORBextractor(1500, 1.2, 8, ini, mn, w, h)
    for m in (None, msk):
        k, d = ex(img, m)
        o = oracle_mod.orb_extract(img, m, 1500, 1.2, 8, ini, mn)
        assert [len(ex.debug_candidates(0, l)) for l in range(8)] == list(o["cand_counts"])
        assert _same(k, d, o), (ini, mn, m is not None)
    ex.close()


def test_erosion_constant_tile_path_and_general_values(adb, oracle_mod):
    """cv::erode(mask, ones(10, 10)) (src/ORBextractor.cc:1130-1131): the level-0 mask the extractor keeps equals the oracle's erosion
    byte for byte -- for masks whose 128 x 32 tiles are all constant (all 255, all 0: the kernel's load-and-store-only path), for
    rectangles (constant and mixed tiles side by side, borders where the outside counts as 255), for single zero pixels next to
    tile seams and image corners, and for a grey-valued noise mask (no constant tile, a true minimum rather than an AND)."""
    from airdos_b200 import synth
    w, h = 640, 480
    img = synth.make_stereo_pair(35, w, h)[0]
    rng = np.random.default_rng(36)
    dots = np.full((h, w), 255, np.uint8)
    for y, x in ((0, 0), (h - 1, w - 1), (31, 127), (32, 128), (36, 133), (250, 383), (251, 384), (479, 0), (5, 639)):
        dots[y, x] = 0
    big = np.full((h, w), 255, np.uint8); big[40:420, 100:560] = 0                   # holds whole all-zero tiles, halo included
    masks = [np.full((h, w), 255, np.uint8), np.zeros((h, w), np.uint8), synth.make_mask(37, w, h, 3), big, dots,
             rng.integers(0, 256, (h, w), dtype=np.uint8), synth.make_human_mask(38, w, h)]
    ex = adb.ORBextractor(800, 1.2, 8, 20, 7, w, h)
    for i, m in enumerate(masks):
        ex(img, m)
        got = ex.pyramid(0, which=1)[0]
        assert got.shape == (h, w) and (got == oracle_mod.erode10(m)).all(), i
    ex.close()


def test_cta_per_cell_fallback_kernel_agrees(adb, oracle_mod, monkeypatch):
    """ADB_FAST_CTA=1 selects the CTA-per-cell FAST kernel (the path for cells wider than the warp kernel's three column tiles)."""
    from airdos_b200 import synth
    monkeypatch.setenv("ADB_FAST_CTA", "1")
    w, h = 640, 480
    img = synth.make_stereo_pair(33, w, h)[1]
    msk = synth.make_mask(34, w, h, 3)
    ex = adb.ORBextractor(1200, 1.2, 8, 20, 7, w, h)
    for m in (None, msk):
        k, d = ex(img, m)
        assert _same(k, d, oracle_mod.orb_extract(img, m, 1200, 1.2, 8, 20, 7))
    ex.close()


def test_portrait_shapes_with_zero_roots_are_refused(adb):
    """nIni = round(w / h) = 0 (src/ORBextractor.cc:545-549) is undefined in the reference: adb_orb_create refuses the shape."""
    with pytest.raises(adb.AdbError) as e:
        adb.ORBextractor(500, 1.2, 8, 20, 7, 200, 640)
    assert e.value.status == 1
    adb.ORBextractor(500, 1.2, 8, 20, 7, 300, 400).close()


def test_five_argument_constructor_provisions_lazily(adb, oracle_mod):
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) as the reference declares it (include/ORBextractor.h:51-52):
    getters work before the first frame, the first operator() provisions for its image size, a different size re-provisions."""
    from airdos_b200 import synth
    ex = adb.ORBextractor(1000, 1.2, 8, 12, 7)
    assert ex.GetLevels() == 8 and abs(ex.GetScaleFactors()[3] - np.float32(1.2) ** 3) < 1e-6 and sum(ex.quotas()) == 1000
    for w, h in ((640, 480), (320, 240), (640, 480)):
        img = synth.make_stereo_pair(70, w, h)[0]
        msk = synth.make_mask(71, w, h, 2)
        k, d = ex(img, msk)
        o = oracle_mod.orb_extract(img, msk, 1000, 1.2, 8, 12, 7)
        assert _same(k, d, o), (w, h)
        assert ex.level_sizes()[0] == (w, h)
        sized = adb.ORBextractor(1000, 1.2, 8, 12, 7, w, h)
        k2, d2 = sized(img, msk)
        assert k2.tobytes() == k.tobytes() and (d2 == d).all()
        sized.close()
    ex.close()


def test_getters_equal_the_reference_constructor_tables(adb):
    """GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares and the per-level quotas of the
    handle against tests/golden/extractor_tables_ref.npz = the members the reference's own constructor fills (src/ORBextractor.cc:411-472,
    compiled from /root/reference by oracle/ref_orb.cpp), as bit patterns."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ref_orb_golden", os.path.join(root, "oracle", "gen_ref_orb_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    gold = np.load(os.path.join(root, "tests", "golden", "extractor_tables_ref.npz"))
    for c, (nf, sf, nl) in enumerate(g.TABLE_CONFIGS):
        ex = adb.ORBextractor(nf, sf, nl, 20, 7)
        assert ex.GetLevels() == nl
        for name, got in (("scale", ex.GetScaleFactors()), ("inv_scale", ex.GetInverseScaleFactors()), ("sigma2", ex.GetScaleSigmaSquares()),
                          ("inv_sigma2", ex.GetInverseScaleSigmaSquares())):
            assert np.asarray(got, np.float32).tobytes() == gold[f"t{c}_{name}"].tobytes(), (nf, sf, nl, name)
        assert list(ex.quotas()) == list(gold[f"t{c}_quota"]), (nf, sf, nl)
        ex.close()


def test_chunked_masked_host_batch_equals_single_calls(adb, oracle_mod):
    """A host batch of >= 32 frames runs as a chunk pipeline (upload / kernels / download overlapped), with the masks
    (the reference passes one with every frame: src/Frame.cc:551-571) uploaded, eroded and resized chunk by chunk.  The result
    must equal one call per frame, and the oracle on a sample."""
    from airdos_b200 import synth
    F, w, h = 40, 640, 480
    base = [synth.make_stereo_pair(50 + f, w, h)[f % 2] for f in range(5)]
    bmask = [synth.make_mask(60 + f, w, h, 3) for f in range(5)]
    imgs = np.stack([base[f % 5] for f in range(F)]); masks = np.stack([bmask[(f * 3) % 5] for f in range(F)])
    ex = adb.ORBextractor(2000, 1.2, 8, 12, 7, w, h, max_batch=F)
    kps, desc, cnt = ex.extract_batch(imgs, masks)
    one = adb.ORBextractor(2000, 1.2, 8, 12, 7, w, h, max_batch=1)
    for f in (0, 9, 10, 19, 20, 31, 39):       # both sides of every chunk boundary (40 frames / 8 chunks = 5 per chunk)
        k1, d1 = one(imgs[f], masks[f])
        assert k1.tobytes() == kps[f, :cnt[f]].tobytes() and (d1 == desc[f, :cnt[f]]).all(), f
    for f in (7, 23):
        o = oracle_mod.orb_extract(imgs[f], masks[f], 2000, 1.2, 8, 12, 7)
        assert _same(kps[f, :cnt[f]], desc[f, :cnt[f]], o), f
    # and the unmasked chunk pipeline still agrees with the masked one where the mask keeps everything
    full = np.full_like(masks, 255)
    ka, da, ca = ex.extract_batch(imgs, full)
    ka, da, ca = ka.copy(), da.copy(), ca.copy()
    kb, db, cb = ex.extract_batch(imgs, None)
    assert (ca == cb).all() and all(ka[f, :ca[f]].tobytes() == kb[f, :cb[f]].tobytes() for f in range(F))
    ex.close(); one.close()


def test_device_resident_path_and_unaligned_input(adb, oracle_mod):
    import torch
    from airdos_b200 import synth
    F, w, h = 4, 640, 480
    imgs = np.stack([synth.make_stereo_pair(30 + f)[0] for f in range(F)])
    ex = adb.ORBextractor(2000, 1.2, 8, 12, 7, w, h, max_batch=F)
    ref_k, ref_d, ref_c = ex.extract_batch(imgs)
    # aligned device buffer: level 0 is read in place
    t = torch.from_numpy(imgs).cuda()
    ex.extract_batch_device(t.data_ptr(), F)
    ex.sync()
    k, d, c = ex.download(0, F)
    assert (c == ref_c).all() and k.tobytes() == ref_k.tobytes() and (d == ref_d).all()
    # unaligned base pointer (+1 byte): staged through the internal copy
    buf = torch.zeros(F * w * h + 64, dtype=torch.uint8, device="cuda")
    buf[1:1 + F * w * h] = t.flatten()
    ex.extract_batch_device(buf.data_ptr() + 1, F)
    ex.sync()
    k, d, c = ex.download(0, F)
    assert (c == ref_c).all() and k.tobytes() == ref_k.tobytes() and (d == ref_d).all()
    o = oracle_mod.orb_extract(imgs[2], None, 2000, 1.2, 8, 12, 7)
    assert _same(k[2, :c[2]], d[2, :c[2]], o)
    ex.close()


def test_edge_cases(adb, oracle_mod):
    ex = adb.ORBextractor(1000, 1.2, 8, 12, 7, 640, 480)
    # flat image: no corners anywhere, count 0 (and the ini -> min fallback runs in every cell)
    k, d = ex(np.full((480, 640), 128, np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    # empty image: silent return like src/ORBextractor.cc:1057-1058
    k, d = ex(np.zeros((0, 0), np.uint8))
    assert len(k) == 0
    # all-zero mask rejects everything
    from airdos_b200 import synth
    img = synth.make_stereo_pair(3)[0]
    k, d = ex(img, np.zeros_like(img))
    assert len(k) == 0
    # saturated noise: maximum candidate density (exercises slot capacity and a deep quad-tree)
    noise = np.random.default_rng(0).integers(0, 2, (480, 640)).astype(np.uint8) * 255
    k, d = ex(noise)
    o = oracle_mod.orb_extract(noise, None, 1000, 1.2, 8, 12, 7)
    assert _same(k, d, o)
    # wrong size is refused, not silently resized
    with pytest.raises(adb.AdbError):
        ex(np.zeros((100, 100), np.uint8))
    ex.close()
    # tiny image: upper levels have no FAST cells at all
    small = synth.make_stereo_pair(5, 96, 80)[0]
    ex = adb.ORBextractor(200, 1.2, 8, 20, 7, 96, 80)
    k, d = ex(small)
    o = oracle_mod.orb_extract(small, None, 200, 1.2, 8, 20, 7)
    assert _same(k, d, o)
    ex.close()


def test_full_size_batch_properties(adb, oracle_mod):
    """BASELINE config 2 shape at a full batch: size-independent properties + spot checks vs the oracle."""
    from airdos_b200 import synth
    F = 64
    base = synth.make_stereo_batch(4).reshape(8, 480, 640)
    imgs = np.concatenate([base] * (F // 8))
    ex = adb.ORBextractor(2000, 1.2, 8, 12, 7, 640, 480, max_batch=F)
    kps, desc, cnt = ex.extract_batch(imgs)
    q = np.array(ex.quotas())
    for f in range(F):
        # identical frames give identical results wherever they sit in the batch
        assert cnt[f] == cnt[f % 8]
        assert kps[f, :cnt[f]].tobytes() == kps[f % 8, :cnt[f]].tobytes()
        assert (desc[f, :cnt[f]] == desc[f % 8, :cnt[f]]).all()
        k = kps[f, :cnt[f]]
        assert (np.diff(k["octave"]) >= 0).all()                        # levels concatenated in order
        per = np.bincount(k["octave"], minlength=8)
        assert (per <= q + 2).all() and per.sum() >= 1900
        assert (k["angle"] >= 0).all() and (k["angle"] < 360).all()
    for f in (0, 5):
        assert _same(kps[f, :cnt[f]], desc[f, :cnt[f]], oracle_mod.orb_extract(imgs[f], None, 2000, 1.2, 8, 12, 7))
    ex.close()


@pytest.mark.parametrize("nfeat", [1000, 1003])
def test_fused_gather_targets(adb, oracle_mod, nfeat):
    """adb_orb_set_gather: the descriptor kernel also stores every record into the given target buffers (on one GPU
    the 'peers' are two local buffers; over NVLink they are peer-mapped symmetric memory, exercised by bench.py --gpus 2).
    1000 features: capacity 1024, records leave as staged 16-byte stores; 1003: odd capacity, the word-granular path."""
    import torch
    from airdos_b200 import synth
    F = 3
    imgs = np.stack([synth.make_stereo_pair(50 + f)[0] for f in range(F)])
    ex = adb.ORBextractor(nfeat, 1.2, 8, 12, 7, 640, 480, max_batch=F)
    cap = ex.capacity
    assert (cap % 8 == 0) == (nfeat == 1000)
    tk = [torch.zeros(F, cap, 24, dtype=torch.uint8, device="cuda") for _ in range(2)]
    td = [torch.zeros(F, cap, 32, dtype=torch.uint8, device="cuda") for _ in range(2)]
    tc = [torch.zeros(F, dtype=torch.int32, device="cuda") for _ in range(2)]
    ex.set_gather([t.data_ptr() for t in tk], [t.data_ptr() for t in td], [t.data_ptr() for t in tc])
    kps, desc, cnt = ex.extract_batch(imgs)
    for g in range(2):
        assert (tc[g].cpu().numpy() == cnt).all()
        for f in range(F):
            n = cnt[f]
            assert tk[g][f, :n].cpu().numpy().tobytes() == kps[f, :n].tobytes()
            assert (td[g][f, :n].cpu().numpy() == desc[f, :n]).all()
    ex.set_gather()                      # off again
    tk[0].zero_()
    ex.extract_batch(imgs)
    assert int(tk[0].sum()) == 0
    ex.close()
