"""CPU tests of the matcher oracle against numpy bit counting and brute force."""
import numpy as np


def _popcount_dist(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def test_hamming_matches_bitcount(oracle_mod):
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    for i in range(0, 64, 2):
        assert oracle_mod.hamming(d[i], d[i + 1]) == _popcount_dist(d[i], d[i + 1])
    assert oracle_mod.hamming(d[0], d[0]) == 0
    assert oracle_mod.hamming(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_best2_rule(oracle_mod):
    rng = np.random.default_rng(1)
    q = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (97, 32), dtype=np.uint8)
    t[50] = t[10]; t[51] = t[10]                       # exact duplicates -> ties
    q[3] = t[10]
    dist = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(2)
    bi, bd, sd = oracle_mod.best2(q, t)
    assert (bi == dist.argmin(1)).all()                 # first minimum wins (strict '<')
    assert (bd == dist.min(1)).all()
    assert (sd == np.sort(dist, 1)[:, 1]).all()         # second order statistic, duplicates counted
    assert bi[3] == 10 and bd[3] == 0 and sd[3] == 0
    # candidate lists: empty list -> (-1, 256, 256); order inside the list decides ties
    q3 = q[[0, 3, 1]]
    off = np.array([0, 0, 3, 5], np.int32)
    idx = np.array([51, 10, 50, 7, 7], np.int32)
    bi, bd, sd = oracle_mod.best2(q3, t, off, idx)
    assert (bi[0], bd[0], sd[0]) == (-1, 256, 256)
    assert bi[1] == 51 and bd[1] == 0 and sd[1] == 0    # 51 listed first among the duplicates
    assert bi[2] == 7 and bd[2] == sd[2]


def test_stereo_oracle_on_synthetic_pair(oracle_mod):
    from airdos_b200 import synth
    L, R = synth.make_stereo_pair(0)
    a = oracle_mod.orb_extract(L, None, 1000, 1.2, 8, 12, 7, want_pyramid=True)
    b = oracle_mod.orb_extract(R, None, 1000, 1.2, 8, 12, 7, want_pyramid=True)
    sc = oracle_mod.orb_params(1000, 1.2, 8, 640, 480)["scale"]
    mbf = synth.BF; mb = mbf / synth.FX
    ur, dp, hi, hd = oracle_mod.stereo_match(a["kps"], a["desc"], b["kps"], b["desc"], a["pyramid"], b["pyramid"], sc, mb, mbf)
    m = dp > 0
    assert m.sum() > 200                                 # the synthetic pair is a real stereo pair
    assert (ur[~m] == -1).all() and (dp[~m] == -1).all()
    disp = a["kps"]["x"][m] - ur[m]
    assert (disp > 0).all() and (disp < mbf / mb).all()
    assert np.allclose(dp[m], mbf / disp, rtol=1e-6)
    # matched rows agree with the Hamming stage and the distance threshold (TH_HIGH + TH_LOW) / 2
    assert (hd[m] < 75).all() and (hi[m] >= 0).all()
    # the true disparity of the generator is bf / Z with Z in [3, 40]: matches must sit in that range mostly
    assert np.median(disp) > synth.BF / 40 * 0.8


def _distinct_case(rng, n_points=200):
    lens = rng.integers(0, 40, n_points); lens[:4] = [0, 1, 2, 128]
    ptr = np.zeros(n_points + 1, np.int32); ptr[1:] = np.cumsum(lens)
    base = rng.integers(0, 256, (n_points, 32), dtype=np.uint8)
    desc = np.repeat(base, lens, axis=0)
    flip = rng.random(desc.shape) < 0.04                      # observations = noisy copies of the point's descriptor
    desc = desc ^ (flip * rng.integers(1, 256, desc.shape)).astype(np.uint8)
    desc[ptr[5]:ptr[6]] = desc[ptr[5]]                        # all identical: every median is 0, the first row wins
    return desc, ptr


def test_distinctive_descriptor_rule(oracle_mod):
    rng = np.random.default_rng(2)
    desc, ptr = _distinct_case(rng)
    got = oracle_mod.distinctive(desc, ptr)
    for p in range(len(ptr) - 1):
        d = desc[ptr[p]:ptr[p + 1]]
        if len(d) == 0:
            assert got[p] == -1
            continue
        m = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2)
        med = np.sort(m, 1)[:, int(0.5 * (len(d) - 1))]
        assert got[p] == int(np.argmin(med))                  # least median, first on ties
