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


def _stereo_numpy(kl, dl, kr, dr, pyr_l, pyr_r, scale, mb, mbf):
    """Frame::ComputeStereoMatches (src/Frame.cc:829-1003) written again from the reference in plain Python / numpy float32
    (bit counting by unpackbits, SAD windows as array slices): an independent check of oracle/match_oracle.cpp."""
    f32 = np.float32
    n_l = len(kl)
    u_right = np.full(n_l, -1, f32); depth = np.full(n_l, -1, f32)
    inv_scale = (f32(1.0) / scale).astype(f32)
    n_rows = pyr_l[0].shape[0]
    rows = [[] for _ in range(n_rows)]
    for i_r in range(len(kr)):
        y = kr["y"][i_r]; r = f32(2.0) * scale[kr["octave"][i_r]]
        hi = min(int(np.ceil(f32(y + r))), n_rows - 1); lo = max(int(np.floor(f32(y - r))), 0)
        for yi in range(lo, hi + 1):
            rows[yi].append(i_r)
    bits_l = np.unpackbits(dl, axis=1).astype(np.int16); bits_r = np.unpackbits(dr, axis=1).astype(np.int16)
    min_d = f32(0); max_d = f32(f32(mbf) / f32(mb))
    dist_idx = []
    for i_l in range(n_l):
        lvl = int(kl["octave"][i_l]); v_l = kl["y"][i_l]; u_l = kl["x"][i_l]
        cands = rows[int(v_l)]
        if not cands:
            continue
        min_u = f32(u_l - max_d); max_u = f32(u_l - min_d)
        if max_u < 0:
            continue
        best, best_r = 100, 0
        for i_r in cands:
            o = int(kr["octave"][i_r])
            if o < lvl - 1 or o > lvl + 1:
                continue
            u_r = kr["x"][i_r]
            if min_u <= u_r <= max_u:
                d = int(np.abs(bits_l[i_l] - bits_r[i_r]).sum())
                if d < best:
                    best, best_r = d, i_r
        if best >= 75:
            continue
        sf = inv_scale[lvl]

        def rnd(x):     # C round(): half away from zero
            return float(np.floor(abs(float(x)) + 0.5) * (1 if x >= 0 else -1))
        su_l = rnd(f32(u_l * sf)); sv_l = rnd(f32(v_l * sf)); su_r0 = rnd(f32(kr["x"][best_r] * sf))
        img_l, img_r = pyr_l[lvl], pyr_r[lvl]
        w, big_l = 5, 5
        cy, cx = int(sv_l), int(su_l)
        ini_u = su_r0 + big_l - w; end_u = su_r0 + big_l + w + 1
        if ini_u < 0 or end_u >= img_r.shape[1]:
            continue
        win_l = img_l[cy - w:cy + w + 1, cx - w:cx + w + 1].astype(f32)
        win_l = win_l - win_l[w, w]
        best_sad, best_inc, dists = 2 ** 31 - 1, 0, np.zeros(2 * big_l + 1, f32)
        for inc in range(-big_l, big_l + 1):
            cxr = int(su_r0) + inc
            win_r = img_r[cy - w:cy + w + 1, cxr - w:cxr + w + 1].astype(f32)
            win_r = win_r - win_r[w, w]
            dist = f32(np.abs((win_l - win_r).astype(np.float64)).sum())     # cv::norm(NORM_L1): double accumulation
            if dist < f32(best_sad):
                best_sad, best_inc = int(dist), inc
            dists[big_l + inc] = dist
        if best_inc in (-big_l, big_l):
            continue
        d1, d2, d3 = dists[big_l + best_inc - 1], dists[big_l + best_inc], dists[big_l + best_inc + 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
        if delta < -1 or delta > 1 or np.isnan(delta):
            continue
        best_u = f32(scale[lvl] * f32(f32(f32(su_r0) + f32(best_inc)) + delta))
        disp = f32(u_l - best_u)
        if min_d <= disp < max_d:
            if disp <= 0:
                disp = f32(0.01); best_u = f32(np.float64(u_l) - 0.01)
            depth[i_l] = f32(f32(mbf) / disp); u_right[i_l] = best_u
            dist_idx.append((best_sad, i_l))
    if dist_idx:
        dist_idx.sort()
        median = f32(dist_idx[len(dist_idx) // 2][0])
        th = f32(f32(f32(1.5) * f32(1.4)) * median)
        for sad, i_l in reversed(dist_idx):
            if f32(sad) < th:
                break
            u_right[i_l] = -1; depth[i_l] = -1
    return u_right, depth


def test_stereo_oracle_matches_an_independent_numpy_restatement(oracle_mod):
    from airdos_b200 import synth
    for frame in (1, 2):
        L, R = synth.make_stereo_pair(frame)
        a = oracle_mod.orb_extract(L, None, 1000, 1.2, 8, 12, 7, want_pyramid=True)
        b = oracle_mod.orb_extract(R, None, 1000, 1.2, 8, 12, 7, want_pyramid=True)
        sc = np.asarray(oracle_mod.orb_params(1000, 1.2, 8, 640, 480)["scale"], np.float32)
        mbf = synth.BF; mb = mbf / synth.FX
        ur, dp, _, _ = oracle_mod.stereo_match(a["kps"], a["desc"], b["kps"], b["desc"], a["pyramid"], b["pyramid"], sc, mb, mbf)
        ur2, dp2 = _stereo_numpy(a["kps"], a["desc"], b["kps"], b["desc"], a["pyramid"], b["pyramid"], sc, mb, mbf)
        assert (dp > 0).sum() > 200
        assert (ur.view(np.uint32) == ur2.view(np.uint32)).all() and (dp.view(np.uint32) == dp2.view(np.uint32)).all()
