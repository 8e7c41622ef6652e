"""Pin the ORB oracle against cv2 (OpenCV 4.13.0) and emit the golden fixtures.

TEST INFRASTRUCTURE.  Run in the build container (cv2 importable):

    python -m oracle.crosscheck_cv2            # check every primitive, print a report
    python -m oracle.crosscheck_cv2 --golden   # additionally (re)write tests/golden/*.npz

OpenCV is the un-vendored third-party dependency that owns most of the extractor's arithmetic
(SURVEY.md section 8c); the reference calls it at src/ORBextractor.cc:104, 812-824, 1100, 1131, 1139-1152.
This script checks the oracle's restatement of each of those calls against the real library
and then runs an independent cv2-composed pipeline (pyramid, per-cell FAST with the ini/min
fallback, orientation, blur, rBRIEF) against ``oracle.orb_extract`` end to end.  Only the
quad-tree (pure reference logic, no OpenCV) is shared between the two pipelines; it is
cross-checked separately against an independently written sort-based formulation in
tests/test_oracle_orb.py.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

import oracle
from airdos_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _images(rng):
    yield "noise", rng.integers(0, 256, (97, 131), dtype=np.uint8)
    yield "synth", synth.make_stereo_pair(0)[0]
    yield "synth_small", synth.make_stereo_pair(3, 320, 240)[0]
    yield "flat", np.full((64, 80), 77, np.uint8)
    g = (np.add.outer(np.arange(120), np.arange(160)) % 256).astype(np.uint8)
    yield "ramp", g


def check_primitives(cv2) -> int:
    rng = np.random.default_rng(7)
    bad = 0
    for name, img in _images(rng):
        h, w = img.shape
        # copyMakeBorder REFLECT_101
        ref = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
        bad += int((ref != oracle.border101(img)).sum())
        # resize: every ratio the pyramid uses plus odd ones
        for s in (1.2, 1.2 ** 2, 1.37, 2.0, 3.1):
            dw, dh = int(round(w / s)), int(round(h / s))
            ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
            bad += int((ref != oracle.resize(img, dw, dh)).sum())
        # GaussianBlur 7x7 sigma 2
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        bad += int((ref != oracle.blur7(img)).sum())
        # erode 10x10 on a binary-ish mask
        m = np.where(rng.random((h, w)) < 0.02, 0, 255).astype(np.uint8)
        ref = cv2.erode(m, np.ones((10, 10), np.uint8))
        bad += int((ref != oracle.erode10(m)).sum())
        # FAST at the thresholds the configs use, with and without mask
        for t in (7, 12, 20):
            det = cv2.FastFeatureDetector_create(t, True)
            ref = np.array([[k.pt[0], k.pt[1], k.response] for k in det.detect(img, None)], np.float64).reshape(-1, 3)
            got = oracle.fast(img, t)
            if ref.shape != got.shape or (ref != got).any():
                bad += 1
                print("  FAST mismatch", name, t, ref.shape, got.shape)
            mk = synth.make_mask(11, w, h, 2) if w >= 240 else m
            ref = np.array([[k.pt[0], k.pt[1], k.response] for k in det.detect(img, mk)], np.float64).reshape(-1, 3)
            got = oracle.fast(img, t, mk)
            if ref.shape != got.shape or (ref != got).any():
                bad += 1
                print("  FAST(mask) mismatch", name, t, ref.shape, got.shape)
        print(f"  primitives on {name:12s} {w}x{h}: cumulative mismatches {bad}")
    # FAST on many cell-sized random sub-images (the shape the extractor actually calls it on)
    for i in range(300):
        cw, ch = int(rng.integers(7, 44)), int(rng.integers(7, 44))
        img = (rng.integers(0, 256, (ch, cw)) if i % 2 else
               np.clip(rng.normal(128, 40, (ch, cw)), 0, 255)).astype(np.uint8)
        det = cv2.FastFeatureDetector_create(int(rng.integers(1, 40)), True)
        ref = np.array([[k.pt[0], k.pt[1], k.response] for k in det.detect(img, None)], np.float64).reshape(-1, 3)
        got = oracle.fast(img, det.getThreshold())
        if ref.shape != got.shape or (ref != got).any():
            bad += 1
    # fastAtan2 on integer moment pairs
    ys = rng.integers(-200000, 200000, 20000).astype(np.float32)
    xs = rng.integers(-200000, 200000, 20000).astype(np.float32)
    ys[:4] = [0, 0, 1, -1]; xs[:4] = [0, 5, 0, 0]
    for y, x in zip(ys, xs):
        a, b = np.float32(cv2.fastAtan2(float(y), float(x))), np.float32(oracle.fast_atan2(y, x))
        if a.tobytes() != b.tobytes():
            bad += 1
    print(f"  after random cells + fastAtan2: cumulative mismatches {bad}")
    return bad


# ----------------------------------------------------------------------------------------
# Independent cv2-composed pipeline (follows src/ORBextractor.cc:767-864, 1054-1156).
def cv2_pipeline(cv2, img, mask, nfeatures, scale, nlevels, ini_th, min_th):
    from include_pattern import PATTERN_X, PATTERN_Y  # noqa: E402  (generated below)
    p = oracle.orb_params(nfeatures, scale, nlevels, img.shape[1], img.shape[0])
    E = 19
    pyr, mpyr = [], []
    for l in range(nlevels):
        sz = (int(p["w"][l]), int(p["h"][l]))
        if l == 0:
            full = cv2.copyMakeBorder(img, E, E, E, E, cv2.BORDER_REFLECT_101)
            if mask is not None:
                mfull = cv2.copyMakeBorder(cv2.erode(mask, np.ones((10, 10), np.uint8)), E, E, E, E, cv2.BORDER_REFLECT_101)
        else:
            r = cv2.resize(pyr[l - 1][E:-E, E:-E], sz, interpolation=cv2.INTER_LINEAR)
            full = cv2.copyMakeBorder(r, E, E, E, E, cv2.BORDER_REFLECT_101)
            if mask is not None:
                r = cv2.resize(mpyr[l - 1][E:-E, E:-E], sz, interpolation=cv2.INTER_LINEAR)
                mfull = cv2.copyMakeBorder(r, E, E, E, E, cv2.BORDER_REFLECT_101)
        pyr.append(full)
        if mask is not None:
            mpyr.append(mfull)
    kps_all, desc_all = [], []
    umax = p["umax"]
    for l in range(nlevels):
        roi = pyr[l][E:-E, E:-E]
        mroi = mpyr[l][E:-E, E:-E] if mask is not None else None
        lh, lw = roi.shape
        minB, maxBX, maxBY = 16, lw - 16, lh - 16
        width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
        ncols, nrows = int(width / np.float32(30)), int(height / np.float32(30))
        wcell, hcell = int(np.ceil(width / ncols)), int(np.ceil(height / nrows))
        cand = []
        for i in range(nrows):
            iniy = minB + i * hcell
            maxy = iniy + hcell + 6
            if iniy >= maxBY - 3:
                continue
            maxy = min(maxy, maxBY)
            for j in range(ncols):
                inix = minB + j * wcell
                maxx = inix + wcell + 6
                if inix >= maxBX - 6:
                    continue
                maxx = min(maxx, maxBX)
                sub = roi[iniy:maxy, inix:maxx]
                msub = mroi[iniy:maxy, inix:maxx] if mroi is not None else None
                k = cv2.FastFeatureDetector_create(ini_th, True).detect(sub, msub)
                if len(k) == 0:
                    k = cv2.FastFeatureDetector_create(min_th, True).detect(sub, msub)
                for kp in k:
                    cand.append((kp.pt[0] + j * wcell, kp.pt[1] + i * hcell, kp.response))
        cand = np.array(cand, np.float32).reshape(-1, 3)
        kept = oracle.distribute(cand, minB, maxBX, minB, maxBY, int(p["quota"][l]))
        if len(kept) == 0:
            continue
        blurred = cv2.GaussianBlur(roi.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        sc = np.float32(p["scale"][l])
        for x, y, resp in kept:
            x, y = np.float32(x + minB), np.float32(y + minB)
            cx, cy = int(np.rint(x)), int(np.rint(y))
            m01 = m10 = 0
            for u in range(-15, 16):
                m10 += u * int(roi[cy, cx + u])
            for v in range(1, 16):
                d = int(umax[v])
                plus = roi[cy + v, cx - d:cx + d + 1].astype(np.int64)
                minus = roi[cy - v, cx - d:cx + d + 1].astype(np.int64)
                us = np.arange(-d, d + 1)
                m01 += v * int((plus - minus).sum())
                m10 += int((us * (plus + minus)).sum())
            ang = np.float32(cv2.fastAtan2(float(np.float32(m01)), float(np.float32(m10))))
            rad = np.float32(ang * np.float32(np.float32(np.pi) / np.float32(180.0)))
            a, b = np.float32(np.cos(np.float64(rad))), np.float32(np.sin(np.float64(rad)))
            px, py = PATTERN_X.astype(np.float32), PATTERN_Y.astype(np.float32)
            rr = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int64)
            cc = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int64)
            vals = blurred[cy + rr, cx + cc].astype(np.int32)
            bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
            desc = np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)
            ox, oy = (x, y) if l == 0 else (np.float32(x * sc), np.float32(y * sc))
            kps_all.append((ox, oy, np.float32(int(np.float32(31) * sc)), ang, np.float32(resp), l))
            desc_all.append(desc)
    kps = np.array(kps_all, dtype=oracle.KP_DTYPE)
    desc = np.array(desc_all, np.uint8).reshape(-1, 32)
    return kps, desc, [q[E:-E, E:-E] for q in pyr]


def _install_pattern_module():
    import re
    import types
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "airdos_orb_pattern.h")).read()
    def grab(tag):
        body = hdr.split("#define " + tag)[1].split("}")[0]
        return np.array([int(t) for t in re.findall(r"-?\d+", body)], np.int32)
    m = types.ModuleType("include_pattern")
    m.PATTERN_X, m.PATTERN_Y = grab("AIRDOS_ORB_PATTERN_X"), grab("AIRDOS_ORB_PATTERN_Y")
    assert len(m.PATTERN_X) == 512 and len(m.PATTERN_Y) == 512
    sys.modules["include_pattern"] = m


def check_pipeline(cv2, write_golden: bool) -> int:
    _install_pattern_module()
    bad = 0
    cases = [
        ("cfg1_left_640x480_1000", synth.make_stereo_pair(0)[0], None, 1000, 12, 7),
        ("cfg1_right_640x480_1000", synth.make_stereo_pair(0)[1], None, 1000, 12, 7),
        ("mask_640x480_1000", synth.make_stereo_pair(1)[0], synth.make_mask(5), 1000, 12, 7),
        ("shipped_640x360_1500", synth.make_stereo_pair(2, 640, 360)[0], None, 1500, 12, 7),
        ("kitti_th_320x240_500", synth.make_stereo_pair(4, 320, 240)[0], None, 500, 20, 7),
    ]
    for name, img, mask, nf, ini, mn in cases:
        kps, desc, pyr = cv2_pipeline(cv2, img, mask, nf, 1.2, 8, ini, mn)
        o = oracle.orb_extract(img, mask, nf, 1.2, 8, ini, mn, want_pyramid=True)
        pb = sum(int((a != b).sum()) for a, b in zip(pyr, o["pyramid"]))
        ok = (len(kps) == len(o["kps"]) and kps.tobytes() == o["kps"].tobytes() and (desc == o["desc"]).all())
        print(f"  pipeline {name:28s} n={len(kps):5d}/{len(o['kps']):5d} pyramid_mismatch={pb} kp+desc_equal={ok}")
        bad += pb + (0 if ok else 1)
        if write_golden and ok and pb == 0:
            os.makedirs(GOLDEN, exist_ok=True)
            np.savez_compressed(os.path.join(GOLDEN, f"orb_{name}.npz"), image=img,
                                mask=mask if mask is not None else np.zeros((0, 0), np.uint8),
                                params=np.array([nf, 8, ini, mn], np.int32), scale=np.float32(1.2),
                                kps=kps, desc=desc, pyramid_sizes=np.array([q.shape for q in pyr], np.int32),
                                pyramid_sums=np.array([int(q.astype(np.int64).sum()) for q in pyr], np.int64),
                                source="cv2 %s composed pipeline (oracle/crosscheck_cv2.py)" % cv2.__version__)
    return bad


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", action="store_true")
    a = ap.parse_args()
    import cv2
    cv2.setNumThreads(1)
    print("cv2", cv2.__version__)
    bad = check_primitives(cv2)
    bad += check_pipeline(cv2, a.golden)
    print("TOTAL MISMATCHES:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
