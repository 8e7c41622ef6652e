import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import airdos_b200 as adb
from airdos_b200 import synth

def oracle_cands(roi, ini_th, min_th, mroi=None):
    lh, lw = roi.shape
    minB, maxBX, maxBY = 16, lw - 16, lh - 16
    width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
    ncols, nrows = int(width / np.float32(30)), int(height / np.float32(30))
    wcell, hcell = int(np.ceil(width / ncols)), int(np.ceil(height / nrows))
    cand = []
    for i in range(nrows):
        iniy = minB + i * hcell; maxy = iniy + hcell + 6
        if iniy >= maxBY - 3: continue
        maxy = min(maxy, maxBY)
        for j in range(ncols):
            inix = minB + j * wcell; maxx = inix + wcell + 6
            if inix >= maxBX - 6: continue
            maxx = min(maxx, maxBX)
            sub = roi[iniy:maxy, inix:maxx]
            msub = None if mroi is None else mroi[iniy:maxy, inix:maxx]
            k = oracle.fast(sub, ini_th, msub)
            if len(k) == 0: k = oracle.fast(sub, min_th, msub)
            for x, y, s in k: cand.append((x + j * wcell, y + i * hcell, s))
    return np.array(cand, np.int32).reshape(-1, 3)

oracle.build()
img = synth.make_stereo_pair(0)[0]
ex = adb.ORBextractor(1000, 1.2, 8, 12, 7, 640, 480, max_batch=1)
kps, desc = ex(img)
pyr = ex.pyramid(0)
for l in range(8):
    ours = ex.debug_candidates(0, l)
    ref = oracle_cands(pyr[l], 12, 7)
    so = set(map(tuple, ours)); sr = set(map(tuple, ref))
    print("level", l, "ours", len(ours), "ref", len(ref), "common", len(so & sr), "order_equal", len(ours) == len(ref) and bool((ours == ref).all()))
    extra = sorted(so - sr)[:8]; miss = sorted(sr - so)[:8]
    print("   extra", extra, "\n   missing", miss)
    if l == 0:
        # position-only overlap
        po = set((a, b) for a, b, c in ours); pr = set((a, b) for a, b, c in ref)
        print("   positions common", len(po & pr), "score hist ours", np.bincount(ours[:, 2] // 32, minlength=8), "ref", np.bincount(ref[:, 2] // 32, minlength=8))
