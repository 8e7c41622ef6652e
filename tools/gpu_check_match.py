"""CUDA matcher vs oracle (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import airdos_b200 as adb
from airdos_b200 import synth

oracle.build()
ok = True
F = 3
pairs = synth.make_stereo_batch(F)
exL = adb.ORBextractor(2000, 1.2, 8, 12, 7, 640, 480, max_batch=F)
exR = adb.ORBextractor(2000, 1.2, 8, 12, 7, 640, 480, max_batch=F)
kl, dl, cl = exL.extract_batch(pairs[:, 0])
kr, dr, cr = exR.extract_batch(pairs[:, 1])
mbf = synth.BF; mb = mbf / synth.FX
ur, dp, bi, bd = adb.compute_stereo_matches(exL, exR, F, mb, mbf)
for f in range(F):
    nl, nr = cl[f], cr[f]
    o = oracle.stereo_match(kl[f, :nl], dl[f, :nl], kr[f, :nr], dr[f, :nr], exL.pyramid(f), exR.pyramid(f),
                            np.array(exL.GetScaleFactors(), np.float32), mb, mbf)
    e = [bool((a[f, :nl].view(np.uint32) == b.view(np.uint32)).all()) for a, b in zip((ur, dp, bi, bd), o)]
    print(f"stereo f{f}: nL={nl} nR={nr} matched={(o[1] > 0).sum()} ham_matched={(o[2] >= 0).sum()} uRight/depth/idx/dist equal: {e}")
    if not all(e):
        ok = False
        for name, a, b in zip(("ur", "dp", "bi", "bd"), (ur, dp, bi, bd), o):
            bad = np.nonzero(a[f, :nl] != b)[0]
            print("   ", name, len(bad), bad[:6], a[f, bad[:6]], b[bad[:6]])
# best2: all pairs and candidate lists
m = adb.ORBmatcher()
rng = np.random.default_rng(3)
q, t = dl[0, :cl[0]], dr[0, :cr[0]]
for name, off, idx in [("allpairs", None, None)]:
    a = m.best2(q, t); b = oracle.best2(q, t)
    e = [bool((x == y).all()) for x, y in zip(a, b)]
    print("best2", name, e); ok &= all(e)
lens = rng.integers(0, 40, len(q)); off = np.zeros(len(q) + 1, np.int32); off[1:] = np.cumsum(lens)
idx = rng.integers(0, len(t), off[-1]).astype(np.int32)
a = m.best2(q, t, off, idx); b = oracle.best2(q, t, off, idx)
e = [bool((x == y).all()) for x, y in zip(a, b)]
print("best2 lists", e); ok &= all(e)
# ties: many duplicate descriptors
t2 = np.repeat(t[:50], 8, axis=0)
a = m.best2(q[:300], t2); b = oracle.best2(q[:300], t2)
e = [bool((x == y).all()) for x, y in zip(a, b)]
print("best2 ties", e); ok &= all(e)
print("ALL OK" if ok else "MISMATCHES")
