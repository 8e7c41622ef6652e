"""Stage-by-stage comparison of the CUDA extractor against the oracle (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import airdos_b200 as adb
from airdos_b200 import synth

def check(name, img, mask, nf, ini, mn):
    h, w = img.shape
    ex = adb.ORBextractor(nf, 1.2, 8, ini, mn, w, h, max_batch=2)
    imgs = np.stack([img, img[::-1].copy()])
    masks = None if mask is None else np.stack([mask, mask[::-1].copy()])
    t = time.time()
    kps, desc, cnt = ex.extract_batch(imgs, masks)
    dt = time.time() - t
    ok_all = True
    for f in range(2):
        o = oracle.orb_extract(imgs[f], None if masks is None else masks[f], nf, 1.2, 8, ini, mn, want_pyramid=True)
        pyr = ex.pyramid(f)
        pb = [int((a != b).sum()) for a, b in zip(pyr, o["pyramid"])]
        cc = [len(ex.debug_candidates(f, l)) for l in range(8)]
        n = cnt[f]
        same_n = n == len(o["kps"])
        kp_ok = same_n and kps[f, :n].tobytes() == o["kps"].tobytes()
        d_ok = same_n and bool((desc[f, :n] == o["desc"]).all())
        print(f"{name} f{f}: n={n}/{len(o['kps'])} pyr_mismatch={pb} cand={cc} oracle_cand={list(o['cand_counts'])} kp_ok={kp_ok} desc_ok={d_ok} ({dt*1e3:.1f} ms)")
        if not (kp_ok and d_ok):
            ok_all = False
            m = min(n, len(o["kps"]))
            a, b = kps[f, :m], o["kps"][:m]
            for fld in a.dtype.names:
                bad = np.nonzero(a[fld] != b[fld])[0]
                print("   field", fld, "mismatches", len(bad), "first", bad[:5], a[fld][bad[:5]], b[fld][bad[:5]])
            db = np.nonzero((desc[f, :m] != o["desc"][:m]).any(1))[0]
            print("   desc rows differing", len(db), db[:10])
    ex.close()
    return ok_all

if __name__ == "__main__":
    oracle.build()
    ok = True
    ok &= check("cfg1", synth.make_stereo_pair(0)[0], None, 1000, 12, 7)
    ok &= check("cfg2k", synth.make_stereo_pair(1)[1], None, 2000, 12, 7)
    ok &= check("mask", synth.make_stereo_pair(1)[0], synth.make_mask(5), 1000, 12, 7)
    ok &= check("shipped", synth.make_stereo_pair(2, 640, 360)[0], None, 1500, 12, 7)
    ok &= check("kitti", synth.make_stereo_pair(4, 320, 240)[0], None, 500, 20, 7)
    print("ALL OK" if ok else "MISMATCHES")
