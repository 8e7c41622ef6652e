"""Writes tests/golden/pose_ref.npz: results of the REFERENCE's own Optimizer::PoseOptimization (src/Optimizer.cc:232-429, the whole function
compiled from /root/reference by `make -C oracle ref`, oracle/ref_lba.cpp) on seeded frames: the edges it builds from the frame's
correspondences (null map points skipped, mono where mvuRight < 0), its four rounds of optimize(10) restarted from mTcw, the float chi2
gates with the reference's own OnlyPose edge types (outliers re-evaluated at the round's pose, inliers with the error of the last
evaluated trial), the kernel dropped after the third round, the `edges().size() < 10` exit, mvbOutlier, the return value and the pose
written back.  LM control: the reference's (oracle/ref_lm.cpp); solver steps: the oracle's pose session.  Run in the build container:

    python oracle/gen_ref_pose_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
CASES = [(7, 600), (17, 300), (5, 40), (6, 9), (8, 2)]      # (seed, correspondences): 4 frames each (make_pose_frames: one of them with 7)


def make_frames(c):
    """Frames in the reference's terms: mTcw as a float 4 x 4, key-points with octaves, mvpMapPoints with every 11th entry null."""
    import oracle
    from airdos_b200 import synth
    seed, n = CASES[c]
    cam, frames, _ = synth.make_pose_frames(4, n, seed=seed)
    sig = oracle.orb_tables(2000, 1.2, 8)["inv_sigma2"]
    lib = oracle.ba_lib()
    lib.ba_oracle_pose_to_tcw.argtypes = [C.c_void_p] * 3
    out = []
    for fr in frames:
        q = np.ascontiguousarray(fr["pose_q"], np.float64); t = np.ascontiguousarray(fr["pose_t"], np.float64); T = np.zeros(16, np.float32)
        lib.ba_oracle_pose_to_tcw(q.ctypes.data, t.ctypes.data, T.ctypes.data)
        octave = np.array([int(np.argmin(np.abs(sig - np.float32(w)))) for w in fr["inv_sigma2"]], np.int32)
        has = np.ones(len(octave), np.uint8); has[::11] = 0
        out.append(dict(tcw=T.reshape(4, 4), inv_level_sigma2=sig, uvr=np.asarray(fr["obs"], np.float32), octave=octave, xw=np.asarray(fr["xw"], np.float32),
                        has_point=has, **cam))
    return out


def main():
    import oracle
    oracle.build()
    LM = C.CDLL(os.path.join(REF, "libref_lm.so")); LBA = C.CDLL(os.path.join(REF, "libref_lba.so"))
    out = {}
    same = total = 0
    for c in range(len(CASES)):
        for f, io in enumerate(make_frames(c)):
            r = oracle.ref_pose_optimization(LBA, LM, io)
            tag = f"c{c}f{f}"
            for k in ("outlier", "tcw", "rows", "final_state", "round_iterations", "round_robust", "round_active"):
                out[f"{tag}_{k}"] = r[k]
            out[f"{tag}_n_inliers"] = np.int32(r["n_inliers"])
            for k, v in r["frame"].items():
                out[f"{tag}_fr_{k}"] = v
            out[f"{tag}_cam"] = np.array([r["cam"][k] for k in ("fx", "fy", "cx", "cy", "bf")])
            total += 1
            if len(r["frame"]["pose_q"]):
                pb, rows = oracle.pose_optimize_traced(r["cam"], r["frame"])
                st = np.concatenate([pb.pose_q.ravel(), pb.pose_t.ravel()])
                ok = (rows.shape == r["rows"].shape and (rows == r["rows"]).all() and (st == r["final_state"]).all()
                      and (pb.outlier[:int(io["has_point"].sum())] == r["outlier"][io["has_point"] != 0]).all() and int(pb.n_inliers[0]) == r["n_inliers"])
            else:
                ok = r["n_inliers"] == 0                       # fewer than 3 correspondences: return 0 before any optimisation (:342-343)
            same += int(ok)
            print(f"{tag}: {int(io['has_point'].sum())} correspondences, rounds {list(r['round_iterations'])} active {list(r['round_active'])}, "
                  f"{len(r['rows'])} trials, returns {r['n_inliers']}: oracle {'identical' if ok else 'DIFFERENT'}")
    print(f"PoseOptimization: the oracle equals the reference function in {same} of {total} frames (trials, pose, mvbOutlier, return value)")
    path = os.path.join(ROOT, "tests", "golden", "pose_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
