"""Writes tests/golden/lba_ref.npz: results of the REFERENCE's own Optimizer::LocalBundleAdjustment (src/Optimizer.cc:431-731, the whole
function compiled from /root/reference by `make -C oracle ref`, oracle/ref_lba.cpp) on seeded windows.  The function runs with the
reference's own g2o vertex / edge types (gates: chi2(), isDepthPositive()), Converter functions and Levenberg-Marquardt control
(oracle/ref_lm.cpp); its solver steps are the oracle's (see ref_lba.cpp's header).  The fixture keeps the window, the problem the
function built (vertex order, fixed flags, edges, information, Huber deltas), every LM trial, the final estimates, the erase list and
the poses / positions it wrote back.  Run in the build container (needs /root/reference):

    python oracle/gen_ref_lba_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")

# (seed, local key-frames, points, fixed observers, mono fraction, key-frame with mnId 0 inside the window, point noise)
CASES = [(3, 8, 300, 2, 0.2, True, 0.0), (4, 6, 200, 0, 0.0, True, 0.0), (5, 10, 400, 3, 0.5, False, 0.0), (6, 5, 150, 1, 1.0, False, 0.0),
         (7, 7, 250, 2, 0.3, True, 0.6)]


def make_window(i):
    """A covisibility window in the reference's terms: key-frame 0 is pKF (the newest), `covisible` its neighbours, the remaining key-frames
    are found by the function itself through the observations (fixed cameras).  Poses go through float 4 x 4 matrices (KeyFrame::GetPose),
    positions through float vectors, key-points carry the octave whose mvInvLevelSigma2 becomes the edge information."""
    import oracle
    from airdos_b200 import synth
    seed, n_kf, n_pts, n_fixed, mono, zero_in, pn = CASES[i]
    d = synth.make_ba_problem(n_kf, n_pts, 5, seed=seed, mono_frac=mono, n_fixed_extra=n_fixed)
    rng = np.random.default_rng(seed)
    d["points"] = d["points"] + rng.normal(0, pn, d["points"].shape)
    K = len(d["pose_t"])
    cur = n_kf - 1
    order = [cur] + [k for k in range(n_kf) if k != cur] + list(range(n_kf, K))
    pos = {k: j for j, k in enumerate(order)}
    ids = np.array([(k if zero_in else k + 3) * 2 for k in order], np.int32)       # mnId 0 (when present) is fixed by the function itself
    lib = oracle.ba_lib()
    lib.ba_oracle_pose_to_tcw.argtypes = [C.c_void_p] * 3
    tcw = np.zeros((K, 4, 4), np.float32)
    for j, k in enumerate(order):
        q = np.ascontiguousarray(d["pose_q"][k]); t = np.ascontiguousarray(d["pose_t"][k]); T = np.zeros(16, np.float32)
        lib.ba_oracle_pose_to_tcw(q.ctypes.data, t.ctypes.data, T.ctypes.data)
        tcw[j] = T.reshape(4, 4)
    sig = oracle.orb_tables(2000, 1.2, 8)["inv_sigma2"]                             # mvInvLevelSigma2 as the extractor's constructor leaves it
    octave = np.array([int(np.argmin(np.abs(sig - np.float32(w)))) for w in d["edge_info"]], np.int32)
    return dict(kf_id=ids, kf_tcw=tcw, covisible=np.arange(1, n_kf, dtype=np.int32), fx=d["fx"], fy=d["fy"], cx=d["cx"], cy=d["cy"], bf=d["bf"],
                inv_level_sigma2=sig, mp_id=np.arange(len(d["points"]), dtype=np.int32) * 3 + 1, mp_pos=d["points"].astype(np.float32),
                obs_kf=np.array([pos[k] for k in d["edge_pose"]], np.int32), obs_mp=d["edge_point"].astype(np.int32),
                obs_uvr=d["edge_obs"].astype(np.float32), obs_octave=octave)


# Optimizer::BundleAdjustment (src/Optimizer.cc:60-230, called by GlobalBundleAdjustemnt :52-58): (window case, nIterations, nLoopKF, bRobust)
GBA_CASES = [(0, 10, 0, True), (1, 5, 0, False), (2, 20, 7, True), (3, 10, 0, True)]


def make_gba_window(j):
    """The window of case GBA_CASES[j][0] plus one map point nobody observes (the function removes its vertex, :178-180)."""
    w = make_window(GBA_CASES[j][0])
    w = dict(w)
    w["mp_id"] = np.concatenate([w["mp_id"], [int(w["mp_id"].max()) + 5]]).astype(np.int32)
    w["mp_pos"] = np.concatenate([w["mp_pos"], [[1.0, 2.0, 30.0]]]).astype(np.float32)
    return w


def gba_main(LBA, LM, out):
    import oracle
    for j, (c, its, loop_kf, robust) in enumerate(GBA_CASES):
        w = make_gba_window(j)
        r = oracle.ref_local_bundle_adjustment(LBA, LM, w, global_ba=(its, loop_kf, robust))
        o = oracle.ba_global_options(its, robust)
        p, res, st = oracle.ba_solve(r["problem"], o)
        fs = np.concatenate([p["pose_q"].ravel(), p["pose_t"].ravel(), p["points"].ravel()])
        tr = res.trace_rows[:, [0, 1, 2, 4]]
        print(f"global BA {j}: {len(r['pose_id'])} key-frames ({int(r['problem']['pose_fixed'].sum())} fixed), {len(r['point_id'])} points, "
              f"{len(r['problem']['edge_pose'])} edges, {its} iterations asked, {list(r['round_iterations'])} run, robust {list(r['round_robust'])}, "
              f"Huber deltas {r['huber']} (oracle {o.huber_mono, o.huber_stereo}) | oracle: trials identical "
              f"{tr.shape == r['rows'].shape and bool((tr == r['rows']).all())}, state identical {bool((fs == r['final_state']).all())}")
        for k in ("kf_tcw", "mp_pos", "mp_updates", "rows", "final_state", "pose_id", "point_id", "round_iterations", "round_robust"):
            out[f"g{j}_{k}"] = r[k]
        out[f"g{j}_huber"] = np.array(r["huber"])
        for k, v in r["problem"].items():
            out[f"g{j}_p_{k}"] = np.asarray(v)


# Optimizer::LocalBundleAdjustmentHumanTrajactory (src/Optimizer.cc:1496-2222):
# (seed, local key-frames, points, fixed observers, trajectories, poses per trajectory, joints displaced by ~1 m, first pose's key-frame outside the window)
HBA_CASES = [(9, 12, 600, 2, 2, 4, 0, False), (10, 14, 500, 1, 3, 5, 4, False), (12, 10, 300, 0, 2, 4, 2, True), (13, 12, 400, 2, 2, 3, 0, False)]
HBA_SIGMAS = dict(sigma_human=0.5, sigma_rigidity=20.0, sigma_motion=20.0, th_motion=4.0, th_rigidity=1.0)


def make_human_window(c):
    """A covisibility window with human trajectories in the reference's terms (MapHumanTrajectory -> MapHumanPose -> 14 MapHumanKey, the
    14 Rigidbody bone lengths of a trajectory, the segment table Map::body1 / body2): every observation stereo (the function builds
    EdgeStereoSE3ProjectXYZ only), optionally joints displaced so that rigidity / motion edges end up as outliers, and optionally a
    trajectory whose first pose hangs on a key-frame outside the window (its vertices are never created; the motion edges that would
    touch them are skipped).  Trajectories of 3 poses are not longer than Map::thLongTrajectory and are left out by the function."""
    import oracle
    from airdos_b200 import synth
    seed, n_kf, n_pts, n_fixed, H, S, n_displaced, outside = HBA_CASES[c][:8]
    obs_per_point = HBA_CASES[c][8] if len(HBA_CASES[c]) > 8 else 5
    d = synth.make_ba_problem(n_kf, n_pts, obs_per_point, seed=seed, mono_frac=0.0, n_fixed_extra=n_fixed, humans=H, human_poses=S)
    d["edge_obs"][:, 2] = np.where(d["edge_obs"][:, 2] < 0, 0.25, d["edge_obs"][:, 2])
    rng = np.random.default_rng(seed + 500)
    for j in rng.choice(len(d["joints"]), n_displaced, replace=False):
        d["joints"][j] += rng.normal(0, 0.7, 3)
    K = len(d["pose_t"]); cur = n_kf - 1
    order = [cur] + [k for k in range(n_kf) if k != cur] + list(range(n_kf, K))
    nh = H * S
    ref_kf = d["jedge_pose"].reshape(nh, 14)[:, 0].copy()
    extra = []
    if outside:                                    # one more key-frame that nothing else observes: not local, not fixed
        extra = [K]
        ref_kf[0] = K
    pos = {k: j for j, k in enumerate(order + extra)}
    ids = np.array([k * 2 for k in order + extra], np.int32)
    lib = oracle.ba_lib()
    lib.ba_oracle_pose_to_tcw.argtypes = [C.c_void_p] * 3
    tcw = np.zeros((len(ids), 4, 4), np.float32)
    for j, k in enumerate(order + extra):
        kk = min(k, K - 1)
        q = np.ascontiguousarray(d["pose_q"][kk]); t = np.ascontiguousarray(d["pose_t"][kk]); T = np.zeros(16, np.float32)
        lib.ba_oracle_pose_to_tcw(q.ctypes.data, t.ctypes.data, T.ctypes.data)
        tcw[j] = T.reshape(4, 4)
    sig = oracle.orb_tables(2000, 1.2, 8)["inv_sigma2"]
    octave = np.array([int(np.argmin(np.abs(sig - np.float32(w)))) for w in d["edge_info"]], np.int32)
    w = dict(kf_id=ids, kf_tcw=tcw, covisible=np.arange(1, n_kf, dtype=np.int32), fx=d["fx"], fy=d["fy"], cx=d["cx"], cy=d["cy"], bf=d["bf"],
             inv_level_sigma2=sig, mp_id=np.arange(len(d["points"]), dtype=np.int32) * 3 + 1, mp_pos=d["points"].astype(np.float32),
             obs_kf=np.array([pos[k] for k in d["edge_pose"]], np.int32), obs_mp=d["edge_point"].astype(np.int32),
             obs_uvr=d["edge_obs"].astype(np.float32), obs_octave=octave)
    hum = dict(traj_id=np.arange(H) * 2 + 1, traj_track_id=np.arange(H) + 10, traj_n_poses=np.full(H, S), rigid_id=(np.arange(H * 14) + 3).reshape(H, 14),
               rigid_dist=d["dists"].reshape(H, 14).astype(np.float32), hp_id=np.arange(nh) + 100, hp_traj=np.repeat(np.arange(H), S),
               hp_ref_kf=np.array([pos[int(k)] for k in ref_kf]), hp_time=np.tile(np.arange(S, dtype=np.float64), H),
               key_id=(np.arange(nh * 14) * 2 + 5).reshape(nh, 14), key_pos=d["joints"].reshape(nh, 14, 3).astype(np.float32),
               key_uvr=d["jedge_obs"].reshape(nh, 14, 3).astype(np.float32), current_hp=np.array([t * S + S - 1 for t in range(H)]), **HBA_SIGMAS)
    return w, hum


def hba_options(oracle, r):
    o = oracle.ba_default_options()
    o.huber_mono, o.huber_stereo = r["huber"]
    o.huber_rigid, o.huber_motion = r["huber_rigid"], r["huber_motion"]
    o.chi2_rigid, o.chi2_motion = HBA_SIGMAS["th_rigidity"], HBA_SIGMAS["th_motion"]
    return o


def hba_main(LBA, LM, out):
    import oracle
    for c in range(len(HBA_CASES)):
        w, hum = make_human_window(c)
        r = oracle.ref_local_bundle_adjustment(LBA, LM, w, humans=hum)
        p = r["problem"]
        line = (f"human window {c}: {len(r['pose_id'])} key-frames, {len(r['point_id'])} points, {len(r['joint_id'])} joints, {len(r['dist_id'])} bone lengths, "
                f"{len(r['motion_id'])} motions; edges {len(p['edge_pose'])} static / {len(p['jedge_pose'])} joint / {len(p['redge_i'])} rigidity / "
                f"{len(p['medge_p1'])} motion; rounds {list(r['round_iterations'])}, {len(r['rows'])} trials")
        if len(r["round_iterations"]):
            pp, res, st = oracle.ba_solve(p, hba_options(oracle, r))
            tr = res.trace_rows[:, [0, 1, 2, 4]]
            fs = np.concatenate([pp[k].ravel() for k in ("pose_q", "pose_t", "points", "joints", "dists", "motion_q", "motion_t")])
            kind, chi, dep = r["edge_kind"], r["edge_final_chi2"], r["edge_final_depth_positive"]
            flags = (bool((res.jedge_outlier == ((chi > 7.815) | (dep == 0))[kind == 2]).all()) and bool((res.redge_outlier == (chi > HBA_SIGMAS["th_rigidity"])[kind == 3]).all())
                     and bool((res.medge_outlier == (chi > HBA_SIGMAS["th_motion"])[kind == 4]).all()) and int(res.edge_outlier.sum()) == len(r["erased"]))
            line += (f" | oracle: trials identical {tr.shape == r['rows'].shape and bool((tr == r['rows']).all())}, state identical "
                     f"{fs.shape == r['final_state'].shape and bool((fs == r['final_state']).all())}, gates identical {flags} "
                     f"(outliers {int(res.edge_outlier.sum())} static, {int(res.jedge_outlier.sum())} joint, {int(res.redge_outlier.sum())} rigidity, {int(res.medge_outlier.sum())} motion)")
        print(line)
        for k in ("kf_tcw", "mp_pos", "erased", "mp_updates", "rows", "final_state", "pose_id", "point_id", "joint_id", "dist_id", "motion_id", "round_iterations",
                  "round_robust", "edge_kind", "edge_final_chi2", "edge_final_depth_positive"):
            out[f"h{c}_{k}"] = r[k]
        out[f"h{c}_huber"] = np.array(list(r["huber"]) + [r["huber_rigid"], r["huber_motion"]])
        for k, v in r["humans"].items():
            out[f"h{c}_hum_{k}"] = np.asarray(v)
        for k, v in p.items():
            out[f"h{c}_p_{k}"] = np.asarray(v)


def main():
    import oracle
    oracle.build()
    LM = C.CDLL(os.path.join(REF, "libref_lm.so")); LBA = C.CDLL(os.path.join(REF, "libref_lba.so"))
    out = {}
    for i in range(len(CASES)):
        w = make_window(i)
        r = oracle.ref_local_bundle_adjustment(LBA, LM, w)
        o = oracle.ba_default_options(); o.huber_mono, o.huber_stereo = r["huber"]
        p, res, st = oracle.ba_solve(r["problem"], o)
        fs = np.concatenate([p["pose_q"].ravel(), p["pose_t"].ravel(), p["points"].ravel()])
        tr = res.trace_rows[:, [0, 1, 2, 4]]
        flagged = np.flatnonzero(res.edge_outlier)
        print(f"window {i}: {len(r['pose_id'])} key-frames ({int(r['problem']['pose_fixed'].sum())} fixed), {len(r['point_id'])} points, "
              f"{len(r['problem']['edge_pose'])} edges, rounds {list(r['round_iterations'])} robust {list(r['round_robust'])}, {len(r['rows'])} trials, "
              f"{len(r['erased'])} erased | oracle: status {st}, trials identical {tr.shape == r['rows'].shape and bool((tr == r['rows']).all())}, "
              f"state identical {bool((fs == r['final_state']).all())}, outliers {len(flagged)}")
        for k in ("kf_tcw", "mp_pos", "erased", "mp_updates", "rows", "final_state", "pose_id", "point_id", "round_iterations", "round_robust"):
            out[f"w{i}_{k}"] = r[k]
        out[f"w{i}_huber"] = np.array(r["huber"])
        for k, v in r["problem"].items():
            out[f"w{i}_p_{k}"] = np.asarray(v)
    gba_main(LBA, LM, out)
    hba_main(LBA, LM, out)
    path = os.path.join(ROOT, "tests", "golden", "lba_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
