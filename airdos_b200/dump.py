"""The reference's own map dump as a bundle-adjustment problem.

Tracking::SaveMap (src/Tracking.cc:1745-1838) writes, with default ostream formatting (6 significant digits):
    KF.txt      <KF id> <16 floats: Twc = KeyFrame::GetPoseInverse(), row major>\\n
    MP.txt      <MP id + maxKFid + 1> <x> <y> <z>\\n            (bad points included)
    Match.txt   <KF id> <MP id + maxKFid + 1> <u> <v> <u_right> <invSigma2>   -- and NO newline: line 1807 reads
                `<< invSigma2; '\\n';`, so the record separator is missing and invSigma2 runs into the next KF id
                ("0.69444412" = invSigma2 0.694444 followed by KF 12)
    HMTraj.txt  <track id> <pose id> <key id> <isBad> <isLost> <x> <y> <z>\\n   (14 joints per pose)
    Motion.txt  <track id> <16 floats: mTMotion>\\n
HMTraj.txt / Motion.txt carry the articulated part as far as the reference dumps it: joint positions with their bad / lost
flags and one motion per trajectory, but neither the joints' image observations nor the bone-length vertices nor time stamps.
`load_map_dump(..., humans=True)` therefore rebuilds what the dump determines -- joints, motions, the 14 rigidity edges per pose
(topology of include/Map.h:49-56, bone lengths initialised to the mean joint distance over the trajectory) and the 5 motion edges
per consecutive pose pair with delta_t = 1 -- and leaves the joint reprojection edges empty.
`load_map_dump` undoes the Match.txt bug with the level table: invSigma2 is one of 1 / scaleFactor^(2 l), whose printed
forms are prefix free, so the token splits uniquely.  `save_map_dump` writes the same files (bug included by default) so
that fixtures can be produced without running SLAM.  Host logic only; the problem dict is what
airdos_b200.ba.Optimizer.GlobalBundleAdjustemnt / LocalBundleAdjustment take.
"""
from __future__ import annotations

import os

import numpy as np


def _fmt(v) -> str:
    """operator<<(ostream&, float) with the default precision (6 significant digits, %g)."""
    return "%g" % float(np.float32(v))


def inv_sigma2_table(scale_factor: float = 1.2, n_levels: int = 8):
    """mvInvLevelSigma2 (src/ORBextractor.cc:418-432) as float32 and as the dump prints it."""
    sf = np.float32(1.0); vals = []
    for l in range(n_levels):
        if l:
            sf = np.float32(sf * np.float32(scale_factor))
        vals.append(np.float32(1.0) / np.float32(sf * sf))
    return np.array(vals, np.float32), [_fmt(v) for v in vals]


def _split_match_tokens(text: str, printed: list[str]):
    """Tokens of Match.txt with the run-together `<invSigma2><next KF id>` token split.  One pass: the glued-on KF id of the next
    record is carried in `pending` instead of being inserted into the token list (a real dump holds 10^5 - 10^6 records)."""
    out = []
    raw = text.split()
    by_len = sorted(printed, key=len, reverse=True)
    i, n = 0, len(raw)
    pending = None
    while i < n or pending is not None:
        if pending is not None:
            rec = [pending] + raw[i:i + 5]; i += 5
            pending = None
        else:
            rec = raw[i:i + 6]; i += 6
        if len(rec) < 6:
            raise ValueError("Match.txt: truncated record")
        tok = rec[5]
        best = next((p for p in by_len if tok.startswith(p)), None)      # longest printed level value that prefixes the token
        if best is None:
            raise ValueError(f"Match.txt: {tok!r} does not start with a level invSigma2 ({printed})")
        out.append(rec[:5] + [best])
        rest = tok[len(best):]
        if rest:                       # the bug: the next record's KF id is glued on
            pending = rest
    return out


def load_map_dump(path: str, cam: dict, scale_factor: float = 1.2, n_levels: int = 8, humans: bool = False) -> dict:
    """-> problem dict (float64 SoA of adb_ba_problem).  cam = dict(fx, fy, cx, cy, bf).  Poses: KF.txt holds Twc; the
    problem wants world -> camera as unit quaternion + translation (Converter::toSE3Quat of Tcw).  Key-frame 0 is fixed
    (src/Optimizer.cc:89); map points without observations are dropped like the reference skips them."""
    from . import ba
    kf_ids, poses_q, poses_t = [], [], []
    for line in open(os.path.join(path, "KF.txt")):
        f = line.split()
        if not f:
            continue
        Twc = np.array(f[1:17], np.float32).reshape(4, 4)
        Rwc, twc = Twc[:3, :3].astype(np.float64), Twc[:3, 3].astype(np.float64)
        Tcw = np.eye(4, dtype=np.float32)
        Tcw[:3, :3] = (Rwc.T).astype(np.float32); Tcw[:3, 3] = (-Rwc.T @ twc).astype(np.float32)
        q, t = ba.pose_from_tcw(Tcw)
        kf_ids.append(int(f[0])); poses_q.append(q); poses_t.append(t)
    kf_index = {k: i for i, k in enumerate(kf_ids)}
    mp_ids, pts = [], []
    for line in open(os.path.join(path, "MP.txt")):
        f = line.split()
        if f:
            mp_ids.append(int(f[0])); pts.append([np.float32(v) for v in f[1:4]])
    _, printed = inv_sigma2_table(scale_factor, n_levels)
    recs = _split_match_tokens(open(os.path.join(path, "Match.txt")).read(), printed)
    mp_pos = dict(zip(mp_ids, pts))
    used = sorted({int(r[1]) for r in recs if int(r[0]) in kf_index and int(r[1]) in mp_pos})   # a Match id missing from MP.txt is skipped
    mp_index = {m: i for i, m in enumerate(used)}
    ep, em, obs, info = [], [], [], []
    for r in recs:
        k, mid = int(r[0]), int(r[1])
        if k not in kf_index or mid not in mp_pos:
            continue
        ep.append(kf_index[k]); em.append(mp_index[mid])
        ur = np.float32(r[4])
        obs.append([np.float32(r[2]), np.float32(r[3]), ur if ur >= 0 else np.float32(-1)])
        info.append(np.float32(r[5]))
    n_p = len(kf_ids)
    out = dict(fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], bf=cam["bf"],
               pose_q=np.array(poses_q, np.float64).reshape(n_p, 4), pose_t=np.array(poses_t, np.float64).reshape(n_p, 3),
               pose_fixed=np.array([1 if k == 0 else 0 for k in kf_ids], np.uint8),
               points=np.array([mp_pos[m] for m in used], np.float64).reshape(len(used), 3),
               edge_pose=np.array(ep, np.int32), edge_point=np.array(em, np.int32),
               edge_obs=np.array(obs, np.float64).reshape(len(ep), 3), edge_info=np.array(info, np.float64),
               kf_ids=np.array(kf_ids, np.int64), mp_ids=np.array(used, np.int64))
    if humans:
        out.update(load_human_dump(path))
    return out


# include/Map.h:49-56: 14 bones over 14 of the 18 AlphaPose joints, 5 joints carry the motion edges
BODY1 = [1, 1, 8, 2, 5, 2, 3, 5, 6, 8, 9, 11, 12, 1]
BODY2 = [2, 5, 11, 8, 11, 3, 4, 6, 7, 9, 10, 12, 13, 0]
MAIN_SKELETON = [1, 2, 5, 11, 8]


def load_human_dump(path: str, sigma_rigidity: float = 20.0, sigma_motion: float = 20.0) -> dict:
    """HMTraj.txt + Motion.txt (src/Tracking.cc:1812-1830) -> the articulated arrays of a problem dict.  Per trajectory the poses
    keep file order; a pose is 14 consecutive HMTraj lines.  Rigidity / motion edges that touch a bad or lost joint are left out."""
    from . import ba
    tracks = {}      # track id -> list of (pose id, [(key id, bad, lost, xyz)] * 14) in file order
    order = []
    for line in open(os.path.join(path, "HMTraj.txt")):
        f = line.split()
        if len(f) < 8:
            continue
        tid, pid = int(f[0]), int(f[1])
        if tid not in tracks:
            tracks[tid] = []; order.append(tid)
        if not tracks[tid] or tracks[tid][-1][0] != pid:
            tracks[tid].append((pid, []))
        tracks[tid][-1][1].append((int(f[2]), int(f[3]), int(f[4]), [np.float32(v) for v in f[5:8]]))
    motions = {}
    for line in open(os.path.join(path, "Motion.txt")):
        f = line.split()
        if len(f) >= 17:
            motions[int(f[0])] = np.array(f[1:17], np.float32).reshape(4, 4)
    joints, key_ids, bad, lost, pose_of_joint, track_of_joint = [], [], [], [], [], []
    dists, r_i, r_j, r_d, m_p1, m_p2, m_m, mq, mt, track_ids = [], [], [], [], [], [], [], [], [], []
    for tr, tid in enumerate(order):
        poses = [p for p in tracks[tid] if len(p[1]) == 14]
        base = []
        for pid, keys in poses:
            base.append(len(joints))
            for (kid, b, l, xyz) in keys:
                joints.append(xyz); key_ids.append(kid); bad.append(b); lost.append(l); pose_of_joint.append(pid); track_of_joint.append(tid)
        J = np.array(joints, np.float64).reshape(-1, 3)
        ok = lambda j: not (bad[j] or lost[j])      # noqa: E731
        d0 = len(dists)
        for b in range(14):     # bone length = mean distance of its two joints over the poses that see both
            vals = [np.linalg.norm(J[j0 + BODY1[b]] - J[j0 + BODY2[b]]) for j0 in base if ok(j0 + BODY1[b]) and ok(j0 + BODY2[b])]
            dists.append(float(np.float32(np.mean(vals))) if vals else 0.0)
        for j0 in base:
            for b in range(14):
                if ok(j0 + BODY1[b]) and ok(j0 + BODY2[b]):
                    r_i.append(j0 + BODY1[b]); r_j.append(j0 + BODY2[b]); r_d.append(d0 + b)
        for a, c in zip(base[:-1], base[1:]):
            for j in MAIN_SKELETON:
                if ok(a + j) and ok(c + j):
                    m_p1.append(a + j); m_p2.append(c + j); m_m.append(tr)
        M = motions.get(tid, np.eye(4, dtype=np.float32))
        T = np.eye(4, dtype=np.float32); T[:3, :3] = M[:3, :3]; T[:3, 3] = M[:3, 3]
        q, t = ba.pose_from_tcw(T)
        mq.append(q); mt.append(t); track_ids.append(tid)
    n_t = len(order)
    return dict(joints=np.array(joints, np.float64).reshape(-1, 3), jedge_pose=np.zeros(0, np.int32), jedge_joint=np.zeros(0, np.int32),
                jedge_obs=np.zeros((0, 3)), jedge_info=np.zeros(0), dists=np.array(dists, np.float64),
                redge_i=np.array(r_i, np.int32), redge_j=np.array(r_j, np.int32), redge_dist=np.array(r_d, np.int32),
                redge_info=np.full(len(r_i), float(sigma_rigidity)), motion_q=np.array(mq, np.float64).reshape(n_t, 4),
                motion_t=np.array(mt, np.float64).reshape(n_t, 3), medge_p1=np.array(m_p1, np.int32), medge_p2=np.array(m_p2, np.int32),
                medge_motion=np.array(m_m, np.int32), medge_dt=np.ones(len(m_p1)), medge_info=np.full(len(m_p1), float(sigma_motion)),
                joint_key_ids=np.array(key_ids, np.int64), joint_bad=np.array(bad, np.uint8), joint_lost=np.array(lost, np.uint8),
                joint_pose_ids=np.array(pose_of_joint, np.int64), joint_track_ids=np.array(track_of_joint, np.int64), track_ids=np.array(track_ids, np.int64))


def save_map_dump(path: str, problem: dict, kf_ids=None, match_newlines: bool = False, human_poses: int = 4, joint_flags=None) -> None:
    """Writes KF.txt / MP.txt / Match.txt / HMTraj.txt / Motion.txt in the reference's format from a problem dict (the last two
    empty for a static window).  match_newlines=False reproduces src/Tracking.cc:1806-1807 (records run together).  Articulated
    part: joints are track-major, `human_poses` consecutive poses of 14 joints per trajectory (synth.make_ba_problem's layout);
    joint_flags = optional [n_joints][2] isBad / isLost."""
    from . import ba
    os.makedirs(path, exist_ok=True)
    n_p = len(problem["pose_q"])
    kf_ids = list(range(n_p)) if kf_ids is None else list(kf_ids)
    max_kf = max(kf_ids)
    with open(os.path.join(path, "KF.txt"), "w") as f:
        for i in range(n_p):
            Tcw = ba.pose_to_tcw(problem["pose_q"][i], problem["pose_t"][i]).astype(np.float64)
            Twc = np.eye(4); Twc[:3, :3] = Tcw[:3, :3].T; Twc[:3, 3] = -Tcw[:3, :3].T @ Tcw[:3, 3]
            f.write(str(kf_ids[i]) + " " + " ".join(_fmt(v) for v in Twc.astype(np.float32).ravel()) + "\n")
    with open(os.path.join(path, "MP.txt"), "w") as f:
        for j, X in enumerate(problem["points"]):
            f.write(f"{j + max_kf + 1} " + " ".join(_fmt(v) for v in X) + "\n")
    with open(os.path.join(path, "Match.txt"), "w") as f:
        for e in range(len(problem["edge_pose"])):
            o = problem["edge_obs"][e]
            f.write(f"{kf_ids[problem['edge_pose'][e]]} {problem['edge_point'][e] + max_kf + 1} {_fmt(o[0])} {_fmt(o[1])} {_fmt(o[2])} "
                    f"{_fmt(problem['edge_info'][e])}" + ("\n" if match_newlines else ""))
    n_j = len(problem.get("joints", ()))
    n_t = len(problem.get("motion_t", ()))
    with open(os.path.join(path, "HMTraj.txt"), "w") as f, open(os.path.join(path, "Motion.txt"), "w") as g:
        for tr in range(n_t):
            T = ba.pose_to_tcw(problem["motion_q"][tr], problem["motion_t"][tr])
            g.write(f"{tr} " + " ".join(_fmt(v) for v in T.ravel()) + "\n")
            for s_ in range(human_poses):
                j0 = (tr * human_poses + s_) * 14
                if j0 + 14 > n_j:
                    break
                for k in range(14):
                    fl = (0, 0) if joint_flags is None else (int(joint_flags[j0 + k][0]), int(joint_flags[j0 + k][1]))
                    f.write(f"{tr} {tr * human_poses + s_} {j0 + k} {fl[0]} {fl[1]} " + " ".join(_fmt(v) for v in problem["joints"][j0 + k]) + "\n")
