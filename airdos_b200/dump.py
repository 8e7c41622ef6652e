"""The reference's own map dump as a bundle-adjustment problem.

Tracking::SaveMap (src/Tracking.cc:1745-1838) writes, with default ostream formatting (6 significant digits):
    KF.txt      <KF id> <16 floats: Twc = KeyFrame::GetPoseInverse(), row major>\\n
    MP.txt      <MP id + maxKFid + 1> <x> <y> <z>\\n            (bad points included)
    Match.txt   <KF id> <MP id + maxKFid + 1> <u> <v> <u_right> <invSigma2>   -- and NO newline: line 1807 reads
                `<< invSigma2; '\\n';`, so the record separator is missing and invSigma2 runs into the next KF id
                ("0.69444412" = invSigma2 0.694444 followed by KF 12)
    HMTraj.txt  <track id> <pose id> <key id> <isBad> <isLost> <x> <y> <z>\\n   (14 joints per pose)
    Motion.txt  <track id> <16 floats: mTMotion>\\n
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
    """Tokens of Match.txt with the run-together `<invSigma2><next KF id>` token split."""
    out = []
    raw = text.split()
    i = 0
    while i < len(raw):
        rec = raw[i:i + 6]
        if len(rec) < 6:
            raise ValueError("Match.txt: truncated record")
        tok = rec[5]
        # longest printed level value that prefixes the token
        best = max((p for p in printed if tok.startswith(p)), key=len, default=None)
        if best is None:
            raise ValueError(f"Match.txt: {tok!r} does not start with a level invSigma2 ({printed})")
        out.append(rec[:5] + [best])
        rest = tok[len(best):]
        i += 6
        if rest:                       # the bug: the next record's KF id is glued on
            raw.insert(i, rest)
    return out


def load_map_dump(path: str, cam: dict, scale_factor: float = 1.2, n_levels: int = 8) -> dict:
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
    used = sorted({int(r[1]) for r in recs if int(r[0]) in kf_index})
    mp_pos = dict(zip(mp_ids, pts))
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
    return dict(fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], bf=cam["bf"],
                pose_q=np.array(poses_q, np.float64).reshape(n_p, 4), pose_t=np.array(poses_t, np.float64).reshape(n_p, 3),
                pose_fixed=np.array([1 if k == 0 else 0 for k in kf_ids], np.uint8),
                points=np.array([mp_pos[m] for m in used], np.float64).reshape(len(used), 3),
                edge_pose=np.array(ep, np.int32), edge_point=np.array(em, np.int32),
                edge_obs=np.array(obs, np.float64).reshape(len(ep), 3), edge_info=np.array(info, np.float64),
                kf_ids=np.array(kf_ids, np.int64), mp_ids=np.array(used, np.int64))


def save_map_dump(path: str, problem: dict, kf_ids=None, match_newlines: bool = False) -> None:
    """Writes KF.txt / MP.txt / Match.txt (and empty HMTraj.txt / Motion.txt) in the reference's format from a static
    problem dict.  match_newlines=False reproduces src/Tracking.cc:1806-1807 (records run together)."""
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
    for name in ("HMTraj.txt", "Motion.txt"):
        open(os.path.join(path, name), "w").close()
