"""Seeded synthetic inputs shaped like the TartanAir-Shibuya stream AirDOS runs on.

numpy only (no cv2) so the same generator runs in the tests, in ``bench.py`` and on the GPU
box.  Shapes follow BASELINE.json's configs (640x480 u8 stereo pairs; BA windows of K key-frames /
P points / 6 observations per point); distributions follow SURVEY.md Appendix E.
"""
from __future__ import annotations

import numpy as np

# Examples/Stereo/config/tartanair.yaml:20-25 (reference camera); cy moved to 240 for 480 rows.
FX = 772.548
FY = 772.548
CX = 320.0
CY = 240.0
BF = 193.137


def _upsample4(a: np.ndarray) -> np.ndarray:
    """Bilinear x4 up-sampling of a 2-D float array (edge clamped)."""
    h, w = a.shape
    ys = (np.arange(h * 4) + 0.5) / 4 - 0.5
    xs = (np.arange(w * 4) + 0.5) / 4 - 0.5
    y0 = np.clip(np.floor(ys).astype(int), 0, h - 1)
    x0 = np.clip(np.floor(xs).astype(int), 0, w - 1)
    y1 = np.clip(y0 + 1, 0, h - 1)
    x1 = np.clip(x0 + 1, 0, w - 1)
    fy = np.clip(ys - y0, 0, 1)[:, None]
    fx = np.clip(xs - x0, 0, 1)[None, :]
    top = a[y0][:, x0] * (1 - fx) + a[y0][:, x1] * fx
    bot = a[y1][:, x0] * (1 - fx) + a[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


def _smooth(a: np.ndarray) -> np.ndarray:
    """Separable [1 2 1]/4 binomial smoothing, edge replicated."""
    p = np.pad(a, 1, mode="edge")
    a = (p[1:-1, :-2] + 2 * p[1:-1, 1:-1] + p[1:-1, 2:]) * 0.25
    p = np.pad(a, 1, mode="edge")
    return (p[:-2, 1:-1] + 2 * p[1:-1, 1:-1] + p[2:, 1:-1]) * 0.25


def make_image(seed: int, width: int = 640, height: int = 480, n_shapes: int = 200) -> np.ndarray:
    """One textured u8 image with plenty of corners (float64 scene, before camera noise)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(40, 216, size=((height + 3) // 4, (width + 3) // 4)).astype(np.float64)
    img = _upsample4(base)[:height, :width]
    for _ in range(n_shapes):
        w = int(rng.integers(6, 60))
        h = int(rng.integers(6, 60))
        x = int(rng.integers(0, width - 6))
        y = int(rng.integers(0, height - 6))
        img[y:y + h, x:x + w] = float(rng.integers(0, 256))
    return _smooth(img)


def _finish(scene: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    noisy = scene + rng.normal(0.0, 2.0, size=scene.shape)
    return np.clip(np.rint(noisy), 0, 255).astype(np.uint8)


def make_stereo_pair(frame: int, width: int = 640, height: int = 480):
    """(left, right) u8 images.  The right image is the left scene displaced by d = bf / Z with a
    piece-wise constant depth Z in [3, 40] m per 40-row band (linear interpolation in x)."""
    scene = make_image(1000 + frame, width, height)
    rng = np.random.default_rng(2000 + frame)
    bands = (height + 39) // 40
    depth = rng.uniform(3.0, 40.0, size=bands)
    disp = np.repeat(BF / depth, 40)[:height]
    xs = np.arange(width)[None, :] + disp[:, None]          # right(x) = left(x + d)
    x0 = np.floor(xs).astype(int)
    fx = xs - x0
    x0c = np.clip(x0, 0, width - 1)
    x1c = np.clip(x0 + 1, 0, width - 1)
    rows = np.arange(height)[:, None]
    right_scene = scene[rows, x0c] * (1 - fx) + scene[rows, x1c] * fx
    return _finish(scene, rng), _finish(right_scene, rng)


def make_stereo_batch(n_pairs: int, width: int = 640, height: int = 480, start: int = 0) -> np.ndarray:
    """u8 array [n_pairs, 2, height, width] (index 0 = left, 1 = right)."""
    out = np.empty((n_pairs, 2, height, width), np.uint8)
    for f in range(n_pairs):
        out[f, 0], out[f, 1] = make_stereo_pair(start + f, width, height)
    return out


def make_mask(seed: int, width: int = 640, height: int = 480, n_rect: int = 3) -> np.ndarray:
    """Extractor mask as Frame::ExtractORB builds it (src/Frame.cc:553-560): 255 = keep, 0 = human."""
    rng = np.random.default_rng(seed)
    m = np.full((height, width), 255, np.uint8)
    for _ in range(n_rect):
        w = int(rng.integers(40, 200))
        h = int(rng.integers(60, 300))
        x = int(rng.integers(0, width - 40))
        y = int(rng.integers(0, height - 60))
        m[y:y + h, x:x + w] = 0
    return m


def make_human_mask(seed: int, width: int = 640, height: int = 480, n_people: int = 2) -> np.ndarray:
    """Segmentation-shaped extractor mask as AirDOS feeds it (src/Frame.cc:551-571, System.IsMask: 1 in the shipped
    Examples/Stereo/config/tartanair.yaml): 255 = static scene, 0 = pixels of a detected person.  A person is a head disc, a
    torso ellipse and two leg strips, 80-220 px tall, anywhere in the lower two thirds of the image (TartanAir-Shibuya street
    scenes); 2-10 % of the image ends up masked."""
    rng = np.random.default_rng(seed)
    m = np.full((height, width), 255, np.uint8)
    yy, xx = np.mgrid[0:height, 0:width]
    for _ in range(n_people):
        hgt = float(rng.uniform(80, 220)); cx = float(rng.uniform(40, width - 40)); top = float(rng.uniform(height * 0.15, height - hgt * 0.6))
        head_r = 0.08 * hgt
        m[(xx - cx) ** 2 + (yy - (top + head_r)) ** 2 <= head_r ** 2] = 0
        m[((xx - cx) / (0.17 * hgt)) ** 2 + ((yy - (top + 0.38 * hgt)) / (0.25 * hgt)) ** 2 <= 1.0] = 0
        for side in (-1, 1):
            lx = cx + side * 0.08 * hgt
            m[(np.abs(xx - lx) <= 0.055 * hgt) & (yy >= top + 0.55 * hgt) & (yy <= top + hgt)] = 0
    return m


# ------------------------------------------------------------------------------------------
# Bundle-adjustment windows (BASELINE.json configs[3] / configs[4])
ORB_QUOTA_2000 = np.array([434, 362, 302, 251, 209, 175, 145, 122], np.float64)


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _small_rot(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def tcw_to_pose(T):
    """Converter::toSE3Quat on a float32 4x4 (src/Converter.cc:37-47): Eigen::Quaterniond(R) then
    normalise with w >= 0.  Returns (q[x,y,z,w], t) in float64."""
    T = np.asarray(T, np.float32).astype(np.float64)
    m = T[:3, :3]
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if tr > 0:
        s = np.sqrt(tr + 1.0)
        q[3] = 0.5 * s
        s = 0.5 / s
        q[0], q[1], q[2] = (m[2, 1] - m[1, 2]) * s, (m[0, 2] - m[2, 0]) * s, (m[1, 0] - m[0, 1]) * s
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[i] = 0.5 * s
        s = 0.5 / s
        q[3] = (m[k, j] - m[j, k]) * s
        q[j] = (m[j, i] + m[i, j]) * s
        q[k] = (m[k, i] + m[i, k]) * s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q), T[:3, 3].copy()


def make_ba_problem(n_kf: int = 50, n_points: int = 20000, obs_per_point: int = 6, seed: int = 4000,
                    outlier_frac: float = 0.03, mono_frac: float = 0.0, n_fixed_extra: int = 0,
                    noise: bool = True, humans: int = 0, human_poses: int = 4):
    """Synthetic local-BA window.  Returns a dict of numpy arrays laid out like adb_ba_problem plus
    the ground truth ('gt_pose_t', 'gt_points').  KF 0 is fixed (+ n_fixed_extra trailing fixed observers)."""
    rng = np.random.default_rng(seed)
    K = n_kf + n_fixed_extra
    yaw = np.cumsum(rng.normal(0, np.deg2rad(2.0), K)); yaw[0] = 0
    centers = np.zeros((K, 3)); Rwc = np.zeros((K, 3, 3))
    for k in range(K):
        Rwc[k] = _rot_y(yaw[k])
        if k:
            centers[k] = centers[k - 1] + 0.5 * np.array([np.sin(yaw[k]), 0, np.cos(yaw[k])])
    Rcw = np.transpose(Rwc, (0, 2, 1))
    tcw = -np.einsum("kij,kj->ki", Rcw, centers)
    W, H = 640, 480
    quota = ORB_QUOTA_2000 / ORB_QUOTA_2000.sum()
    pts, e_pose, e_point, e_obs, e_info = [], [], [], [], []
    n_done = 0
    while n_done < n_points:
        m = max(1024, (n_points - n_done) * 2)
        anchor = rng.integers(0, K, m)
        u = rng.uniform(20, W - 20, m); v = rng.uniform(20, H - 20, m); z = rng.uniform(3.0, 40.0, m)
        Xc = np.stack([(u - CX) / FX * z, (v - CY) / FY * z, z], 1)
        Xw = np.einsum("mij,mj->mi", Rwc[anchor], Xc) + centers[anchor]
        for i in range(m):
            lo, hi = max(0, anchor[i] - 8), min(K - 1, anchor[i] + 8)
            win = np.arange(lo, hi + 1)
            pc = np.einsum("kij,j->ki", Rcw[win], Xw[i]) + tcw[win]
            zz = pc[:, 2]
            ok = zz > 1.0
            uu = np.where(ok, pc[:, 0] / np.where(ok, zz, 1) * FX + CX, -1)
            vv = np.where(ok, pc[:, 1] / np.where(ok, zz, 1) * FY + CY, -1)
            ok &= (uu > 5) & (uu < W - 5) & (vv > 5) & (vv < H - 5)
            vis = win[ok]
            if len(vis) < obs_per_point:
                continue
            sel = np.sort(rng.choice(vis, obs_per_point, replace=False))
            for k in sel:
                p = Rcw[k] @ Xw[i] + tcw[k]
                lvl = rng.choice(8, p=quota)
                sig = 1.2 ** lvl if noise else 0.0
                uo, vo = p[0] / p[2] * FX + CX, p[1] / p[2] * FY + CY
                uro = uo - BF / p[2]
                n3 = rng.normal(0, 1, 3) * sig
                if noise and rng.random() < outlier_frac:
                    n3 = n3 + rng.uniform(-30, 30, 3)
                mono = rng.random() < mono_frac
                e_pose.append(k); e_point.append(n_done)
                uro_n = uro + n3[2]
                # a negative right-image coordinate means 'not seen by the right camera': monocular edge (mvuRight < 0)
                e_obs.append([uo + n3[0], vo + n3[1], -1.0 if (mono or uro_n < 0) else uro_n])
                e_info.append(np.float32(1.0) / (np.float32(1.2 ** lvl) * np.float32(1.2 ** lvl)))
            pts.append(Xw[i]); n_done += 1
            if n_done >= n_points:
                break
    gt_points = np.array(pts)
    # initial estimates: perturbed, then through float32 (the map stores cv::Mat float)
    pose_q = np.zeros((K, 4)); pose_t = np.zeros((K, 3))
    for k in range(K):
        R, t = Rcw[k], tcw[k]
        if k > 0 and noise:
            R = _small_rot(rng.normal(0, np.deg2rad(0.2), 3)) @ R
            t = t + rng.normal(0, 0.01, 3)
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
        pose_q[k], pose_t[k] = tcw_to_pose(T.astype(np.float32))
    points = (gt_points + (rng.normal(0, 0.05, gt_points.shape) if noise else 0)).astype(np.float32).astype(np.float64)
    fixed = np.zeros(K, np.uint8); fixed[0] = 1; fixed[n_kf:] = 1
    prob = dict(fx=FX, fy=FY, cx=CX, cy=CY, bf=BF, pose_q=pose_q, pose_t=pose_t, pose_fixed=fixed, points=points,
                edge_pose=np.array(e_pose, np.int32), edge_point=np.array(e_point, np.int32),
                edge_obs=np.array(e_obs, np.float32).astype(np.float64), edge_info=np.array(e_info, np.float32).astype(np.float64),
                gt_pose_t=tcw.copy(), gt_pose_R=Rcw.copy(), gt_points=gt_points)
    if humans:
        _add_humans(prob, rng, humans, human_poses, Rcw, tcw, centers, Rwc, noise)
    return prob


# Map.h:49-56 topology of the reference: 14 bones over 14 of the 18 AlphaPose joints, 5 motion joints
BODY1 = [1, 1, 8, 2, 5, 2, 3, 5, 6, 8, 9, 11, 12, 1]
BODY2 = [2, 5, 11, 8, 11, 3, 4, 6, 7, 9, 10, 12, 13, 0]
MAIN_SKELETON = [1, 2, 5, 11, 8]
_CANON = np.array([[0, -1.60, 0], [0, -1.45, 0], [-0.18, -1.42, 0], [-0.25, -1.15, 0], [-0.27, -0.90, 0], [0.18, -1.42, 0],
                   [0.25, -1.15, 0], [0.27, -0.90, 0], [-0.10, -0.92, 0], [-0.11, -0.50, 0], [-0.11, -0.08, 0],
                   [0.10, -0.92, 0], [0.11, -0.50, 0], [0.11, -0.08, 0]])   # 1.7 m skeleton, y down


def _add_humans(prob, rng, n_tracks, n_poses, Rcw, tcw, centers, Rwc, noise):
    """n_tracks trajectories x n_poses consecutive skeletons (14 joints), walking 1.2 m/s, dt = 1."""
    K = len(tcw)
    joints, jedge_pose, jedge_joint, jedge_obs, jedge_info = [], [], [], [], []
    dists, r_i, r_j, r_d = [], [], [], []
    m_p1, m_p2, m_m, m_dt = [], [], [], []
    for tr in range(n_tracks):
        k0 = int(rng.integers(2, max(3, K - n_poses - 2)))
        heading = rng.uniform(0, 2 * np.pi)
        vel = 1.2 * np.array([np.cos(heading), 0, np.sin(heading)])
        base = centers[k0] + Rwc[k0] @ np.array([rng.uniform(-1.5, 1.5), 1.0, rng.uniform(6, 12)])
        bone_len = [np.linalg.norm(_CANON[a] - _CANON[b]) for a, b in zip(BODY1, BODY2)]
        d0 = len(dists)
        dists += [l + (rng.normal(0, 0.02) if noise else 0) for l in bone_len]
        prev = None
        for s in range(n_poses):
            kf = min(k0 + s, K - 1)
            sk = _CANON + base + vel * s
            j0 = len(joints)
            for j in range(14):
                p = Rcw[kf] @ sk[j] + tcw[kf]
                uo, vo = p[0] / p[2] * FX + CX, p[1] / p[2] * FY + CY
                n3 = rng.normal(0, 2.0, 3) if noise else np.zeros(3)
                joints.append(sk[j] + (rng.normal(0, 0.05, 3) if noise else 0))
                jedge_pose.append(kf); jedge_joint.append(j0 + j)
                jedge_obs.append([uo + n3[0], vo + n3[1], max(uo - BF / p[2] + n3[2], 0.0)])   # joints sit 6-12 m away: always positive
                jedge_info.append(0.5)
            for b in range(14):
                r_i.append(j0 + BODY1[b]); r_j.append(j0 + BODY2[b]); r_d.append(d0 + b)
            if prev is not None:
                for j in MAIN_SKELETON:
                    m_p1.append(prev + j); m_p2.append(j0 + j); m_m.append(tr); m_dt.append(1.0)
            prev = j0
    nm = n_tracks
    prob.update(joints=np.array(joints, np.float32).astype(np.float64), jedge_pose=np.array(jedge_pose, np.int32),
                jedge_joint=np.array(jedge_joint, np.int32), jedge_obs=np.array(jedge_obs, np.float32).astype(np.float64),
                jedge_info=np.array(jedge_info, np.float64), dists=np.array(dists, np.float32).astype(np.float64),
                redge_i=np.array(r_i, np.int32), redge_j=np.array(r_j, np.int32), redge_dist=np.array(r_d, np.int32),
                redge_info=np.full(len(r_i), 20.0), motion_q=np.tile(np.array([0, 0, 0, 1.0]), (nm, 1)), motion_t=np.zeros((nm, 3)),
                medge_p1=np.array(m_p1, np.int32), medge_p2=np.array(m_p2, np.int32), medge_motion=np.array(m_m, np.int32),
                medge_dt=np.array(m_dt, np.float64), medge_info=np.full(len(m_p1), 20.0))


def make_pose_frames(n_frames: int = 4, n_points: int = 600, seed: int = 7, outlier_frac: float = 0.1, mono_frac: float = 0.2):
    """Frames for Optimizer::PoseOptimization: each sees its own MapPoints (float32 world positions), stereo / mono
    observations with octave noise and gross outliers, and starts from a perturbed pose.  Returns (cam, frames, gt_t)."""
    rng = np.random.default_rng(seed)
    cam = dict(fx=FX, fy=FY, cx=CX, cy=CY, bf=BF)
    frames, gts = [], []
    for f in range(n_frames):
        n = n_points if f != 1 else 7            # one frame below the 10-edge early exit
        R = _small_rot(rng.normal(0, 0.1, 3)); t = rng.normal(0, 0.5, 3)
        u = rng.uniform(20, 620, n); v = rng.uniform(20, 460, n); z = rng.uniform(3, 40, n)
        Xc = np.stack([(u - CX) / FX * z, (v - CY) / FY * z, z], 1)
        Xw = ((Xc - t) @ R).astype(np.float32)                     # Xc = R Xw + t
        Xc = Xw.astype(np.float64) @ R.T + t
        lvl = rng.choice(8, n, p=ORB_QUOTA_2000 / ORB_QUOTA_2000.sum())
        sig = 1.2 ** lvl
        uo = Xc[:, 0] / Xc[:, 2] * FX + CX + rng.normal(0, 1, n) * sig
        vo = Xc[:, 1] / Xc[:, 2] * FY + CY + rng.normal(0, 1, n) * sig
        ur = uo - BF / Xc[:, 2] + rng.normal(0, 1, n) * sig
        bad = rng.random(n) < outlier_frac
        uo[bad] += rng.uniform(-40, 40, bad.sum()); vo[bad] += rng.uniform(-40, 40, bad.sum())
        ur[(rng.random(n) < mono_frac) | (ur < 0)] = -1.0
        T = np.eye(4); T[:3, :3] = _small_rot(rng.normal(0, 0.01, 3)) @ R; T[:3, 3] = t + rng.normal(0, 0.05, 3)
        q, tt = tcw_to_pose(T.astype(np.float32))
        frames.append(dict(pose_q=q, pose_t=tt, xw=Xw, obs=np.stack([uo, vo, ur], 1).astype(np.float32),
                           inv_sigma2=(np.float32(1) / (np.float32(1.2) ** lvl.astype(np.float32)) ** 2).astype(np.float32)))
        gts.append(t)
    return cam, frames, np.array(gts)


# ------------------------------------------------------------------------------------------
# Tracking search problems: a current Frame and the map points of the last frame
def make_tracking_problem(seed: int = 0, n_kp: int = 2000, n_q: int = 1500, width: int = 640, height: int = 480,
                          th: float = 7.0, dup_frac: float = 0.1, n_levels: int = 8, scale: float = 1.2, last_dz: float = 0.02):
    """Problem dict for ORBmatcher.search_by_projection, last-frame variant (src/ORBmatcher.cc:1328-1470): key-points of the
    current frame (positions uniform in the image, octaves by the ORB quota, ~70 % with a stereo match), and n_q map
    points of the last frame: most are a current key-point back-projected through its stereo depth (descriptor with a few
    flipped bits, octave +-1), `dup_frac` of them duplicates that compete for the same key-point (the reference's
    sequential closure rule decides), the rest random.  The last pose is the current one moved by a few centimetres."""
    from .capi import KP_DTYPE
    rng = np.random.default_rng(seed)
    sf = (np.float32(scale) ** np.arange(n_levels)).astype(np.float32)
    for i in range(1, n_levels):   # mvScaleFactor[i] = mvScaleFactor[i-1] * scaleFactor (src/ORBextractor.cc:418-424)
        sf[i] = np.float32(sf[i - 1] * np.float32(scale))
    kps = np.zeros(n_kp, KP_DTYPE)
    kps["x"] = rng.uniform(20, width - 20, n_kp).astype(np.float32)
    kps["y"] = rng.uniform(20, height - 20, n_kp).astype(np.float32)
    quota = ORB_QUOTA_2000[:n_levels] / ORB_QUOTA_2000[:n_levels].sum()
    kps["octave"] = rng.choice(n_levels, n_kp, p=quota)
    kps["angle"] = rng.uniform(0, 360, n_kp).astype(np.float32)
    kps["size"] = 31 * sf[kps["octave"]]
    desc = rng.integers(0, 256, (n_kp, 32), dtype=np.uint8)
    depth = rng.uniform(3.0, 40.0, n_kp)
    has_stereo = rng.random(n_kp) < 0.7
    u_right = np.where(has_stereo, kps["x"] - BF / depth, -1.0).astype(np.float32)
    taken = (rng.random(n_kp) < 0.03).astype(np.uint8)
    # poses: current = small motion from identity; last = current moved a little further back
    Tc = np.eye(4); Tc[:3, :3] = _small_rot(rng.normal(0, 0.01, 3)); Tc[:3, 3] = rng.normal(0, 0.05, 3)
    Tl = np.eye(4); Tl[:3, :3] = _small_rot(rng.normal(0, 0.01, 3)) @ Tc[:3, :3]; Tl[:3, 3] = Tc[:3, 3] + np.array([0.0, 0.0, last_dz])   # |dz| > mb switches the forward / backward level rule
    Rcw, tcw = Tc[:3, :3], Tc[:3, 3]
    src = rng.integers(0, n_kp, n_q)
    n_dup = int(dup_frac * n_q)
    src[n_q - n_dup:] = src[rng.integers(0, n_q - n_dup, n_dup)]          # duplicates of earlier queries
    perm = rng.permutation(n_q); src = src[perm]
    is_random = rng.random(n_q) < 0.15
    # back-project the source key-point (with a little pixel noise) into the world through the current pose
    u = kps["x"][src] + rng.normal(0, 1.5, n_q); v = kps["y"][src] + rng.normal(0, 1.5, n_q)
    z = depth[src]
    xc = np.stack([(u - CX) * z / FX, (v - CY) * z / FY, z], 1)
    xw = (xc - tcw) @ Rcw          # Rcw^T (xc - tcw)
    xw[is_random] = rng.uniform([-10, -5, 2], [10, 5, 40], (int(is_random.sum()), 3))
    xw[rng.random(n_q) < 0.02, 2] = -5.0   # behind the camera
    q_desc = desc[src].copy()
    flips = rng.integers(0, 60, n_q)
    for i in range(n_q):
        bits = rng.choice(256, flips[i], replace=False)
        np.bitwise_xor.at(q_desc[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    q_desc[is_random] = rng.integers(0, 256, (int(is_random.sum()), 32), dtype=np.uint8)
    octave = np.clip(kps["octave"][src] + rng.integers(-1, 2, n_q), 0, n_levels - 1).astype(np.int32)
    rot = np.where(rng.random(n_q) < 0.8, rng.normal(4.0, 3.0, n_q), rng.uniform(0, 360, n_q))
    q_angle = np.mod(kps["angle"][src] + rot, 360.0).astype(np.float32)
    q_flags = ((rng.random(n_q) < 0.93).astype(np.uint8) | ((rng.random(n_q) < 0.5).astype(np.uint8) << 1)).astype(np.uint8)
    return dict(kps=kps, u_right=u_right, desc=desc, taken=taken, bounds=(0.0, 0.0, float(width), float(height)),
                q_flags=q_flags, q_desc=q_desc, q_angle=q_angle, last_xw=xw.astype(np.float32), last_octave=octave,
                tcw_cur=Tc.astype(np.float32), tcw_last=Tl.astype(np.float32), cam=(FX, FY, CX, CY, BF, BF / FX),
                scale_factors=sf, th=th, mono=0)


def tracking_problem_as_map_points(pr: dict, proj: dict, nn_ratio: float = 0.8) -> dict:
    """The same scene as a SearchByProjection(Frame, vpMapPoints, th) problem (src/ORBmatcher.cc:45-129): `proj` are the
    projected query arrays (from the oracle's projection); level window = (nPredictedLevel - 1, nPredictedLevel)."""
    out = {k: pr[k] for k in ("kps", "u_right", "desc", "taken", "bounds", "q_desc", "q_angle")}
    lvl = np.asarray(pr["last_octave"], np.int32)
    out.update(q_u=proj["q_u"], q_v=proj["q_v"], q_ur=proj["q_ur"], q_radius=proj["q_radius"], q_min_level=(lvl - 1).astype(np.int32),
               q_max_level=lvl.copy(), q_flags=proj["q_flags"], use_ratio=1, nn_ratio=nn_ratio, check_orientation=0)
    return out


def tracking_problem_as_local_map(pr: dict, seed: int = 0, nn_ratio: float = 0.8, th: float = 1.0) -> dict:
    """The same scene as Tracking::SearchLocalPoints sees it (src/Tracking.cc:1319-1343): local map points BEFORE the
    visibility test -- world position, mean viewing direction, scale-invariance distances -- so Frame::isInFrustum,
    MapPoint::PredictScale and SearchByProjection(F, vpMapPoints, th) all run behind the C-ABI.  A fifth of the points
    fails one of the tests (wrong side, out of the distance range, oblique view)."""
    rng = np.random.default_rng(seed + 77)
    out = {k: pr[k] for k in ("kps", "u_right", "desc", "taken", "bounds", "q_desc", "tcw_cur", "cam", "scale_factors")}
    T = pr["tcw_cur"].astype(np.float64)
    ow = -T[:3, :3].T @ T[:3, 3]
    xw = pr["last_xw"].astype(np.float64)
    po = xw - ow
    dist = np.linalg.norm(po, axis=1)
    normal = po / dist[:, None] + rng.normal(0, 0.15, po.shape)
    oblique = rng.random(len(xw)) < 0.07
    normal[oblique] = rng.normal(0, 1, (int(oblique.sum()), 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    lvl = np.asarray(pr["last_octave"], np.float64)
    max_d = dist * 1.2 ** lvl * rng.uniform(0.9, 1.1, len(xw))           # PredictScale then lands near the source octave
    min_d = max_d / 1.2 ** 7
    far = rng.random(len(xw)) < 0.07
    max_d[far] = dist[far] * 0.5
    out.update(mp_xw=xw.astype(np.float32), mp_normal=normal.astype(np.float32), mp_min_distance=min_d.astype(np.float32),
               mp_max_distance=max_d.astype(np.float32), ow=ow.astype(np.float32), q_flags=pr["q_flags"].copy(),
               log_scale_factor=float(np.log(np.float32(1.2))), th=th, nn_ratio=nn_ratio, use_ratio=1, view_cos_limit=0.5)
    return out


def tracking_problem_as_fuse(pr: dict, seed: int = 0, th: float = 3.0) -> dict:
    """The same scene as ORBmatcher::Fuse(pKF, vpMapPoints, th) sees it (src/ORBmatcher.cc:825-975; LocalMapping::SearchInNeighbors
    calls it with the default th = 3.0): the key-frame's key-points and the neighbours' map points."""
    out = tracking_problem_as_local_map(pr, seed=seed, th=th)
    sf = np.asarray(pr["scale_factors"], np.float32)
    out.update(fuse=1, inv_level_sigma2=(np.float32(1.0) / (sf * sf)).astype(np.float32), taken=None)
    return out


def make_bow_problem(seed: int = 0, mode: int = 0, n1: int = 1800, n2: int = 2000, n_nodes: int = 400, n_levels: int = 8, scale: float = 1.2):
    """Problem dict for ORBmatcher.search_by_bow.  Side 2 = a frame / key-frame with random key-points; side 1 = n1 features that
    mostly re-observe a side-2 feature (descriptor with a few flipped bits, same vocabulary node, small rotation), shifted
    along x like a second key-frame translated sideways (F12 of a pure x translation: horizontal epipolar lines, epipole at
    infinity).  Vocabulary nodes are synthetic (DBoW2 is out of scope): node = source feature mod n_nodes, 10 % of the
    side-1 features land in a random node; duplicates compete for the same partner (the first-come rule of SearchByBoW)."""
    from .capi import KP_DTYPE
    rng = np.random.default_rng(seed + 9000)
    sf = np.ones(n_levels, np.float32)
    for i in range(1, n_levels):
        sf[i] = np.float32(sf[i - 1] * np.float32(scale))
    quota = ORB_QUOTA_2000[:n_levels] / ORB_QUOTA_2000[:n_levels].sum()
    k2 = np.zeros(n2, KP_DTYPE)
    k2["x"] = rng.uniform(20, 620, n2).astype(np.float32); k2["y"] = rng.uniform(20, 460, n2).astype(np.float32)
    k2["octave"] = rng.choice(n_levels, n2, p=quota); k2["angle"] = rng.uniform(0, 360, n2).astype(np.float32)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    ur2 = np.where(rng.random(n2) < 0.7, k2["x"] - rng.uniform(5, 60, n2), -1.0).astype(np.float32)
    node2 = np.arange(n2) % n_nodes
    src = rng.integers(0, n2, n1)
    src[n1 - n1 // 5:] = src[rng.integers(0, n1 - n1 // 5, n1 // 5)]      # duplicates
    src = src[rng.permutation(n1)]
    k1 = np.zeros(n1, KP_DTYPE)
    k1["x"] = (k2["x"][src] + rng.uniform(5, 40, n1)).astype(np.float32)
    k1["y"] = (k2["y"][src] + rng.normal(0, 1.0, n1) * sf[k2["octave"][src]]).astype(np.float32)
    k1["octave"] = k2["octave"][src]
    rot = np.where(rng.random(n1) < 0.8, rng.normal(3.0, 3.0, n1), rng.uniform(0, 360, n1))
    k1["angle"] = np.mod(k2["angle"][src] + rot, 360.0).astype(np.float32)
    d1 = d2[src].copy()
    flips = rng.integers(0, 70, n1)
    for i in range(n1):
        bits = rng.choice(256, flips[i], replace=False)
        np.bitwise_xor.at(d1[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    ur1 = np.where(rng.random(n1) < 0.7, k1["x"] - rng.uniform(5, 60, n1), -1.0).astype(np.float32)
    node1 = node2[src].copy()
    stray = rng.random(n1) < 0.1
    node1[stray] = rng.integers(0, n_nodes + 40, int(stray.sum()))       # some nodes exist on one side only
    common = sorted(set(node1.tolist()) & set(node2.tolist()))
    p1, i1, p2, i2 = [0], [], [0], []
    for nd in common:                                                    # the FeatureVector walk keeps ascending node ids
        a = np.nonzero(node1 == nd)[0]; b = np.nonzero(node2 == nd)[0]
        i1 += a.tolist(); i2 += b.tolist(); p1.append(len(i1)); p2.append(len(i2))
    pr = dict(mode=mode, kps1=k1, desc1=d1, kps2=k2, desc2=d2, b_ptr1=np.array(p1, np.int32), b_idx1=np.array(i1, np.int32),
              b_ptr2=np.array(p2, np.int32), b_idx2=np.array(i2, np.int32), nn_ratio=0.7, check_orientation=1)
    if mode == 0:
        pr["flags1"] = (rng.random(n1) < 0.85).astype(np.uint8)         # key-points of the key-frame that hold a good map point
    else:
        pr["flags1"] = (rng.random(n1) < 0.6).astype(np.uint8)          # not triangulated yet
        pr.update(flags2=(rng.random(n2) < 0.6).astype(np.uint8), u_right1=ur1, u_right2=ur2,
                  f12=np.array([0, 0, 0, 0, 0, -1, 0, 1, 0], np.float32), epipole=(1.0e6, 240.0), scale_factors2=sf,
                  level_sigma2_2=(sf * sf).astype(np.float32))
    return pr
