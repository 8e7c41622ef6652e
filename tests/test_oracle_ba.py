"""CPU tests pinning the BA oracle (oracle/ba_oracle.cpp).  The reference ships no BA tests and cannot
be built here, so the restatement is pinned by independent mathematics:
  * SE3 exp / oplus against scipy's matrix exponential,
  * the first LM step (Jacobians, Huber weights, Schur complement, back-substitution) against a
    dense numpy normal-equation solve whose Jacobian comes from central differences of an
    independently written residual,
  * convergence to the exact optimum on noise-free data, LM bookkeeping invariants."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import expm

from airdos_b200 import synth


def _q2R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _se3_exp(d):
    w, v = d[:3], d[3:]
    A = np.zeros((4, 4))
    A[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    A[:3, 3] = v
    return expm(A)


def test_pose_oplus_is_left_multiplication_by_exp(oracle_mod):
    lib = oracle_mod.ba_lib()
    rng = np.random.default_rng(0)
    for scale in (1e-7, 1e-3, 0.3):
        q = rng.normal(size=4); q /= np.linalg.norm(q); q = q if q[3] > 0 else -q
        t = rng.normal(size=3)
        d = rng.normal(size=6) * scale
        T = np.eye(4); T[:3, :3] = _q2R(q); T[:3, 3] = t
        ref = _se3_exp(d) @ T
        q2, t2 = q.copy(), t.copy()
        lib.ba_oracle_pose_oplus(q2.ctypes.data_as(C.c_void_p), t2.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p))
        assert np.allclose(_q2R(q2), ref[:3, :3], atol=1e-9 if scale < 0.1 else 1e-9)
        assert np.allclose(t2, ref[:3, 3], atol=1e-9)
        assert abs(np.linalg.norm(q2) - 1) < 1e-14 and q2[3] >= 0


def _residuals(d, pose_q, pose_t, points):
    """Independent residual: plain double pinhole stereo model (no float invz)."""
    R = np.array([_q2R(q) for q in pose_q])
    Xc = np.einsum("eij,ej->ei", R[d["edge_pose"]], points[d["edge_point"]]) + pose_t[d["edge_pose"]]
    u = Xc[:, 0] / Xc[:, 2] * d["fx"] + d["cx"]
    v = Xc[:, 1] / Xc[:, 2] * d["fy"] + d["cy"]
    ur = u - d["bf"] / Xc[:, 2]
    e = d["edge_obs"] - np.stack([u, v, ur], 1)
    mono = d["edge_obs"][:, 2] < 0
    e[mono, 2] = 0
    return e


@pytest.mark.parametrize("mono_frac,robust", [(0.0, 1), (0.3, 1), (0.0, 0)])
def test_first_lm_step_matches_dense_numpy_normal_equations(oracle_mod, mono_frac, robust):
    from airdos_b200 import ba_types as T
    d = synth.make_ba_problem(5, 60, 4, seed=11, mono_frac=mono_frac)
    K, P, E = len(d["pose_t"]), len(d["points"]), len(d["edge_pose"])
    free = [k for k in range(K) if not d["pose_fixed"][k]]
    off = {k: 6 * i for i, k in enumerate(free)}
    nd = 6 * len(free)
    n = nd + 3 * P

    def state(x):
        pq, pt, X = d["pose_q"].copy(), d["pose_t"].copy(), d["points"].copy()
        for k in free:
            Tm = np.eye(4); Tm[:3, :3] = _q2R(pq[k]); Tm[:3, 3] = pt[k]
            Tn = _se3_exp(x[off[k]:off[k] + 6]) @ Tm
            pq[k], pt[k] = synth.tcw_to_pose(Tn)[0], Tn[:3, 3]
            # tcw_to_pose casts through float32: redo in float64
            m = Tn[:3, :3]
            w = np.sqrt(max(0, 1 + m[0, 0] + m[1, 1] + m[2, 2])) / 2
            pq[k] = np.array([(m[2, 1] - m[1, 2]) / (4 * w), (m[0, 2] - m[2, 0]) / (4 * w), (m[1, 0] - m[0, 1]) / (4 * w), w])
        return pq, pt, X + x[nd:].reshape(P, 3)

    e0 = _residuals(d, *state(np.zeros(n)))
    Jm = np.zeros((3 * E, n))
    h = 1e-6
    for j in range(n):
        dx = np.zeros(n); dx[j] = h
        Jm[:, j] = ((_residuals(d, *state(dx)) - _residuals(d, *state(-dx))) / (2 * h)).ravel()
    # g2o convention: e = z - h(x); J = de/dx ; H = J^T W J ; b = -J^T W e
    w0 = np.repeat(d["edge_info"], 3)
    chi = (e0 ** 2 * d["edge_info"][:, None]).sum(1)
    delta = np.where(d["edge_obs"][:, 2] >= 0, np.float32(np.sqrt(7.815)), np.float32(np.sqrt(5.991))).astype(np.float64)
    rho1 = np.where(chi <= delta ** 2, 1.0, delta / np.sqrt(np.maximum(chi, 1e-300))) if robust else np.ones(E)
    Wd = w0 * np.repeat(rho1, 3)
    Hm = Jm.T @ (Wd[:, None] * Jm)
    bm = -Jm.T @ (Wd * e0.ravel())
    lam = 1e-5 * np.abs(np.diag(Hm)).max()
    x_ref = np.linalg.solve(Hm + lam * np.eye(n), bm)

    p = T.Problem(d)
    o = oracle_mod.ba_default_options()
    xd = np.zeros(nd); xl = np.zeros(3 * P)
    got = oracle_mod.ba_lib().ba_oracle_first_step(C.byref(p.c), C.byref(o), lam, robust, xd.ctypes.data_as(C.c_void_p), nd,
                                                    xl.ctypes.data_as(C.c_void_p))
    assert got == nd
    x = np.concatenate([xd, xl])
    # the oracle uses the reference's float invz in the residual -> agreement to ~1e-5 of the step size
    assert np.abs(x - x_ref).max() < 2e-4 * np.abs(x_ref).max() + 1e-9


def test_noise_free_problem_converges_to_ground_truth(oracle_mod):
    d = synth.make_ba_problem(8, 400, 5, seed=5, noise=False)
    rng = np.random.default_rng(1)
    d["points"] = d["points"] + rng.normal(0, 0.03, d["points"].shape)
    d["pose_t"][1:] += rng.normal(0, 0.01, d["pose_t"][1:].shape)
    o = oracle_mod.ba_default_options()
    o.iterations[0] = 30; o.iterations[1] = 30
    p, r, st = oracle_mod.ba_solve(d, o)
    assert st == 0
    assert r.c.chi2_round[1] < 1e-3 * max(1.0, r.c.chi2_initial) or r.c.chi2_round[1] < 1.0
    # gauge: pose 0 fixed at its float32-rounded ground truth -> solution is the ground truth up to that rounding
    assert np.abs(p["pose_t"].reshape(-1, 3) - d["gt_pose_t"]).max() < 2e-3
    assert r.edge_outlier.sum() == 0


def test_lm_bookkeeping_and_gates(oracle_mod):
    d = synth.make_ba_problem(10, 800, 6, seed=1)
    p, r, st = oracle_mod.ba_solve(d)
    tr = r.trace_rows
    assert st == 0 and list(r.c.iterations_run) == [5, 10] and r.c.trials_run == len(tr)
    acc = tr[:, 4] == 1
    assert (tr[acc, 2] < tr[acc, 1]).all()                    # accepted trials decrease the robust chi2
    assert r.c.chi2_initial == tr[0, 1]
    # lambda_0 = tau * max diag, then x1/3 per fully successful step (rho ~ 1): levenberg.cpp:129-141
    assert np.allclose(tr[1:5, 0] / tr[0:4, 0], 1 / 3, rtol=1e-6)
    # erase list = chi2 gate or negative depth; the 3 % gross outliers must be in it
    frac = r.edge_outlier.mean()
    assert 0.03 < frac < 0.12
    assert ((r.edge_chi2 > 7.815) <= (r.edge_outlier == 1)).all()
    # fixed pose untouched, free poses moved
    assert (p["pose_t"].reshape(-1, 3)[0] == d["pose_t"][0]).all()
    assert np.abs(p["pose_t"].reshape(-1, 3)[1:] - d["pose_t"][1:]).max() > 1e-5
    # stop flag set before the start: nothing happens (src/Optimizer.cc:620-622)
    stop = np.ones(1, np.uint8)
    p2, r2, st2 = oracle_mod.ba_solve(d, stop=stop)
    assert st2 == 6 and (p2["pose_t"].reshape(-1, 3) == d["pose_t"]).all()


def test_dynamic_problem_runs_and_improves(oracle_mod):
    d = synth.make_ba_problem(12, 600, 6, seed=9, humans=2, human_poses=4)
    assert len(d["joints"]) == 2 * 4 * 14 and len(d["redge_i"]) == 2 * 4 * 14 and len(d["medge_p1"]) == 2 * 3 * 5
    p, r, st = oracle_mod.ba_solve(d)
    assert st == 0 and r.c.iterations_run[0] >= 1
    assert r.c.chi2_round[0] < r.c.chi2_initial
    # bone lengths stay near the 1.7 m skeleton's, motions near 1.2 m/s
    assert np.abs(p["dists"] - d["dists"]).max() < 1.5   # rigidity information 20 is weak against pixel residuals
    assert np.all(np.linalg.norm(p["motion_t"].reshape(-1, 3), axis=1) < 3.0)


def test_pose_optimization_oracle(oracle_mod):
    cam, frames, gt = synth.make_pose_frames(4, 600, seed=7)
    pb = oracle_mod.pose_optimize(cam, frames)
    init = np.array([f["pose_t"] for f in frames])
    big = [0, 2, 3]
    assert (np.abs(pb.pose_t[big] - gt[big]).max(1) < 0.02).all()
    assert (np.abs(pb.pose_t[big] - gt[big]).max(1) < np.abs(init[big] - gt[big]).max(1)).all()
    # ~10 % gross outliers + the chi2 tail; the return value is nInitialCorrespondences - nBad
    for f in big:
        a, b = pb.frame_ptr[f], pb.frame_ptr[f + 1]
        assert pb.n_inliers[f] == (b - a) - pb.outlier[a:b].sum()
        assert 0.08 < pb.outlier[a:b].mean() < 0.3
    # fewer than 3 correspondences: returns 0 and leaves the pose alone (src/Optimizer.cc:345-346)
    tiny = dict(frames[0]); tiny["xw"] = tiny["xw"][:2]; tiny["obs"] = tiny["obs"][:2]; tiny["inv_sigma2"] = tiny["inv_sigma2"][:2]
    pb2 = oracle_mod.pose_optimize(cam, [tiny])
    assert pb2.n_inliers[0] == 0 and (pb2.pose_t[0] == frames[0]["pose_t"]).all()


def _residuals_g2o(d, pose_q, pose_t, points):
    """Residual with the reference's arithmetic quirk: EdgeStereoSE3ProjectXYZ::cam_project keeps 1 / z in a float
    (Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:152-161); the monocular edge divides in double."""
    R = np.array([_q2R(q) for q in pose_q])
    Xc = np.einsum("eij,ej->ei", R[d["edge_pose"]], points[d["edge_point"]]) + pose_t[d["edge_pose"]]
    mono = d["edge_obs"][:, 2] < 0
    invz = np.where(mono, 1.0 / Xc[:, 2], (1.0 / Xc[:, 2]).astype(np.float32).astype(np.float64))
    u = Xc[:, 0] * invz * d["fx"] + d["cx"]
    v = Xc[:, 1] * invz * d["fy"] + d["cy"]
    ur = u - d["bf"] * invz
    e = d["edge_obs"] - np.stack([u, v, ur], 1)
    e[mono, 2] = 0
    return e


@pytest.mark.parametrize("seed,mono_frac,point_noise,pose_noise,tau,its", [(23, 0.0, 0.05, 0.02, 1e-5, 5), (26, 0.7, 8.0, 0.5, 1e-10, 6)],
                         ids=["all_accepted", "with_rejections"])
def test_lm_trajectory_matches_an_independent_numpy_levenberg(oracle_mod, seed, mono_frac, point_noise, pose_noise, tau, its):
    """The whole LM control loop of g2o (optimization_algorithm_levenberg.cpp:61-189: lambda_0 = tau * max diag, trial loop with
    push / pop, rho = (chi2 - chi2_trial) / (x.(lambda x + b) + 1e-3), lambda *= max(1/3, min(2/3, 1 - (2 rho - 1)^3)) or
    *= ni, ni *= 2) restated independently: dense normal equations over ALL unknowns (no Schur complement), Jacobian by central
    differences of the smooth projection model, Huber IRLS weights, numpy solve.  The oracle's trace (lambda, chi2 before /
    after, accept flag per trial) must follow it."""
    d = synth.make_ba_problem(4, 40, 4, seed=seed, mono_frac=mono_frac)
    rng = np.random.default_rng(seed - 20)
    d["pose_t"][1:] += rng.normal(0, pose_noise, d["pose_t"][1:].shape)      # far enough from the optimum for several steps
    d["points"] = d["points"] + rng.normal(0, point_noise, d["points"].shape)
    K, P, E = len(d["pose_t"]), len(d["points"]), len(d["edge_pose"])
    free = [k for k in range(K) if not d["pose_fixed"][k]]
    off = {k: 6 * i for i, k in enumerate(free)}
    nd = 6 * len(free); n = nd + 3 * P
    delta = np.where(d["edge_obs"][:, 2] >= 0, np.float32(np.sqrt(7.815)), np.float32(np.sqrt(5.991))).astype(np.float64)

    def apply(st, x):
        pq, pt, X = st[0].copy(), st[1].copy(), st[2].copy()
        for k in free:
            Tm = np.eye(4); Tm[:3, :3] = _q2R(pq[k]); Tm[:3, 3] = pt[k]
            Tn = _se3_exp(x[off[k]:off[k] + 6]) @ Tm
            m = Tn[:3, :3]
            w = np.sqrt(max(0, 1 + m[0, 0] + m[1, 1] + m[2, 2])) / 2
            pq[k] = np.array([(m[2, 1] - m[1, 2]) / (4 * w), (m[0, 2] - m[2, 0]) / (4 * w), (m[1, 0] - m[0, 1]) / (4 * w), w])
            pt[k] = Tn[:3, 3]
        return pq, pt, X + x[nd:].reshape(P, 3)

    def robust_chi2(st):
        e = _residuals_g2o(d, *st)
        chi = (e ** 2 * d["edge_info"][:, None]).sum(1)
        rho = np.where(chi <= delta ** 2, chi, 2 * delta * np.sqrt(chi) - delta ** 2)
        return float(rho.sum()), e, chi

    st = (d["pose_q"].copy(), d["pose_t"].copy(), d["points"].copy())
    lam, ni, trace = None, 2.0, []
    for it in range(its):
        cur, e0, chi = robust_chi2(st)
        Jm = np.zeros((3 * E, n)); h = 1e-6
        for j in range(n):
            dx = np.zeros(n); dx[j] = h
            Jm[:, j] = ((_residuals(d, *apply(st, dx)) - _residuals(d, *apply(st, -dx))) / (2 * h)).ravel()
        rho1 = np.where(chi <= delta ** 2, 1.0, delta / np.sqrt(np.maximum(chi, 1e-300)))
        Wd = np.repeat(d["edge_info"] * rho1, 3)
        Hm = Jm.T @ (Wd[:, None] * Jm); bm = -Jm.T @ (Wd * e0.ravel())
        if it == 0:
            lam = tau * np.abs(np.diag(Hm)).max()
        q = 0
        while True:
            x = np.linalg.solve(Hm + lam * np.eye(n), bm)
            trial = apply(st, x)
            tmp = robust_chi2(trial)[0]
            rho = (cur - tmp) / (float(x @ (lam * x + bm)) + 1e-3)
            ok = rho > 0 and np.isfinite(tmp)
            trace.append((lam, cur, tmp, ok))
            if ok:
                lam *= max(1.0 / 3.0, min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)); ni = 2.0
                st, cur = trial, tmp
            else:
                lam *= ni; ni *= 2
            q += 1
            if not (rho < 0 and q < 10):
                break
    o = oracle_mod.ba_default_options()
    o.iterations[0] = its; o.iterations[1] = 0; o.tau = tau
    p, r, status = oracle_mod.ba_solve(d, o)
    tr = r.trace_rows
    assert status == 0 and len(tr) == len(trace) >= 5
    if mono_frac > 0:
        assert not all(t[3] for t in trace), "this case is meant to exercise rejected trials"
    ref = np.array([(t[0], t[1], t[2]) for t in trace])
    assert np.allclose(tr[:, 0], ref[:, 0], rtol=1e-5), (tr[:, 0], ref[:, 0])       # lambda of every trial
    # chi2 before / after; the far-from-optimum case (chi2 ~ 1e6, rejected overshoots) amplifies the finite-difference Jacobian
    assert np.allclose(tr[:, 1:3], ref[:, 1:3], rtol=1e-6 if mono_frac == 0 else 1e-5)
    assert [bool(a) for a in tr[:, 4]] == [t[3] for t in trace]
    assert np.abs(p["pose_t"] - st[1]).max() < (1e-6 if mono_frac == 0 else 1e-4)
    assert np.abs(p["points"] - st[2]).max() < (1e-5 if mono_frac == 0 else 1e-3)


def test_dynamic_lm_trajectory_with_rigidity_edges_matches_numpy(oracle_mod):
    """The articulated part of LocalBundleAdjustmentHumanTrajactory that is a true least-squares model -- non-marginalised
    joint vertices with stereo observations (src/Optimizer.cc:1775-1843), VertexDistanceDouble bone lengths and
    EdgeRigidBodyDouble e = |Ji - Jj| - d (include/g2o_edge_rigidbody.h:67-149) -- against the same independent numpy LM:
    dense normal equations over [poses | bone lengths | joints | points], finite-difference Jacobian.  (The motion edges are
    left out here: by convention D.6 their Jacobian is deliberately not the derivative of the residual.)"""
    d = synth.make_ba_problem(8, 60, 4, seed=31, humans=1, human_poses=2)
    for k in ("medge_p1", "medge_p2", "medge_motion"):
        d[k] = np.zeros(0, np.int32)
    d["medge_dt"] = np.zeros(0); d["medge_info"] = np.zeros(0)
    d["motion_q"] = np.zeros((0, 4)); d["motion_t"] = np.zeros((0, 3))
    rng = np.random.default_rng(5)
    d["joints"] = d["joints"] + rng.normal(0, 0.05, d["joints"].shape)
    d["dists"] = d["dists"] + rng.normal(0, 0.05, d["dists"].shape)
    K, P, E = len(d["pose_t"]), len(d["points"]), len(d["edge_pose"])
    NJ, ND = len(d["joints"]), len(d["dists"])
    free = [k for k in range(K) if not d["pose_fixed"][k]]
    off = {k: 6 * i for i, k in enumerate(free)}
    nd = 6 * len(free)
    o_d, o_j, o_p = nd, nd + ND, nd + ND + 3 * NJ
    n = o_p + 3 * P
    opt = oracle_mod.ba_default_options()
    opt.iterations[0] = 5; opt.iterations[1] = 0
    hs, hr = opt.huber_stereo, opt.huber_rigid

    def apply(st, x):
        pq, pt, X, J, D = [a.copy() for a in st]
        for k in free:
            Tm = np.eye(4); Tm[:3, :3] = _q2R(pq[k]); Tm[:3, 3] = pt[k]
            Tn = _se3_exp(x[off[k]:off[k] + 6]) @ Tm
            m = Tn[:3, :3]
            w = np.sqrt(max(0, 1 + m[0, 0] + m[1, 1] + m[2, 2])) / 2
            pq[k] = np.array([(m[2, 1] - m[1, 2]) / (4 * w), (m[0, 2] - m[2, 0]) / (4 * w), (m[1, 0] - m[0, 1]) / (4 * w), w])
            pt[k] = Tn[:3, 3]
        return pq, pt, X + x[o_p:].reshape(P, 3), J + x[o_j:o_p].reshape(NJ, 3), D + x[o_d:o_j]

    def residuals(st, smooth):
        pq, pt, X, J, D = st
        f = _residuals if smooth else _residuals_g2o
        e_s = f(d, pq, pt, X)
        dj = dict(d, edge_pose=d["jedge_pose"], edge_point=d["jedge_joint"], edge_obs=d["jedge_obs"])
        e_j = f(dj, pq, pt, J)
        e_r = np.linalg.norm(J[d["redge_i"]] - J[d["redge_j"]], axis=1) - D[d["redge_dist"]]
        return e_s, e_j, e_r

    def parts(st, smooth=False):
        e_s, e_j, e_r = residuals(st, smooth)
        e = np.concatenate([e_s.ravel(), e_j.ravel(), e_r])
        info = np.concatenate([np.repeat(d["edge_info"], 3), np.repeat(d["jedge_info"], 3), d["redge_info"]])
        chi = np.concatenate([(e_s ** 2 * d["edge_info"][:, None]).sum(1), (e_j ** 2 * d["jedge_info"][:, None]).sum(1), e_r ** 2 * d["redge_info"]])
        dl = np.concatenate([np.where(d["edge_obs"][:, 2] >= 0, hs, opt.huber_mono), np.full(len(e_j), hs), np.full(len(e_r), hr)])
        reps = np.concatenate([np.full(len(e_s), 3), np.full(len(e_j), 3), np.full(len(e_r), 1)])
        return e, info, chi, dl, reps

    def robust_chi2(st):
        e, info, chi, dl, reps = parts(st)
        return float(np.where(chi <= dl ** 2, chi, 2 * dl * np.sqrt(chi) - dl ** 2).sum())

    st = (d["pose_q"].copy(), d["pose_t"].copy(), d["points"].copy(), d["joints"].copy(), d["dists"].copy())
    lam, ni, trace = None, 2.0, []
    for it in range(5):
        cur = robust_chi2(st)
        e0, info, chi, dl, reps = parts(st)
        h = 1e-6
        Jm = np.zeros((len(e0), n))
        for j in range(n):
            dx = np.zeros(n); dx[j] = h
            Jm[:, j] = (parts(apply(st, dx), True)[0] - parts(apply(st, -dx), True)[0]) / (2 * h)
        rho1 = np.where(chi <= dl ** 2, 1.0, dl / np.sqrt(np.maximum(chi, 1e-300)))
        Wd = info * np.repeat(rho1, reps)
        Hm = Jm.T @ (Wd[:, None] * Jm); bm = -Jm.T @ (Wd * e0)
        if it == 0:
            lam = opt.tau * np.abs(np.diag(Hm)).max()
        q = 0
        while True:
            x = np.linalg.solve(Hm + lam * np.eye(n), bm)
            trial = apply(st, x)
            tmp = robust_chi2(trial)
            rho = (cur - tmp) / (float(x @ (lam * x + bm)) + 1e-3)
            ok = rho > 0 and np.isfinite(tmp)
            trace.append((lam, cur, tmp, ok))
            if ok:
                lam *= max(1.0 / 3.0, min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)); ni = 2.0
                st, cur = trial, tmp
            else:
                lam *= ni; ni *= 2
            q += 1
            if not (rho < 0 and q < 10):
                break
    p, r, status = oracle_mod.ba_solve(d, opt)
    tr = r.trace_rows
    assert status == 0 and len(tr) == len(trace) >= 5
    ref = np.array([(t[0], t[1], t[2]) for t in trace])
    assert np.allclose(tr[:, 0], ref[:, 0], rtol=1e-5) and np.allclose(tr[:, 1:3], ref[:, 1:3], rtol=1e-6)
    assert [bool(a) for a in tr[:, 4]] == [t[3] for t in trace]
    assert np.abs(p["joints"] - st[3]).max() < 1e-5 and np.abs(p["dists"] - st[4]).max() < 1e-5
    assert np.abs(p["pose_t"] - st[1]).max() < 1e-6


def test_full_dynamic_window_initial_chi2_matches_numpy(oracle_mod):
    """All four edge families of LocalBundleAdjustmentHumanTrajactory at once, incl. LandmarkMotionTernaryEdge
    (include/g2o_dyn_slam3d.h:65-76: e = p1 - M^-1 p2 with M = [R | t * delta_t]): the robust chi2 the LM starts from, for
    non-trivial motion estimates and delta_t != 1, against a plain numpy evaluation."""
    d = synth.make_ba_problem(10, 80, 4, seed=41, humans=2, human_poses=3)
    rng = np.random.default_rng(8)
    nm = len(d["motion_t"])
    d["motion_t"] = rng.normal(0, 0.3, (nm, 3))
    q = np.concatenate([rng.normal(0, 0.05, (nm, 3)), np.ones((nm, 1))], 1)
    d["motion_q"] = q / np.linalg.norm(q, axis=1, keepdims=True)
    d["medge_dt"] = rng.uniform(0.5, 1.5, len(d["medge_dt"]))
    opt = oracle_mod.ba_default_options()
    opt.iterations[0] = 1; opt.iterations[1] = 0
    _, r, status = oracle_mod.ba_solve(d, opt)
    assert status == 0
    # the same window without its motion edges: the difference of the two initial chi2 isolates the motion family exactly
    d0 = dict(d, medge_p1=np.zeros(0, np.int32), medge_p2=np.zeros(0, np.int32), medge_motion=np.zeros(0, np.int32),
              medge_dt=np.zeros(0), medge_info=np.zeros(0), motion_q=np.zeros((0, 4)), motion_t=np.zeros((0, 3)))
    _, r0, status0 = oracle_mod.ba_solve(d0, opt)
    assert status0 == 0

    def rho(chi, dl):
        return np.where(chi <= dl ** 2, chi, 2 * dl * np.sqrt(chi) - dl ** 2).sum()

    e_s = _residuals_g2o(d, d["pose_q"], d["pose_t"], d["points"])
    chi_s = (e_s ** 2 * d["edge_info"][:, None]).sum(1)
    dl_s = np.where(d["edge_obs"][:, 2] >= 0, opt.huber_stereo, opt.huber_mono)
    dj = dict(d, edge_pose=d["jedge_pose"], edge_point=d["jedge_joint"], edge_obs=d["jedge_obs"])
    e_j = _residuals_g2o(dj, d["pose_q"], d["pose_t"], d["joints"])
    chi_j = (e_j ** 2 * d["jedge_info"][:, None]).sum(1)
    J = d["joints"]
    e_r = np.linalg.norm(J[d["redge_i"]] - J[d["redge_j"]], axis=1) - d["dists"][d["redge_dist"]]
    chi_r = e_r ** 2 * d["redge_info"]
    Rm = np.array([_q2R(qq) for qq in d["motion_q"]])[d["medge_motion"]]
    tm = d["motion_t"][d["medge_motion"]] * d["medge_dt"][:, None]
    e_m = J[d["medge_p1"]] - np.einsum("eji,ej->ei", Rm, J[d["medge_p2"]] - tm)      # p1 - R^T (p2 - t * dt)
    chi_m = (e_m ** 2).sum(1) * d["medge_info"]
    total = rho(chi_s, dl_s) + rho(chi_j, opt.huber_stereo) + rho(chi_r, opt.huber_rigid) + rho(chi_m, opt.huber_motion)
    assert len(chi_m) > 0 and (chi_m > opt.huber_motion ** 2).any()              # the Huber branch of the motion edges is exercised
    # the float 1 / z of the stereo projection makes the static part sensitive to the last bit of z (~1e-8 relative) ...
    assert abs(r.c.chi2_initial - total) < 1e-7 * total
    # ... the motion family alone is exact
    assert abs((r.c.chi2_initial - r0.c.chi2_initial) - rho(chi_m, opt.huber_motion)) < 1e-9 * rho(chi_m, opt.huber_motion)


def test_pose_optimization_matches_an_independent_numpy_schedule(oracle_mod):
    """Optimizer::PoseOptimization (src/Optimizer.cc:232-429) restated independently in numpy: four rounds, each restarted
    from the initial pose, ten g2o LM iterations per round on the edges still at level 0 (6 x 6 normal equations, Jacobian by
    central differences), re-classification of EVERY edge after each round (outliers re-evaluated at the new pose, inliers
    judged by the error of the last evaluated trial, float chi2 against 5.991f / 7.815f), robust kernel off in the last round."""
    cam, frames, _ = synth.make_pose_frames(3, 300, seed=17)
    pb = oracle_mod.pose_optimize(cam, frames)
    th_m, th_s = np.float32(5.991), np.float32(7.815)
    d_m, d_s = float(np.float32(np.sqrt(5.991))), float(np.float32(np.sqrt(7.815)))
    for f, fr in enumerate(frames):
        n = len(fr["xw"])
        if n < 10:
            continue                                   # the < 10 edges early exit has its own test above
        X = fr["xw"].astype(np.float64); obs = fr["obs"].astype(np.float64); w = fr["inv_sigma2"].astype(np.float64)
        stereo = ~(fr["obs"][:, 2] < 0)
        dl = np.where(stereo, d_s, d_m)

        def err(q, t, smooth=False):
            Xc = X @ _q2R(q).T + t
            invz = 1.0 / Xc[:, 2]
            if not smooth:
                invz = np.where(stereo, invz.astype(np.float32).astype(np.float64), invz)
            u = Xc[:, 0] * invz * cam["fx"] + cam["cx"]; v = Xc[:, 1] * invz * cam["fy"] + cam["cy"]
            e = obs - np.stack([u, v, u - cam["bf"] * invz], 1)
            e[~stereo, 2] = 0
            return e

        def oplus(q, t, x):
            Tm = np.eye(4); Tm[:3, :3] = _q2R(q); Tm[:3, 3] = t
            Tn = _se3_exp(x) @ Tm
            m = Tn[:3, :3]
            ww = np.sqrt(max(0, 1 + m[0, 0] + m[1, 1] + m[2, 2])) / 2
            return np.array([(m[2, 1] - m[1, 2]) / (4 * ww), (m[0, 2] - m[2, 0]) / (4 * ww), (m[1, 0] - m[0, 1]) / (4 * ww), ww]), Tn[:3, 3]

        q0, t0 = np.asarray(fr["pose_q"], np.float64), np.asarray(fr["pose_t"], np.float64)
        level = np.zeros(n, bool)          # True = level 1 (left out of the optimisation)
        outlier = np.zeros(n, bool)
        robust = True
        for rnd in range(4):
            q, t = q0.copy(), t0.copy()
            act = ~level

            def rchi(e):
                chi = (e ** 2).sum(1) * w
                r = np.where(chi <= dl ** 2, chi, 2 * dl * np.sqrt(chi) - dl ** 2) if robust else chi
                return float(r[act].sum()), chi

            last_e = err(q, t)
            lam, ni, n_bad = None, 2.0, 0
            for it in range(10):
                e0 = err(q, t); last_e = e0
                cur, chi = rchi(e0); ini = cur
                Jm = np.zeros((3 * n, 6)); h = 1e-6
                for j in range(6):
                    dx = np.zeros(6); dx[j] = h
                    Jm[:, j] = ((err(*oplus(q, t, dx), True) - err(*oplus(q, t, -dx), True)) / (2 * h)).ravel()
                rho1 = np.where(chi <= dl ** 2, 1.0, dl / np.sqrt(np.maximum(chi, 1e-300))) if robust else np.ones(n)
                Wd = np.repeat(w * rho1 * act, 3)
                Hm = Jm.T @ (Wd[:, None] * Jm); bm = -Jm.T @ (Wd * e0.ravel())
                if it == 0:
                    lam = 1e-5 * np.abs(np.diag(Hm)).max()
                qn, rho = 0, 0.0
                while True:
                    x = np.linalg.solve(Hm + lam * np.eye(6), bm)
                    q2, t2 = oplus(q, t, x)
                    last_e = err(q2, t2)
                    tmp = rchi(last_e)[0]
                    rho = (cur - tmp) / (float(x @ (lam * x + bm)) + 1e-3)
                    if rho > 0 and np.isfinite(tmp):
                        lam *= max(1.0 / 3.0, min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)); ni = 2.0
                        q, t, cur = q2, t2, tmp
                    else:
                        lam *= ni; ni *= 2
                    qn += 1
                    if not (rho < 0 and qn < 10):
                        break
                if qn == 10 or rho == 0:
                    break
                n_bad = n_bad + 1 if (ini - cur) * 1e3 < ini else 0
                if n_bad >= 3:
                    break
            # classification: former outliers are re-evaluated at the current estimate, the rest keep the last evaluated error
            e_now = err(q, t)
            e_cls = np.where(outlier[:, None], e_now, last_e)
            chi_f = ((e_cls ** 2).sum(1) * w).astype(np.float32)
            outlier = chi_f > np.where(stereo, th_s, th_m)
            level = outlier.copy()
            if rnd == 2:
                robust = False
        a, b = pb.frame_ptr[f], pb.frame_ptr[f + 1]
        assert (pb.outlier[a:b].astype(bool) == outlier).all(), f
        assert pb.n_inliers[f] == n - outlier.sum()
        assert np.abs(pb.pose_t[f] - t).max() < 1e-6 and np.abs(np.abs(pb.pose_q[f] @ q) - 1) < 1e-10


def test_out_of_range_indices_are_rejected(oracle_mod):
    """A bad vertex index in any edge family is ADB_ERR_INVALID (1), never a wild access (same contract as adb_ba_solve)."""
    from airdos_b200 import synth
    d = synth.make_ba_problem(n_kf=6, n_points=120, seed=31, humans=2, human_poses=4)
    for key, bad in (("edge_pose", 6), ("edge_point", -1), ("jedge_pose", 99), ("jedge_joint", 10 ** 6), ("redge_i", -3), ("redge_j", 10 ** 5),
                     ("redge_dist", 28), ("medge_p1", 112), ("medge_p2", -1), ("medge_motion", 2)):
        b = dict(d); b[key] = d[key].copy(); b[key][len(b[key]) // 2] = bad
        _, _, st = oracle_mod.ba_solve(b)
        assert st == 1, key
    assert oracle_mod.ba_solve(d)[2] == 0
