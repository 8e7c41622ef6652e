"""Writes tests/golden/lm_ref.npz: results of the REFERENCE's own Levenberg-Marquardt control, Huber kernel and per-edge quadratic form
(optimization_algorithm_levenberg.cpp whole, SparseOptimizer::optimize, RobustKernelHuber::setDelta / robustify, BaseEdge::chi2 /
robustInformation, BaseBinaryEdge / BaseUnaryEdge::constructQuadraticForm: compiled from /root/reference by `make -C oracle ref`,
oracle/ref_lm.cpp).  The LM control runs over the oracle's arithmetic through function pointers (computeActiveErrors, buildSystem, solve,
update, push / pop are the oracle's steps), so a trace recorded here is what the reference's control flow does with that arithmetic.
Run in the build container (needs /root/reference):

    python oracle/gen_ref_lm_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_lm.so")

# (seed, key-frames, points, make_ba_problem kwargs, point noise, translation noise, iterations, robust, max_trials)
LM_CASES = [
    (1, 10, 800, {}, 0.0, 0.0, 10, 1, 10),                                  # the shipped schedule: all steps accepted
    (9, 12, 600, dict(humans=2, human_poses=4), 0.0, 0.0, 10, 1, 10),       # articulated window (rigidity + motion edges)
    (11, 5, 60, dict(mono_frac=0.5), 0.0, 0.0, 15, 0, 10),                  # no robust kernel (round 2 of the schedule)
    (26, 6, 120, dict(mono_frac=0.7), 8.0, 0.5, 8, 1, 10),                  # a rejected trial (lambda *= ni)
    (28, 6, 120, {}, 20.0, 2.0, 12, 0, 10),                                 # three rejections, ni doubling
    (28, 6, 120, {}, 20.0, 2.0, 12, 0, 2),                                  # the same with maxTrialsAfterFailure = 2: Terminate
    (30, 6, 120, {}, 0.0, 0.0, 20, 1, 10),                                  # the "_nBad >= 3" stop (src modification by Raul Mur-Artal)
    (29, 6, 120, dict(humans=1, human_poses=3), 1.0, 0.3, 10, 1, 10),
    (31, 6, 120, {}, 50.0, 5.0, 10, 1, 10),
]
N_QF = 600


def make_lm_case(i):
    from airdos_b200 import synth
    seed, kf, pts, kw, pn, tn, its, robust, max_trials = LM_CASES[i]
    d = synth.make_ba_problem(kf, pts, 4 if pts < 200 else 6, seed=seed, **kw)
    rng = np.random.default_rng(seed)
    d["pose_t"][1:] += rng.normal(0, tn, d["pose_t"][1:].shape)
    d["points"] = d["points"] + rng.normal(0, pn, d["points"].shape)
    return d, its, bool(robust), max_trials


def make_qf_case(t):
    """Jacobians / residual / information / kernel of one edge: mono and stereo, inliers and outliers of the kernel, pyramid-level weights."""
    rng = np.random.default_rng(93000 + t)
    dim = 2 + (t & 1); robust = (t >> 1) & 1; fixed = t % 7 == 0
    Ji = np.zeros((3, 3)); Jj = np.zeros((3, 6)); er = np.zeros(3)
    Ji[:dim] = rng.normal(0, 50, (dim, 3)); Jj[:dim] = rng.normal(0, 200, (dim, 6)); er[:dim] = rng.normal(0, rng.choice([0.5, 3, 30]), dim)
    w0 = float(np.float32(1.2) ** (-2 * int(rng.integers(0, 8))))
    delta = float(np.float32(np.sqrt(7.815 if dim == 3 else 5.991)))
    return dim, Ji, Jj, er, w0, delta, robust, fixed


N_MULTI = 300


def make_multi_case(t):
    """A rigidity edge (1 residual; joint, joint, bone length) or a motion edge (3 residuals; joint, joint, SE3 motion) with random Jacobians."""
    rng = np.random.default_rng(94000 + t)
    dim, dims = (1, [3, 3, 1]) if t & 1 else (3, [3, 3, 6])
    Js = [rng.normal(0, 3, (dim, d)) for d in dims]
    er = np.zeros(3); er[:dim] = rng.normal(0, rng.choice([0.05, 1, 5]), dim)
    return dim, dims, Js, er, 20.0, (1.0 if dim == 1 else 2.0), (t >> 1) & 1


def huber_inputs():
    """The deltas the Optimizer sets (float thHuber* promoted to double, src/Optimizer.cc:95-96, 538-539, 742-744) x squared errors around
    delta^2: far inside, far outside, and a fine sweep across [delta^2 - 1e-6, delta^2 + 1e-6] which contains both the double delta^2
    and the float `dsqr` the reference compares with."""
    deltas = [float(np.float32(np.sqrt(v))) for v in (5.99, 5.991, 7.815, 4.0, 1.0, 16.0)] + [1.0, 2.0, 0.5]
    d, e = [], []
    for dl in deltas:
        sq = dl * dl
        es = np.concatenate([[0.0, 1e-12, 0.3 * sq, 0.999 * sq, 1.001 * sq, 2 * sq, 10 * sq, 1e4 * sq, 1e12], sq + np.linspace(-1e-6, 1e-6, 81),
                             [float(np.float32(sq)), np.nextafter(float(np.float32(sq)), 0), np.nextafter(float(np.float32(sq)), 1e9), sq]])
        d += [dl] * len(es); e += list(es)
    return np.array(d), np.array(e)


def main():
    import oracle
    oracle.build()
    L = C.CDLL(LIB)
    out = {}
    same = 0
    for i in range(len(LM_CASES)):
        d, its, robust, mt = make_lm_case(i)
        o = oracle.ba_default_options(); o.max_trials = mt
        a = oracle.LmSession(d, o, robust); b = oracle.LmSession(d, o, robust)
        it_o, rows_o, lam_o = a.optimize(its)
        it_r, rows_r, lam_r, n_eval, tau = oracle.ref_lm_optimize(L, b, its)
        ok = it_o == it_r and rows_o.shape == rows_r.shape and (rows_o == rows_r).all() and lam_o == lam_r and (a.state() == b.state()).all()
        same += int(ok)
        out[f"lm{i}_rows"] = rows_r; out[f"lm{i}_meta"] = np.array([it_r, lam_r, n_eval, tau]); out[f"lm{i}_state"] = b.state()
        print(f"case {i}: {it_r} iterations, {len(rows_r)} trials ({int((rows_r[:, 3] == 0).sum())} rejected), lambda_final {lam_r:.6g}, "
              f"tau {tau:g}: oracle loop {'identical' if ok else 'DIFFERENT'}")
        a.close(); b.close()
    print(f"LM control: the oracle's loop equals the reference's in {same} of {len(LM_CASES)} cases (trial rows, final lambda, final state: bit for bit)")
    hd, he = huber_inputs()
    rho = np.array([oracle.huber(a, b, lib=L) for a, b in zip(hd, he)])
    rho_o = np.array([oracle.huber(a, b) for a, b in zip(hd, he)])
    print(f"Huber: {len(hd)} (delta, e2) pairs, oracle bit-identical in {int((rho == rho_o).all(1).sum())}")
    out["huber_delta"], out["huber_e2"], out["huber_rho"] = hd, he, rho
    worst = 0.0
    for t in range(N_QF):
        dim, Ji, Jj, er, w0, delta, robust, fixed = make_qf_case(t)
        r = oracle.edge_quadratic_form(dim, Ji[:dim], Jj[:dim], er, w0, delta, robust, fixed, lib=L)
        q = oracle.edge_quadratic_form(dim, Ji, Jj, er, w0, delta, robust, fixed)
        ru = oracle.pose_quadratic_form(dim, Jj[:dim], er, w0, delta, robust, lib=L)
        qu = oracle.pose_quadratic_form(dim, Jj, er, w0, delta, robust)
        for x, y in list(zip(r, q)) + list(zip(ru, qu)):
            worst = max(worst, float(np.abs(x - y).max() / max(np.abs(x).max(), 1e-300)))
        out[f"qf{t}"] = np.concatenate(list(r) + list(ru))
    print(f"quadratic forms: {N_QF} edges (binary + unary), worst relative difference oracle vs reference {worst:.3g}")
    worst = 0.0
    for t in range(N_MULTI):
        c = make_multi_case(t)
        Hr, br = oracle.multi_quadratic_form(*c, lib=L)
        Ho, bo = oracle.multi_quadratic_form(*c)
        worst = max(worst, float(np.abs(Hr - Ho).max() / max(np.abs(Hr).max(), 1e-300)), float(np.abs(br - bo).max() / max(np.abs(br).max(), 1e-300)))
        out[f"mq{t}"] = np.concatenate([Hr.ravel(), br])
    print(f"multi-edge quadratic forms: {N_MULTI} rigidity / motion edges, worst relative difference oracle vs reference {worst:.3g}")
    path = os.path.join(ROOT, "tests", "golden", "lm_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
