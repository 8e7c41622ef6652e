"""Writes tests/golden/schur_ref.npz: results of the REFERENCE's own BlockSolver<BlockSolverTraits<6, 3>>::solve()
(Thirdparty/g2o/g2o/core/block_solver.hpp:353-483 -- Schur complement over the landmarks, reduced right-hand side, landmark
back-substitution -- compiled from /root/reference by `make -C oracle ref`, oracle/ref_schur.cpp) on the normal equations the oracle
assembles for seeded static windows (the per-edge blocks of that assembly are pinned separately: tests/test_ref_lm.py).  Stored per case
and lambda: the whole solution x, the reduced right-hand side and the reduced matrix.  The linear solver behind the reduced system is
not the reference's (Eigen LDLT there, a plain Cholesky in ref_schur.cpp).  Run in the build container (needs /root/reference):

    python oracle/gen_ref_schur_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_schur.so")

# (seed, key-frames, points, observations per point, make_ba_problem kwargs, robust, fraction of edges switched off, lambdas as multiples of
#  tau * max diagonal -- the controller's first lambda -- from the damped to the almost undamped system)
SCHUR_CASES = [
    (101, 10, 800, 6, {}, 1, 0.0, (1.0, 1e3)),
    (102, 6, 150, 4, dict(mono_frac=0.6), 1, 0.0, (1.0, 1e-3)),
    (103, 14, 1200, 6, {}, 0, 0.1, (1.0, 1e2)),              # no kernel; a tenth of the edges at level 1 (some points lose all their edges)
    (104, 4, 40, 3, dict(mono_frac=1.0), 1, 0.0, (1.0,)),    # monocular only
    (105, 20, 2500, 5, {}, 1, 0.02, (1.0, 10.0)),
    (106, 8, 300, 3, {}, 1, 0.5, (1.0, 1e-2)),               # half of the edges off at 3 observations per point: points drop out of the system
]


# articulated windows through BlockSolverX: (seed, key-frames, points, observations per point, trajectories, poses per trajectory, lambdas)
SCHUR_X_CASES = [
    (201, 10, 600, 5, 2, 4, (1.0, 1e2)),
    (202, 16, 1500, 6, 3, 5, (1.0,)),
    (203, 6, 200, 4, 1, 3, (1.0, 1e-2)),
]


def open_session_x(oracle, i):
    from airdos_b200 import synth
    seed, kf, pts, obs, tracks, poses, _ = SCHUR_X_CASES[i]
    d = synth.make_ba_problem(kf, pts, obs, seed=seed, humans=tracks, human_poses=poses)
    rng = np.random.default_rng(seed)
    d["points"] = d["points"] + rng.normal(0, 0.02, d["points"].shape)
    d["joints"] = d["joints"] + rng.normal(0, 0.01, d["joints"].shape)
    s = oracle.LmSession(d, None, True)
    s.linearize()
    return s


def first_lambda_x(system):
    return 1e-5 * max(float(np.abs(np.diag(system["H"])).max()), float(np.abs(np.einsum("kii->ki", system["Hll"])).max()))


def make_schur_case(i):
    """-> (problem dict, robust, edge levels or None)"""
    from airdos_b200 import synth
    seed, kf, pts, obs, kw, robust, off, _ = SCHUR_CASES[i]
    d = synth.make_ba_problem(kf, pts, obs, seed=seed, **kw)
    rng = np.random.default_rng(seed)
    d["points"] = d["points"] + rng.normal(0, 0.02, d["points"].shape)            # off the optimum: a non-trivial right-hand side
    lvl = (rng.random(len(d["edge_pose"])) < off).astype(np.uint8) if off > 0 else None
    return d, bool(robust), lvl


def open_session(oracle, i):
    d, robust, lvl = make_schur_case(i)
    s = oracle.LmSession(d, None, robust)
    if lvl is not None:
        s.lib.ba_oracle_lm_set_levels.argtypes = [C.c_void_p, C.c_void_p]
        s.lib.ba_oracle_lm_set_levels(s.h, lvl.ctypes.data_as(C.c_void_p))
    s.linearize()
    return s


def first_lambda(system):
    """tau * max diagonal of the whole Hessian (optimization_algorithm_levenberg.cpp:147-161), tau = 1e-5"""
    return 1e-5 * max(float(np.abs(np.einsum("kii->ki", system["Hpp"])).max()), float(np.abs(np.einsum("kii->ki", system["Hll"])).max()))


def main():
    import oracle
    oracle.build()
    lib = C.CDLL(LIB)
    out = {}
    for i, case in enumerate(SCHUR_CASES):
        s = open_session(oracle, i)
        sysd = s.system()
        lam0 = first_lambda(sysd)
        for k, mult in enumerate(case[7]):
            ok, x, hs, bs = oracle.ref_schur_solve(lib, sysd, lam0 * mult)
            ok_o, x_o = s.solve(lam0 * mult)
            assert ok and ok_o
            rel = float(np.abs(x - x_o).max() / np.abs(x).max())
            res = float(np.abs(hs @ x_o[:len(bs)] - bs).max() / np.abs(bs).max())
            print(f"case {i} lambda x{mult:g}: {sysd['n_poses']} free poses, {sysd['n_points']} points, {len(sysd['edge_pose'])} blocks | "
                  f"oracle vs reference: |dx| / |x| = {rel:.2e}, residual of the oracle's poses in the reference's reduced system = {res:.2e}")
            out[f"c{i}_{k}_lambda"] = np.float64(lam0 * mult)
            out[f"c{i}_{k}_x"] = x; out[f"c{i}_{k}_bschur"] = bs; out[f"c{i}_{k}_hschur"] = hs
        out[f"c{i}_sizes"] = np.array([sysd["n_poses"], sysd["n_points"], len(sysd["edge_pose"])], np.int32)
        s.close()
    for i, case in enumerate(SCHUR_X_CASES):
        s = open_session_x(oracle, i)
        sysd = s.system_x()
        lam0 = first_lambda_x(sysd)
        for k, mult in enumerate(case[6]):
            ok, x, hs, bs = oracle.ref_schur_solve_x(lib, sysd, lam0 * mult)
            ok_o, x_o = s.solve(lam0 * mult)
            assert ok and ok_o
            rel = float(np.abs(x - x_o).max() / np.abs(x).max())
            res = float(np.abs(hs @ x_o[:len(bs)] - bs).max() / np.abs(bs).max())
            print(f"articulated case {i} lambda x{mult:g}: vertex widths {np.bincount(sysd['dims'])[[6, 1, 3]].tolist()} (6 / 1 / 3 wide), reduced order "
                  f"{sysd['n_dense']}, {sysd['n_points']} points | oracle vs reference: |dx| / |x| = {rel:.2e}, residual = {res:.2e}")
            out[f"x{i}_{k}_lambda"] = np.float64(lam0 * mult)
            out[f"x{i}_{k}_x"] = x; out[f"x{i}_{k}_bschur"] = bs
            if len(bs) <= 450:                                       # (the reduced matrix of the largest case would double the fixture)
                out[f"x{i}_{k}_hschur_upper"] = hs[np.triu_indices(len(bs))]
        out[f"x{i}_sizes"] = np.array([len(sysd["dims"]), sysd["n_points"], len(sysd["edge_block"]), sysd["n_dense"]], np.int32)
        s.close()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "schur_ref.npz"), **out)
    print("wrote tests/golden/schur_ref.npz")


if __name__ == "__main__":
    main()
