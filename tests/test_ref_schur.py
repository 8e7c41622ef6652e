"""Pins the Schur complement, the reduced right-hand side and the landmark back-substitution of oracle/ba_oracle.cpp (Solver::solve_trial)
to the LITERAL reference: tests/golden/schur_ref.npz holds what the reference's own BlockSolver<BlockSolverTraits<6, 3>>::solve()
(Thirdparty/g2o/g2o/core/block_solver.hpp:353-483, compiled from /root/reference: oracle/ref_schur.cpp; oracle/gen_ref_schur_golden.py wrote
the fixture) computes on the normal equations of six seeded static windows (stereo / monocular mixes, fixed key-frames, edges switched
off so that points drop out, with and without the kernel) at two damping values each, and -- instantiated as BlockSolverX -- on three
articulated windows (key-frames, motions, joints and bone lengths in the reduced system).

The per-edge blocks that go INTO those equations are pinned in tests/test_ref_lm.py (constructQuadraticForm); the LM control around them
likewise; what is not the reference's is the factorisation of the reduced matrix (Eigen's LDLT there, a Cholesky in both the oracle and
ref_schur.cpp) -- therefore the second check below, which involves no factorisation of ours at all: the oracle's pose update must satisfy
the REFERENCE's reduced system to rounding."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "schur_ref.npz")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_schur.so")
HAVE_REF = os.path.exists(REF_LIB) and os.path.isdir("/root/reference")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_ref_schur_golden", os.path.join(ROOT, "oracle", "gen_ref_schur_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    return g


def test_oracle_schur_solve_equals_the_reference_block_solver(oracle_mod):
    g = _gen()
    gold = np.load(GOLD)
    dropped = 0
    for i, case in enumerate(g.SCHUR_CASES):
        s = g.open_session(oracle_mod, i)
        sysd = s.system()
        assert list(gold[f"c{i}_sizes"]) == [sysd["n_poses"], sysd["n_points"], len(sysd["edge_pose"])], i
        dropped += int(sysd["n_points"] < case[2])
        for k in range(len(case[7])):
            lam = float(gold[f"c{i}_{k}_lambda"])
            assert lam == g.first_lambda(sysd) * case[7][k]
            ok, x = s.solve(lam)
            xr, bs, hs = gold[f"c{i}_{k}_x"], gold[f"c{i}_{k}_bschur"], gold[f"c{i}_{k}_hschur"]
            assert ok and x.shape == xr.shape
            n = len(bs)
            # the whole update (poses and landmarks) against the reference's: 1e-12 of the largest component (measured: <= 2.5e-14)
            assert np.abs(x - xr).max() <= 1e-12 * np.abs(xr).max(), (i, k, float(np.abs(x - xr).max() / np.abs(xr).max()))
            # the oracle's pose update in the reference's reduced system, no factorisation involved (measured: <= 3.2e-15)
            assert np.abs(hs @ x[:n] - bs).max() <= 1e-12 * np.abs(bs).max(), (i, k)
            # and the reference's landmark update follows from the reference's pose update through the oracle's blocks: xl = Dinv (bl - W^T xp)
            xl = np.zeros((sysd["n_points"], 3))
            acc = sysd["b"][n:].reshape(-1, 3).copy()
            np.subtract.at(acc, sysd["edge_point"], np.einsum("eij,ei->ej", sysd["W"], xr[:n].reshape(-1, 6)[sysd["edge_pose"]]))
            D = sysd["Hll"] + lam * np.eye(3)
            xl = np.linalg.solve(D, acc[..., None])[..., 0]
            assert np.abs(xl.ravel() - xr[n:]).max() <= 1e-10 * np.abs(xr[n:]).max(), (i, k)
        s.close()
    assert dropped >= 1                                          # at least one window where switched-off edges removed points from the system


def test_oracle_schur_solve_equals_the_reference_block_solver_x(oracle_mod):
    """The same for the AirDOS windows: BlockSolverX (run-time block widths: key-frames and motions 6, joints 3, bone lengths 1 stay in the
    reduced system with their couplings from the rigidity / motion / joint edges, map points are marginalised; src/Optimizer.cc:1508-1516)
    on three articulated windows, reduced order up to 780."""
    g = _gen()
    gold = np.load(GOLD)
    for i, case in enumerate(g.SCHUR_X_CASES):
        s = g.open_session_x(oracle_mod, i)
        sysd = s.system_x()
        assert list(gold[f"x{i}_sizes"]) == [len(sysd["dims"]), sysd["n_points"], len(sysd["edge_block"]), sysd["n_dense"]], i
        assert set(np.unique(sysd["dims"])) == {1, 3, 6}
        for k in range(len(case[6])):
            lam = float(gold[f"x{i}_{k}_lambda"])
            assert lam == g.first_lambda_x(sysd) * case[6][k]
            ok, x = s.solve(lam)
            xr, bs = gold[f"x{i}_{k}_x"], gold[f"x{i}_{k}_bschur"]
            assert ok and x.shape == xr.shape
            assert np.abs(x - xr).max() <= 1e-12 * np.abs(xr).max(), (i, k, float(np.abs(x - xr).max() / np.abs(xr).max()))   # measured <= 1e-14
            if f"x{i}_{k}_hschur_upper" in gold.files:
                n = len(bs)
                hs = np.zeros((n, n)); hs[np.triu_indices(n)] = gold[f"x{i}_{k}_hschur_upper"]
                hs = hs + np.triu(hs, 1).T
                assert np.abs(hs @ x[:n] - bs).max() <= 1e-12 * np.abs(bs).max(), (i, k)
        s.close()


@pytest.mark.skipif(not HAVE_REF, reason="needs oracle/_ref/libref_schur.so built from /root/reference (build container only)")
def test_fixture_is_what_the_reference_code_computes_now(oracle_mod):
    """Live: the reference's solve() compiled here reproduces the committed fixture bit for bit (same compiler, same inputs)."""
    g = _gen()
    gold = np.load(GOLD)
    lib = C.CDLL(REF_LIB)
    for i, case in enumerate(g.SCHUR_CASES):
        s = g.open_session(oracle_mod, i)
        sysd = s.system()
        for k in range(len(case[7])):
            ok, x, hs, bs = oracle_mod.ref_schur_solve(lib, sysd, float(gold[f"c{i}_{k}_lambda"]))
            assert ok and (x == gold[f"c{i}_{k}_x"]).all() and (bs == gold[f"c{i}_{k}_bschur"]).all() and (hs == gold[f"c{i}_{k}_hschur"]).all(), (i, k)
        s.close()
    for i, case in enumerate(g.SCHUR_X_CASES):
        s = g.open_session_x(oracle_mod, i)
        sysd = s.system_x()
        for k in range(len(case[6])):
            ok, x, hs, bs = oracle_mod.ref_schur_solve_x(lib, sysd, float(gold[f"x{i}_{k}_lambda"]))
            assert ok and (x == gold[f"x{i}_{k}_x"]).all() and (bs == gold[f"x{i}_{k}_bschur"]).all(), (i, k)
        s.close()
