"""GPU tests of the in-cluster FP64 Cholesky solve (airdos_b200/csrc/chol.cu) that replaces g2o's LinearSolverDense /
LinearSolverEigen (Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:64-113, linear_solver_eigen.h:92-115) under adb_ba_solve.
Checked against numpy's LAPACK solve in FP64: backward error at rounding level, not-positive-definite input flagged."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spd(n, seed, cond=1e4):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    ev = np.geomspace(1.0, cond, n)
    a = (q * ev) @ q.T
    return (a + a.T) / 2, rng.normal(size=n)


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 64, 97, 294, 500, 1226])
@pytest.mark.parametrize("cluster", [8, 16])
def test_dense_solve_matches_lapack(n, cluster):
    from airdos_b200 import ba
    a, b = _spd(n, n)
    x, info, _ = ba.dense_solve(a, b, cluster=cluster)
    assert info == 0
    ref = np.linalg.solve(a, b)
    # backward error: ||A x - b|| / (||A|| ||x||) at rounding level; forward error bounded by cond * eps
    assert np.linalg.norm(a @ x - b) / (np.linalg.norm(a, 2) * np.linalg.norm(x)) < 1e-14
    assert np.abs(x - ref).max() / np.abs(ref).max() < 1e-9


def test_only_the_lower_triangle_is_read():
    from airdos_b200 import ba
    a, b = _spd(200, 7)
    junk = np.tril(a) + np.triu(np.full_like(a, np.nan), 1)
    x, info, _ = ba.dense_solve(junk, b)
    assert info == 0 and np.allclose(x, np.linalg.solve(a, b), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("n,bad", [(40, 3), (294, 100), (294, 293), (700, 650)])
def test_not_positive_definite_is_flagged(n, bad):
    """g2o: `if (!ldlt.isPositive()) return false` -> LM rejects the trial.  info reports the panel of the failing pivot."""
    from airdos_b200 import ba
    a, b = _spd(n, 11)
    a[bad, bad] = -1.0
    _, info, _ = ba.dense_solve(a, b)
    assert info == (bad // 32) * 32 + 1
    a2, _ = _spd(n, 12)
    a2[bad, :] = np.nan; a2[:, bad] = np.nan
    _, info, _ = ba.dense_solve(a2, b)
    assert info != 0


def test_repeated_solves_are_deterministic():
    from airdos_b200 import ba
    a, b = _spd(294, 3)
    x0, _, _ = ba.dense_solve(a, b, reps=3)
    x1, _, _ = ba.dense_solve(a, b, reps=2)
    assert (x0 == x1).all()
