"""Pins the Levenberg-Marquardt control, the Huber kernel and the per-edge quadratic form of oracle/ba_oracle.cpp to the LITERAL reference:
tests/golden/lm_ref.npz holds what the reference's own functions compute (oracle/ref_lm.cpp compiles, from /root/reference,
Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp as a whole, SparseOptimizer::optimize (core/sparse_optimizer.cpp:354-419),
RobustKernelHuber::setDelta / robustify (core/robust_kernel_impl.cpp:65-92), BaseEdge::chi2 / robustInformation (core/base_edge.h:58-61,
96-102) and BaseBinaryEdge / BaseUnaryEdge::constructQuadraticForm (core/base_binary_edge.hpp:54-117, base_unary_edge.hpp:40-69);
oracle/gen_ref_lm_golden.py wrote the fixture).

The LM fixture is the reference's control flow run over the oracle's arithmetic (its computeActiveErrors / buildSystem / solve / update /
push / pop calls land in the oracle's steps through function pointers): the oracle's own loop must reproduce every trial -- lambda, chi2
before and after, accept / reject -- the final lambda, the number of iterations and the final state BIT FOR BIT, on cases with accepted
steps only, rejected trials (lambda *= ni, ni *= 2), the maxTrialsAfterFailure exit and the "_nBad >= 3" stop."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "lm_ref.npz")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_lm.so")
HAVE_REF = os.path.exists(REF_LIB) and os.path.isdir("/root/reference")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_ref_lm_golden", os.path.join(ROOT, "oracle", "gen_ref_lm_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    return g


def test_oracle_lm_loop_equals_the_reference_control_flow(oracle_mod):
    g = _gen()
    gold = np.load(GOLD)
    kinds = set()
    for i in range(len(g.LM_CASES)):
        d, its, robust, mt = g.make_lm_case(i)
        o = oracle_mod.ba_default_options(); o.max_trials = mt
        s = oracle_mod.LmSession(d, o, robust)
        it, rows, lam = s.optimize(its)
        meta = gold[f"lm{i}_meta"]
        assert it == int(meta[0]) and lam == meta[1], i
        assert rows.shape == gold[f"lm{i}_rows"].shape and (rows == gold[f"lm{i}_rows"]).all(), i
        assert (s.state() == gold[f"lm{i}_state"]).all(), i
        assert meta[3] == o.tau == 1e-5                      # the constructor's _tau (optimization_algorithm_levenberg.cpp:46)
        assert int(meta[2]) == it + len(rows)                # one computeActiveErrors per iteration + one per trial
        s.close()
        if (rows[:, 3] == 0).any(): kinds.add("rejected")
        if it < its: kinds.add("early stop")
    assert kinds == {"rejected", "early stop"}


def test_oracle_huber_equals_the_reference_kernel(oracle_mod):
    """rho(e2), rho'(e2) bit for bit, including squared errors between the double delta^2 and the reference's FLOAT dsqr member
    (core/robust_kernel_impl.h:84), where a double-precision threshold would classify the edge differently."""
    gold = np.load(GOLD)
    hd, he, rho = gold["huber_delta"], gold["huber_e2"], gold["huber_rho"]
    got = np.array([oracle_mod.huber(a, b) for a, b in zip(hd, he)])
    assert (got == rho[:, :2]).all()
    between = [(a, b) for a, b in zip(hd, he) if min(a * a, float(np.float32(a * a))) < b <= max(a * a, float(np.float32(a * a)))]
    assert len(between) >= 6                                 # the sweep does contain such values
    outl = rho[:, 1] < 1
    assert outl.any() and (~outl).any()


def test_oracle_quadratic_forms_equal_the_reference_edges(oracle_mod):
    """hl / gl / hp / gp / W of one reprojection edge and h / g of one OnlyPose edge against constructQuadraticForm of the reference's
    BaseBinaryEdge / BaseUnaryEdge on 600 seeded edges (mono / stereo, robust on / off, fixed pose): 1e-12 relative (the Eigen stand-in
    multiplies in Eigen's order; a real Eigen build may differ in the last bit)."""
    g = _gen()
    gold = np.load(GOLD)
    for t in range(g.N_QF):
        dim, Ji, Jj, er, w0, delta, robust, fixed = g.make_qf_case(t)
        q = oracle_mod.edge_quadratic_form(dim, Ji, Jj, er, w0, delta, robust, fixed)
        u = oracle_mod.pose_quadratic_form(dim, Jj, er, w0, delta, robust)
        got = np.concatenate(list(q) + list(u)); ref = gold[f"qf{t}"]
        o = 0
        for x in list(q) + list(u):
            r = ref[o:o + len(x)]; o += len(x)
            assert np.abs(x - r).max() <= 1e-12 * max(np.abs(r).max(), 1e-300), t
        assert len(got) == len(ref)


def test_oracle_multi_edge_quadratic_forms_equal_the_reference(oracle_mod):
    """BaseMultiEdge::constructQuadraticForm / computeQuadraticForm (core/base_multi_edge.hpp:35-49, 170-222) of the reference on 300 seeded
    rigidity (1 x [3 | 3 | 1]) and motion (3 x [3 | 3 | 6]) edges, kernel on / off: the diagonal blocks the three vertices receive, the
    three off-diagonal blocks and the right-hand sides against what Solver::build_system accumulates for such an edge, 1e-12."""
    g = _gen()
    gold = np.load(GOLD)
    for t in range(g.N_MULTI):
        H, b = oracle_mod.multi_quadratic_form(*g.make_multi_case(t))
        ref = gold[f"mq{t}"]
        got = np.concatenate([H.ravel(), b])
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), t


@pytest.mark.skipif(not HAVE_REF, reason="reference tree / oracle/_ref not present (GPU box)")
def test_fixture_is_what_the_reference_library_computes_now(oracle_mod):
    import ctypes as C
    g = _gen()
    gold = np.load(GOLD)
    L = C.CDLL(REF_LIB)
    for i in (3, 5):
        d, its, robust, mt = g.make_lm_case(i)
        o = oracle_mod.ba_default_options(); o.max_trials = mt
        s = oracle_mod.LmSession(d, o, robust)
        it, rows, lam, n_eval, tau = oracle_mod.ref_lm_optimize(L, s, its)
        assert it == int(gold[f"lm{i}_meta"][0]) and (rows == gold[f"lm{i}_rows"]).all() and (s.state() == gold[f"lm{i}_state"]).all()
        s.close()
    for k in (0, 100, 400, 845):
        assert oracle_mod.huber(float(gold["huber_delta"][k]), float(gold["huber_e2"][k]), lib=L) == tuple(gold["huber_rho"][k][:2])
