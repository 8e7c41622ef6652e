"""How far does last-bit noise move the result of a bundle-adjustment fixture?  (CPU only: the oracle, no GPU.)

The CUDA solver accumulates the Schur complement and chi2 with FP64 atomics, so the summation order -- and the last bits of every LM
step -- differ from run to run.  This script measures the amplification the LM schedule applies to such noise: it re-solves every window
of tests/golden/lba_ref.npz with the oracle after multiplying the initial points / translations by (1 + k * 2.2e-16), k in {-1, 0, 1},
and prints the largest change of the final state.  The tolerances of tests/test_ba_gpu.py sit >= 30x above these numbers
(static windows: up to 2.5e-9 -> 1e-7; articulated windows: 2.5e-11 -> 1e-7; global BA with a fixed key-frame: 6.5e-11 -> 1e-7;
gauge-free global BA, 20 iterations: up to 3.4e-5 over 60 draws, all of it in a handful of far points -- quaternions 2e-9, translations
1.3e-7 -- hence the per-block bars 1e-6 / 1e-5 / 1e-3 of that test).

usage: PYTHONPATH=. python tools/ba_noise_sensitivity.py"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (test infrastructure; this tool is not part of the product path)

oracle.build()
gold = np.load(os.path.join(ROOT, "tests", "golden", "lba_ref.npz"))
rng = np.random.default_rng(0)
STATIC = ("pose_q", "pose_t", "points")
DYNAMIC = STATIC + ("joints", "dists", "motion_q", "motion_t")


def state(p, keys):
    return np.concatenate([np.asarray(p[k]).ravel() for k in keys])


def problem(prefix, i):
    return {k[len(f"{prefix}{i}_p_"):]: (gold[k].item() if gold[k].ndim == 0 else gold[k]) for k in gold.files if k.startswith(f"{prefix}{i}_p_")}


def worst_change(prob, keys, options=None, trials=6):
    base = state(oracle.ba_solve(prob, options)[0], keys)
    worst = 0.0
    for _ in range(trials):
        q = dict(prob)
        for k in ("points", "pose_t"):
            a = np.array(q[k], np.float64)
            q[k] = a * (1 + rng.integers(-1, 2, a.shape) * 2.2e-16)
        worst = max(worst, float(np.abs(state(oracle.ba_solve(q, options)[0], keys) - base).max()))
    return worst


if __name__ == "__main__":
    for prefix, keys in (("w", STATIC), ("h", DYNAMIC)):
        i = 0
        while f"{prefix}{i}_rows" in gold.files:
            print(f"{prefix}{i}: worst |d state| under 1-ulp input noise = {worst_change(problem(prefix, i), keys):.3e}")
            i += 1
    spec = importlib.util.spec_from_file_location("gen_ref_lba_golden", os.path.join(ROOT, "oracle", "gen_ref_lba_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    for j, (c, its, loop_kf, robust) in enumerate(g.GBA_CASES):
        prob = problem("g", j)
        print(f"g{j} ({'fixed key-frame' if prob['pose_fixed'].any() else 'gauge-free'}, {its} iterations): "
              f"{worst_change(prob, STATIC, oracle.ba_global_options(its, robust)):.3e}")
