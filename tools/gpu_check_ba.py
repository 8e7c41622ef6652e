"""CUDA BA vs oracle (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from airdos_b200 import synth, ba

oracle.build()
opt = ba.Optimizer()
ok = True
for name, kw in [("small", dict(n_kf=10, n_points=800, seed=1)), ("mono_mix", dict(n_kf=8, n_points=500, seed=2, mono_frac=0.3)),
                 ("fixed_extra", dict(n_kf=12, n_points=1500, seed=3, n_fixed_extra=4)), ("cfg4", dict(n_kf=50, n_points=20000, seed=4000)),
                 ("dyn_small", dict(n_kf=12, n_points=600, seed=9, humans=2)), ("cfg5", dict(n_kf=80, n_points=30000, seed=5000, humans=4))]:
    t = time.time(); d = synth.make_ba_problem(**kw); tg = time.time() - t
    t = time.time(); po, ro, so = oracle.ba_solve(d); to = time.time() - t
    t = time.time(); pg, rg, sg = opt.LocalBundleAdjustment(d); tg2 = time.time() - t
    t = time.time(); pg, rg, sg = opt.LocalBundleAdjustment(d); tg3 = time.time() - t
    dt = np.abs(pg["pose_t"] - po["pose_t"]).max(); dq = np.abs(pg["pose_q"] - po["pose_q"]).max(); dx = np.abs(pg["points"] - po["points"]).max()
    same_out = bool((rg.edge_outlier == ro.edge_outlier).all())
    ntr = min(len(rg.trace_rows), len(ro.trace_rows))
    trel = np.abs(rg.trace_rows[:ntr, :3] - ro.trace_rows[:ntr, :3]) / np.maximum(np.abs(ro.trace_rows[:ntr, :3]), 1e-30)
    print(f"{name}: E={len(d['edge_pose'])} gen {tg:.1f}s oracle {to*1e3:.0f} ms gpu {tg2*1e3:.0f}/{tg3*1e3:.0f} ms  its {list(rg.c.iterations_run)} vs {list(ro.c.iterations_run)} "
          f"trials {rg.c.trials_run}/{ro.c.trials_run} |dt|={dt:.2e} |dq|={dq:.2e} |dX|={dx:.2e} outliers_equal={same_out} ({rg.edge_outlier.sum()}) "
          f"trace_rel_max={trel.max():.2e} chi {rg.c.chi2_round[1]:.6f} vs {ro.c.chi2_round[1]:.6f}")
    if "joints" in d:
        print("    dyn: |dJ|=%.2e |dD|=%.2e |dmt|=%.2e flags equal %s %s %s" % (np.abs(pg["joints"] - po["joints"]).max(), np.abs(pg["dists"] - po["dists"]).max(),
              np.abs(pg["motion_t"] - po["motion_t"]).max(), (rg.jedge_outlier == ro.jedge_outlier).all(), (rg.redge_outlier == ro.redge_outlier).all(),
              (rg.medge_outlier == ro.medge_outlier).all()))
    print("    stage ms", {k: round(v, 3) for k, v in opt.stage_ms().items()}, "launches", opt.launch_count())
    ok &= dt < 1e-4 and same_out and len(rg.trace_rows) == len(ro.trace_rows)
print("ALL OK" if ok else "MISMATCH")
