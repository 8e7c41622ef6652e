"""bench_ba.py -- bundle-adjustment sections of bench.py.

`ba`          BASELINE.json configs[3]: LocalBundleAdjustment, 50 KF poses, 20k MapPoints, 120k reprojection edges, 10 LM iterations.
`ba_dynamic`  BASELINE.json configs[4]: LocalBundleAdjustmentHumanTrajactory (src/Optimizer.cc:1496-2222), 80 KF, 30k points, 180k stereo
              edges + 16 MapHumanPose skeletons (224 joint / 224 rigidity / 60 motion edges), dense reduced system of order 1226.
Metric = edges / s per LM iteration on one B200 (device events around the LM loop), e2e = wall clock around adb_ba_solve with host arrays,
CPU baseline = oracle port on one host thread (g2o's OpenMP is off by default, Thirdparty/g2o/CMakeLists.txt:48)."""
from __future__ import annotations

import json
import os
import time

import numpy as np

from airdos_b200 import ba, synth

ROOT = os.path.dirname(os.path.abspath(__file__))
# FP64 tensor pipe (DMMA m8n8k4) peak measured on this pool's B200 with tools/probe/fp64_probe.cu: 0.25 warp-DMMA / clk / SM
# = 128 FLOP / clk / SM; at the 1965 MHz boost clock 251.5 GFLOP/s per SM, 37.2 TFLOP/s for 148 SMs (NVIDIA quotes 40).
DMMA_GFLOPS_PER_SM = 251.5


def _opts(its0, its1):
    o = ba.default_options()
    o.iterations[0] = its0; o.iterations[1] = its1
    return o


def _hbm_peak():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            return float(json.load(open(pk))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def run(device: int = 0, steps: int = 5, with_cpu: bool = True, n_kf: int = 50, n_points: int = 20000, seed: int = 4000, humans: int = 0,
        human_poses: int = 4, cpu_full_schedule: bool = True):
    d = synth.make_ba_problem(n_kf, n_points, 6, seed=seed, humans=humans, human_poses=human_poses)
    n_dyn = sum(len(d[k]) for k in ("jedge_pose", "redge_i", "medge_p1") if k in d)
    E = len(d["edge_pose"]) + n_dyn
    opt = ba.Optimizer(device)
    o10 = _opts(10, 0)
    for _ in range(3):
        opt.LocalBundleAdjustment(d, options=o10)                      # warm-up (buffers, module load)
    l0 = opt.launch_count()
    lm_ms, wall, trials = [], [], 0
    stages = {}
    for _ in range(steps):
        t0 = time.perf_counter()
        p, r, st = opt.LocalBundleAdjustment(d, options=o10)
        wall.append(time.perf_counter() - t0)
        s = opt.stage_ms()
        lm_ms.append(s["lm_loop"])
        trials = r.c.trials_run
        for k, v in s.items():
            stages[k] = stages.get(k, 0.0) + v / steps
    launches = (opt.launch_count() - l0) // steps
    lm = float(np.median(lm_ms)); wl = float(np.median(wall))
    n_free = n_kf - 1
    nd = 6 * n_free + (3 * len(d["joints"]) + len(d["dists"]) + 6 * len(d["motion_t"]) if humans else 0)
    name = "LocalBundleAdjustmentHumanTrajactory" if humans else "LocalBundleAdjustment"
    extra = (f" + {humans * human_poses} MapHumanPose skeletons ({len(d['jedge_pose'])} joint / {len(d['redge_i'])} rigidity / {len(d['medge_p1'])} motion edges)"
             if humans else "")
    n_state = n_points * 24 + n_kf * 56 + (len(d["joints"]) * 24 + len(d["dists"]) * 8 + len(d["motion_t"]) * 56 if humans else 0)
    out = {
        "metric": "ba_edges_per_s_per_lm_iteration", "unit": "edges/s", "dtype": "f64",
        "value": E * trials / (lm * 1e-3), "ms_per_lm_iteration": lm / trials,
        "e2e": {"value": E * trials / wl, "unit": "edges/s", "ms_per_solve": wl * 1e3,
                "h2d_bytes_per_step": int(E * (8 + 24 + 8) + n_state), "d2h_bytes_per_step": int(n_state + E * 9)},
        "config": {"workload": f"{name}: {n_kf} KF poses, {n_points} MapPoints, {len(d['edge_pose'])} reprojection edges{extra}, 10 LM iterations (robust), 1 B200",
                   "lm_trials": trials, "reduced_dim": nd},
        "stage_ms_per_solve": {k: round(v, 4) for k, v in stages.items()}, "gpu_launches_per_solve": int(launches),
    }
    # roofline of the HBM-bound part (520 B / edge + 168 B / point + ~1 KB / pose + 288 B per reduced 6x6 block, per LM trial)
    alg = (520 * E + 168 * n_points + 1024 * n_kf + 8 * nd * nd) * trials
    sparse_ms = stages["linearize"] + stages["schur"] + stages["backsub_eval"]
    peak, src = _hbm_peak()
    solve_ms = stages["reduced_solve"] / trials
    cluster = 16 if (nd + 31) // 32 >= 16 else 8
    gf = nd ** 3 / 3 / (solve_ms * 1e-3) / 1e9
    out["roofline"] = {"bound": "hbm", "kernel": "ba_point/pose_linearize + ba_schur_block + ba_backsub + ba_eval", "achieved": alg / (sparse_ms * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": alg / (sparse_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": src,
                       "note": "L2-resident after the first trial (DRAM traffic per trial is a fraction of the algorithmic bytes: profiles/r2c_ncu_full_summary.csv, ba_* rows): latency / L2 bound, not HBM bound",
                       "reduced_solve": {"kernel": "chol_cluster_kernel (one launch, one cluster)", "n": nd, "cluster_ctas": cluster, "ms_per_trial": solve_ms,
                                         "bound": "tensor (FP64 DMMA m8n8k4)", "achieved": gf, "unit": "GFLOP/s",
                                         "peak_cluster": DMMA_GFLOPS_PER_SM * cluster, "frac_of_cluster_peak": gf / (DMMA_GFLOPS_PER_SM * cluster),
                                         "peak_chip": DMMA_GFLOPS_PER_SM * 148, "frac_of_chip_peak": gf / (DMMA_GFLOPS_PER_SM * 148),
                                         "note": "n dependent pivots: the chain of 32x32 diagonal factorisations bounds the solve, not FLOP/s; tensor-pipe "
                                                 "utilisation from ncu in profiles/r2c_ncu_full_summary.csv (chol_cluster_kernel rows: pipe_tensor_* columns); cuSOLVER potrf+potrs on the same box: tools/chol_bench.py, profiles/r2c_chol_vs_cusolver.jsonl"}}
    # parity on the reference's own 5 + 10 schedule, and the CPU baseline (oracle port, 1 thread like g2o without OpenMP)
    if with_cpu:
        import oracle
        oracle.build()
        if cpu_full_schedule:
            pg, rg, _ = opt.LocalBundleAdjustment(d)
            t0 = time.perf_counter()
            po, ro, _ = oracle.ba_solve(d)
            t_full = time.perf_counter() - t0
            sched = "5 + 10 iterations with chi2 gates"
        t0 = time.perf_counter()
        po10, ro10, _ = oracle.ba_solve(d, o10)
        t10 = time.perf_counter() - t0
        if not cpu_full_schedule:
            pg, rg, po, ro, t_full, sched = p, r, po10, ro10, t10, "10 iterations, one round"
        out["parity"] = {"max_abs_pose_translation_diff": float(np.abs(pg["pose_t"] - po["pose_t"]).max()),
                         "outlier_flags_equal": bool((rg.edge_outlier == ro.edge_outlier).all()), "schedule": sched}
        if humans:
            out["parity"].update(max_abs_joint_diff=float(np.abs(pg["joints"] - po["joints"]).max()),
                                 human_edge_flags_equal=bool((rg.jedge_outlier == ro.jedge_outlier).all() and (rg.redge_outlier == ro.redge_outlier).all()
                                                             and (rg.medge_outlier == ro.medge_outlier).all()))
        out["cpu_baseline"] = {"value": E * ro10.c.trials_run / t10, "unit": "edges/s", "cores": 1, "kind": "port",
                               "sample": f"the same window, 10 LM iterations, oracle port on 1 host thread ({t10 * 1e3:.0f} ms; {sched} {t_full * 1e3:.0f} ms)"}
    opt.close()
    return out


def run_dynamic(device: int = 0, steps: int = 3, with_cpu: bool = True):
    """BASELINE.json configs[4]."""
    return run(device, steps, with_cpu, n_kf=80, n_points=30000, seed=5000, humans=4, human_poses=4, cpu_full_schedule=False)


if __name__ == "__main__":
    import sys
    print(json.dumps(run_dynamic() if "--dynamic" in sys.argv else run()))
