"""bench_search.py -- tracking-search section of bench.py: ORBmatcher::SearchByProjection(CurrentFrame,
LastFrame, th) for a stream of frames, 2000 key-points and 1500 last-frame map points each, grid build + candidate
enumeration + Hamming + closure fixed point + rotation check on the device; metric = map-point queries / s."""
from __future__ import annotations

import time

import numpy as np

import airdos_b200 as adb
from airdos_b200 import synth


def run(device: int = 0, steps: int = 5, frames: int = 128, with_cpu: bool = True):
    base = [synth.make_tracking_problem(300 + i, n_kp=2000, n_q=1500, dup_frac=0.1) for i in range(8)]
    probs = [base[i % len(base)] for i in range(frames)]
    nq = sum(len(p["q_flags"]) for p in probs)
    m = adb.ORBmatcher(0.9, True, device)
    for _ in range(3):
        res = m.search_by_projection(probs)
    dev_ms, wall = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        res = m.search_by_projection(probs)
        wall.append(time.perf_counter() - t0)
        dev_ms.append(m.search_last_ms())
    one = []
    for i in range(20):
        t0 = time.perf_counter()
        m.SearchByProjection(probs[i % frames])
        one.append(time.perf_counter() - t0)
    d, w = float(np.median(dev_ms)), float(np.median(wall))
    out = {"metric": "search_by_projection_queries_per_s", "unit": "queries/s", "dtype": "u8/f32",
           "value": nq / (d * 1e-3), "ms_per_step": d,
           "e2e": {"value": nq / w, "unit": "queries/s", "ms_per_step": w * 1e3},
           "config": {"workload": f"SearchByProjection(Current, Last): {frames} frames x (2000 key-points, 1500 map points), th 7, rotation check",
                      "matches_per_frame": float(np.mean([r[0] for r in res])), "single_frame_latency_ms_host_api": float(np.median(one)) * 1e3},
           "gpu_launches_per_step": 2}
    if with_cpu:
        import oracle
        oracle.build()
        t0 = time.perf_counter()
        for p in base:
            ref = oracle.search_by_projection(p)
        t = time.perf_counter() - t0
        g = m.SearchByProjection(base[-1])
        out["cpu_baseline"] = {"value": sum(len(p["q_flags"]) for p in base) / t, "unit": "queries/s", "cores": 1, "kind": "port",
                               "sample": f"{len(base)} of the frames, oracle port on 1 host thread (the reference searches on the tracking thread)"}
        out["parity"] = {"kp_match_equal": bool((g[1] == ref[1]).all()), "nmatches_equal": g[0] == ref[0]}
    # vocabulary-bucket searches (SearchByBoW / SearchForTriangulation), 64 key-frame pairs per call
    bow = {}
    for mode, name in ((0, "SearchByBoW"), (1, "SearchForTriangulation")):
        bprobs = [synth.make_bow_problem(700 + i, mode) for i in range(4)] * 16
        nqb = sum(len(p["b_idx1"]) for p in bprobs)
        for _ in range(2):
            m.search_by_bow(bprobs)
        ms = []
        for _ in range(steps):
            m.search_by_bow(bprobs)
            ms.append(m.search_last_ms())
        bow[name] = {"queries_per_s": nqb / (float(np.median(ms)) * 1e-3), "ms_per_64_pairs": float(np.median(ms))}
        if with_cpu:
            t0 = time.perf_counter()
            for p in bprobs[:4]:
                oracle.search_by_bow(p)
            bow[name]["cpu_port_queries_per_s"] = sum(len(p["b_idx1"]) for p in bprobs[:4]) / (time.perf_counter() - t0)
    out["bow"] = bow
    m.close()
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(run()))
