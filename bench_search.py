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
    # e2e = the C-ABI call with host arrays (struct array filled once, as the C++ shim fills it with pointer assignments):
    # pack into pinned scratch + H2D + kernels + D2H inside the timed region
    prep = m.prepare_projection(probs)
    dev_ms, wall = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        res = m.run_prepared_projection(prep)
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
    try:
        out["best2"] = run_best2(m, device, steps)
    except Exception as e:   # noqa: BLE001
        out["best2"] = {"error": repr(e)}
    m.close()
    return out


def run_best2(m, device: int = 0, steps: int = 5, nq: int = 65536, nt: int = 2048):
    """ORBmatcher best / second-best scan (src/ORBmatcher.cc:85-114) with DescriptorDistance (:1647-1663) as an all-pairs N x M problem
    through adb_match_best2_device: pairs / s against the POPC peak SURVEY.md 8(d) names (SMs x 16 popc / clk x clock / 8 words per pair)
    and bytes / s against HBM ((N + M) x 32 + N x 12 algorithmic bytes)."""
    import ctypes as C
    import json
    import os
    import torch
    from airdos_b200.capi import check, lib
    dev = torch.device("cuda", device)
    g = torch.Generator(device=dev); g.manual_seed(1234)
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)
    t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device=dev, generator=g)
    t[::97] = q[:len(t[::97])]                                   # some exact matches
    o = torch.empty((3, nq), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(dev)          # a real stream handle: a NULL stream argument means "the matcher's own stream" to the C-ABI
    st.wait_stream(torch.cuda.current_stream(dev))

    def call():
        check(lib().adb_match_best2_device(m._m, q.data_ptr(), nq, t.data_ptr(), nt, None, None, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                           C.c_void_p(st.cuda_stream)))
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        call()
    e1.record(st)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    torch.cuda.current_stream(dev).wait_stream(st)
    # parity on a sample against a brute force in torch (bit counting by table)
    pop = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int16, device=dev)
    sample = torch.arange(0, nq, 509, device=dev)
    dist = pop[(q[sample][:, None, :] ^ t[None, :, :]).long()].sum(-1)            # [S, nt]
    srt, idx = torch.sort(dist, dim=1, stable=True)
    ok = bool(torch.equal(o[0][sample].long(), idx[:, 0]) and torch.equal(o[1][sample].long(), srt[:, 0].long()) and torch.equal(o[2][sample].long(), srt[:, 1].long()))
    props = torch.cuda.get_device_properties(dev)
    clk = 1.965e9
    peak_pairs = props.multi_processor_count * 16 * clk / 8
    pairs = nq * nt / (ms * 1e-3)
    hbm = 6650.0
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            hbm = float(json.load(open(pk))["hbm_gbs"])
        except Exception:   # noqa: BLE001
            pass
    alg = (nq + nt) * 32 + nq * 12
    return {"metric": "hamming_pairs_per_s", "value": pairs, "unit": "pairs/s", "ms_per_call": ms,
            "config": {"workload": f"all-pairs best / second-best Hamming, {nq} queries x {nt} descriptors of 256 bit"},
            "roofline": {"bound": "integer pipe (POPC)", "achieved": pairs, "peak": peak_pairs, "unit": "pairs/s", "frac": pairs / peak_pairs,
                         "peak_model": "SMs x 16 POPC / clk / SM x 1.965 GHz / 8 words per pair (SURVEY.md 8d)",
                         "hbm_achieved_GBps": alg / (ms * 1e-3) / 1e9, "hbm_frac": alg / (ms * 1e-3) / 1e9 / hbm},
            "parity": {"best_second_equal_on_sample": ok, "sample": int(len(sample))}}


if __name__ == "__main__":
    import json
    print(json.dumps(run()))
