#!/usr/bin/env python
"""bench.py -- headline benchmark of the AirDOS hot path on B200 (contract: see README / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement, all host threads

Workload (BASELINE.json configs[1]): a stream of 640x480 stereo pairs, 8-level pyramid, 2000
features per frame; one step = ORB extraction of the left and right images of `pairs` stereo
pairs + Frame::ComputeStereoMatches on every pair.  Metric: output key-points per second.
At N > 1 every rank processes its own `pairs` pairs (weak scaling) and one NCCL all-gather of the
fixed-stride (key-point, descriptor, count) records per step makes every rank hold all results
(BASELINE.json configs[2]).  A second section (`ba`) reports the LocalBundleAdjustment metric
(edges/s per LM iteration, configs[3]) when the BA solver is built.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT, NLEVELS, SCALE, INI_TH, MIN_TH = 640, 480, 2000, 8, 1.2, 12, 7
# algorithmic bytes per 640x480 frame at 2000 features
PYR_PX = 950532
BYTES_PER_FRAME = 307200 + PYR_PX + PYR_PX + PYR_PX + 2000 * (32 + 24)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# the committed `ncu --set full` capture of the UNMASKED extractor kernels at 128 frames per launch (profiles/README.md); the r2j_masked
# capture holds the masked instances of the same kernel names and must not be picked up here
TRAFFIC_CAPTURE = "r2c_dram_traffic.json"


def dram_traffic(kernel: str, frames: int):
    """dram__bytes_read + dram__bytes_write of `kernel` per launch of `frames` frames, scaled from the committed ncu capture
    (the single-tile + the wide-cell instance of one FAST call) -> (bytes or None, source file or None)."""
    path = os.path.join(ROOT, "profiles", TRAFFIC_CAPTURE)
    try:
        tj = json.load(open(path))
        if kernel in tj and not kernel.startswith("pyr_resize"):
            return float(np.sum([e["dram_bytes"] for e in tj[kernel][:2]])) / 128.0 * frames, "profiles/" + TRAFFIC_CAPTURE
    except Exception:
        pass
    return None, None


ISSUE_CAPTURE = "r2c_ncu_full_summary.csv"


def issue_roofline(kernel: str, frames: int, kernel_ms: float, sm_mhz, n_sms: int = 148):
    """The bound that actually holds for the FAST kernel ("integer ALU is the realistic limiter for FAST -- report
    both"): executed warp-instructions per second against the SMs' issue rate (4 schedulers x 1 warp-instruction per clock each).
    The instruction count comes from the committed ncu capture (smsp__inst_executed.sum of the single-tile + the wide-cell instance at
    128 frames per launch, profiles/r2c_ncu_full_summary.csv), the time is this run's.  -> dict, or None when the capture is missing."""
    try:
        import csv
        rows = list(csv.reader(l for l in open(os.path.join(ROOT, "profiles", ISSUE_CAPTURE)) if not l.startswith("#")))
        h = rows[0]
        ki, wi, ai = h.index("kernel"), h.index("warp_instructions"), h.index("pipe_alu_pct")
        mine = [r for r in rows[1:] if r[ki] == kernel][:2]
        if not mine or not kernel_ms or kernel_ms <= 0:
            return None
        inst = sum(float(r[wi]) for r in mine) / 128.0 * frames
        clk = float(sm_mhz) * 1e6 if sm_mhz else 1965e6
        peak = n_sms * 4 * clk
        return {"bound": "issue", "achieved": inst / (kernel_ms * 1e-3) / 1e9, "peak": peak / 1e9, "unit": "G warp-instructions/s",
                "frac": inst / (kernel_ms * 1e-3) / peak, "warp_instructions_per_launch": inst,
                "alu_pipe_pct_in_capture": [float(r[ai]) for r in mine], "source": "profiles/" + ISSUE_CAPTURE,
                "note": "instruction count from the committed ncu capture (128 frames per launch), time and clock from this run"}
    except Exception:
        return None


def cpu_reference_run(pairs_arr, threads, steps, warmup):
    """Reference CPU path (oracle port) on the host cores: kp/s over `steps` passes of the sample.  The SAME routine serves the
    `cpu_baseline` leg of our arm and `--impl reference`: at least one full warm pass (thread pool, page faults, clocks), then
    >= 3 timed passes, so that the two numbers agree on the same box."""
    import oracle
    from airdos_b200 import synth
    oracle.build()
    mbf = synth.BF; mb = mbf / synth.FX
    for _ in range(max(warmup, 1)):
        oracle.stereo_pipeline_batch(pairs_arr, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, mb, mbf, threads)
    steps = max(steps, 3)
    t0 = time.perf_counter()
    tot = 0
    for _ in range(steps):
        _, _, n = oracle.stereo_pipeline_batch(pairs_arr, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, mb, mbf, threads)
        tot += n
    dt = time.perf_counter() - t0
    return tot / dt, dt / steps * 1e3, tot // max(steps, 1)


def cv2_primitive_baseline(n_frames: int = 4):
    """Secondary CPU number: the OpenCV calls the reference's extractor makes per frame -- resize +
    copyMakeBorder per level (src/ORBextractor.cc:1139-1152), FastFeatureDetector per 30-px cell with the ini / min rule
    (:812-824), GaussianBlur 7x7 per level (:1100) -- through cv2 (OpenCV's hand-written SIMD) on ONE thread, timing only the cv2
    calls.  The reference does all of this plus its scalar quad-tree / IC_Angle / rBRIEF, so frames / s / core <= 1 / this."""
    try:
        import cv2
    except Exception as e:   # noqa: BLE001
        return {"unavailable": repr(e)}
    from airdos_b200 import synth
    cv2.setNumThreads(1)
    E = 19
    tsum = 0.0
    ncalls = 0
    sizes = [(W, H)]
    sc = np.float32(1.0)
    for _ in range(1, NLEVELS):
        sc = np.float32(sc * np.float32(SCALE))
        sizes.append((int(np.rint(np.float32(W) * (np.float32(1.0) / sc))), int(np.rint(np.float32(H) * (np.float32(1.0) / sc)))))
    det_ini = cv2.FastFeatureDetector_create(INI_TH, True); det_min = cv2.FastFeatureDetector_create(MIN_TH, True)
    tiny = np.zeros((8, 8), np.uint8)          # cost of one Python -> cv2 detect round trip with nothing to do: subtracted per call below
    for _ in range(200):
        det_ini.detect(tiny, None)
    t0 = time.perf_counter()
    for _ in range(2000):
        det_ini.detect(tiny, None)
    call_overhead = (time.perf_counter() - t0) / 2000
    for f in range(n_frames + 1):
        img = synth.make_stereo_pair(900 + f, W, H)[0]
        t_frame = 0.0
        pyr = []
        for l in range(NLEVELS):
            t0 = time.perf_counter()
            if l == 0:
                full = cv2.copyMakeBorder(img, E, E, E, E, cv2.BORDER_REFLECT_101)
            else:
                r = cv2.resize(pyr[l - 1][E:-E, E:-E], sizes[l], interpolation=cv2.INTER_LINEAR)
                full = cv2.copyMakeBorder(r, E, E, E, E, cv2.BORDER_REFLECT_101)
            t_frame += time.perf_counter() - t0
            pyr.append(full)
        for l in range(NLEVELS):
            roi = pyr[l][E:-E, E:-E]
            lh, lw = roi.shape
            minB, maxBX, maxBY = 16, lw - 16, lh - 16
            ncols, nrows = int(np.float32(maxBX - minB) / np.float32(30)), int(np.float32(maxBY - minB) / np.float32(30))
            wcell, hcell = int(np.ceil(np.float32(maxBX - minB) / ncols)), int(np.ceil(np.float32(maxBY - minB) / nrows))
            for i in range(nrows):
                iniy = minB + i * hcell
                if iniy >= maxBY - 3:
                    continue
                maxy = min(iniy + hcell + 6, maxBY)
                for j in range(ncols):
                    inix = minB + j * wcell
                    if inix >= maxBX - 6:
                        continue
                    sub = roi[iniy:maxy, inix:min(inix + wcell + 6, maxBX)]
                    t0 = time.perf_counter()
                    k = det_ini.detect(sub, None)
                    if len(k) == 0:
                        k = det_min.detect(sub, None)
                    t_frame += time.perf_counter() - t0 - call_overhead * (1 if len(k) else 2)
                    ncalls += 1
            t0 = time.perf_counter()
            cv2.GaussianBlur(roi.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            t_frame += time.perf_counter() - t0
        if f > 0:            # frame 0 warms up
            tsum += t_frame
    ms = tsum / n_frames * 1e3
    return {"ms_per_frame_one_thread": ms, "frames_per_s_per_core_upper_bound": 1e3 / ms, "cv2_version": cv2.__version__,
            "python_call_overhead_us_subtracted_per_detect": call_overhead * 1e6,
            "note": "cv2 (SIMD OpenCV) time of the pyramid + per-cell FAST + 7x7 blur calls only; the reference adds its scalar quad-tree, IC_Angle and "
                    "rBRIEF on top, so its extractor cannot beat 1 / ms_per_frame frames/s per core on this host",
            "frames": n_frames, "fast_calls_per_frame": ncalls // (n_frames + 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from airdos_b200 import synth
    threads = os.cpu_count() or 1
    n_pairs = max(threads, 8)
    pairs = synth.make_stereo_batch(n_pairs)
    v, ms, kp = cpu_reference_run(pairs, threads, args.steps, 1)
    sample = f"{n_pairs} stereo pairs ({2 * n_pairs} frames 640x480) per step, extract L+R + stereo match, {threads} host threads"
    print(json.dumps({
        "impl": "reference", "metric": "orb_keypoints_per_s", "value": v, "unit": "keypoints/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "640x480 stereo stream, 8-level pyramid, 2000 feat/frame, ORB extract L+R + stereo match",
                   "arm": "reference CPU path (oracle port of the reference algorithm, all host threads)",
                   "pairs_per_step": n_pairs},
        "cpu_baseline": {"value": v, "unit": "keypoints/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "keypoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--pairs", type=int, default=3072,
                    help="stereo pairs per step per GPU (3072 pairs = 6144 frames = 54 ms per step: 20 steps give a timed region above one second)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ba", action="store_true")
    ap.add_argument("--gather", default="fused", choices=["fused", "fused_p2p", "nccl"],
                    help="N > 1: 'fused' = descriptor kernel stores records over NVLink into the symmetric buffers of all ranks: once, to the NVLS "
                         "multicast mapping, when the NVSwitch can replicate (else per peer); 'fused_p2p' = per-peer stores always; "
                         "'nccl' = separate ncclAllGather per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import airdos_b200 as adb
    from airdos_b200 import dist as adist, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = "not set"
    try:   # pin this rank (and the pages it first-touches: pinned rings below) to the CPUs next to its GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        numa = "nvmlDeviceSetCpuAffinity: %d cpus" % len(os.sched_getaffinity(0))
    except Exception as e:   # noqa: BLE001
        numa = "unavailable (%s)" % type(e).__name__
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = args.pairs
    K, Wm = args.steps, max(args.warmup, 3)

    # ---- synthetic stream: 16 distinct seeded pairs per rank tiled to P pairs (generation is numpy, slow)
    base = synth.make_stereo_batch(16, W, H, start=100 * rank)
    host = np.concatenate([base] * ((P + 15) // 16))[:P]                      # [P, 2, H, W]
    hostL = torch.from_numpy(np.ascontiguousarray(host[:, 0])).pin_memory()
    hostR = torch.from_numpy(np.ascontiguousarray(host[:, 1])).pin_memory()
    dL, dR = hostL.to(dev), hostR.to(dev)
    exL = adb.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, W, H, max_batch=P, device=local)
    exR = adb.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, W, H, max_batch=P, device=local)
    cap = exL.capacity
    mbf = synth.BF; mb = mbf / synth.FX
    sL = torch.cuda.ExternalStream(exL.stream(), device=dev)
    kp_ptr, desc_ptr, cnt_ptr, _ = exL.results_device()
    kpsL_t = adist.as_tensor(kp_ptr, (P, cap, 24), "|u1", dev)
    descL_t = adist.as_tensor(desc_ptr, (P, cap, 32), "|u1", dev)
    cntL_t = adist.as_tensor(cnt_ptr, (P,), "<i4", dev)
    kp_ptrR, desc_ptrR, cnt_ptrR, _ = exR.results_device()
    cntR_t = adist.as_tensor(cnt_ptrR, (P,), "<i4", dev)
    gather_bufs = None
    gather_mode = "none"
    symm = None
    nvlink_bytes = None
    if world > 1 and args.gather in ("fused", "fused_p2p"):
        try:
            symm = adist.SymmetricGather(world, rank, P, cap, dev)      # rendezvous + peer-mapped buffers
            tk, td, tc, mc = symm.targets(prefer_multicast=args.gather == "fused")
            exL.set_gather(tk, td, tc, mc)
            gather_mode = ("fused: orient_describe_kernel stores every record ONCE to the NVLS multicast mapping of the symmetric buffers (multimem.st; the "
                           "NVSwitch replicates to all ranks)" if mc else
                           "fused: orient_describe_kernel stores every record to all peers over NVLink (symmetric memory, one 32-bit-word store per destination)")
            nvlink_copies = 1 if mc else world - 1
        except Exception as e:   # symmetric memory unavailable: separate collective
            symm = None
            gather_mode = "nccl all_gather_into_tensor (symmetric memory unavailable: %r)" % (e,)
    elif world > 1:
        gather_mode = "nccl all_gather_into_tensor"

    def step():
        exL.extract_batch_device(dL.data_ptr(), P)
        exR.extract_batch_device(dR.data_ptr(), P)
        adb.orb.stereo_match_device(exL, exR, P, mb, mbf)
        if world > 1 and symm is None:
            with torch.cuda.stream(sL):
                return adist.all_gather_records(kpsL_t, descL_t, cntL_t)
        return None

    def sync_all():
        exL.sync(); exR.sync()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    for _ in range(Wm):
        gather_bufs = step()
    sync_all()
    if symm is not None:   # every rank now holds every rank's left-image records: check the counts against the owners'
        allc = [torch.zeros(P, dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(allc, cntL_t.clone())
        got = symm.counts_view()
        for r in range(world):
            assert torch.equal(got[r], allc[r]), "fused gather: counts of rank %d differ" % r
        # ... and the PAYLOAD of every rank: key-point records and descriptors as the owner holds them (valid rows), fetched by NCCL
        alld = [torch.empty_like(descL_t) for _ in range(world)]
        allk = [torch.empty_like(kpsL_t) for _ in range(world)]
        dist.all_gather(alld, descL_t.contiguous()); dist.all_gather(allk, kpsL_t.contiguous())
        rows = torch.arange(cap, device=dev)[None, :]
        for r in range(world):
            valid = rows < allc[r][:, None]                               # [P, cap]
            assert torch.equal(symm.desc_view()[r][valid], alld[r][valid]), "fused gather: descriptors of rank %d differ" % r
            assert torch.equal(symm.kps_view()[r][valid], allk[r][valid]), "fused gather: key-points of rank %d differ" % r
        del alld, allk
        nvlink_bytes = int(cntL_t.sum().item()) * 56 * nvlink_copies
    n_kp_step = int(cntL_t.sum().item() + cntR_t.sum().item())
    launches0 = exL.launch_count() + exR.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(sL)
    for _ in range(K):
        gather_bufs = step()
    ev1.record(sL)
    sync_all()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = exL.launch_count() + exR.launch_count() - launches0
    tmax = torch.tensor([ms_total], device=dev)
    nkp = torch.tensor([float(n_kp_step)], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(nkp, op=dist.ReduceOp.SUM)
    ms_step = float(tmax.item()) / K
    value = float(nkp.item()) / (ms_step * 1e-3)

    # ---- per-stage device times (CUDA events on the handles' own streams), same workload, K steps
    exL.profile(True); exR.profile(True)
    stage = np.zeros(5)
    for _ in range(K):
        exL.extract_batch_device(dL.data_ptr(), P); exL.sync()
        stage += np.array(exL.stage_ms())
    stage /= K
    exL.profile(False); exR.profile(False)
    names = ["pyr_resize_strip_kernel(x7)", "fast_cells_warp_kernel(x2)", "quadtree_kernel", "blur7_level_kernel", "orient_describe_kernel"]
    ncand = 0
    for l in range(NLEVELS):
        ncand += len(exL.debug_candidates(0, l))
    nkp_frame = n_kp_step / (2 * P)
    alg_bytes = [307200 + (PYR_PX - 307200),                      # level 0 read + levels 1..7 written
                 PYR_PX + 4 * ncand + 2 * 815,                    # pyramid read once + candidate records + cell counts
                 4 * ncand * 3 + 4 * nkp_frame,                   # candidates read, key/state scratch, kept list
                 2 * PYR_PX,                                      # pyramid read once, blurred pyramid written once
                 (31 * 31 + 37 * 37) * nkp_frame + 56 * nkp_frame]   # 31x31 level box + 37x37 blurred box per key-point + 32-B descriptor + 24-B record
    dom = int(np.argmax(stage))
    peak, peak_src = peaks()
    achieved = alg_bytes[dom] * P / (stage[dom] * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/), scaled to this launch's frames
    traffic, traffic_src = dram_traffic(names[dom].split("(")[0], P)
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes[dom] * P,
                "note": "the FAST kernel is bound by the integer pipe, not by HBM (ncu, profiles/: ~77 % alu-pipe, ~72 % of the issue slots, "
                        "< 3 % of DRAM bandwidth): the exact score is 39 packed 3-input min / max (VIMNMX3.U16x2, two pipe passes each, "
                        "tools/probe/pipe_probe.cu) + 17 LDS + 17 IMAD per 32 pixels of ~120 warp-instructions in all; the HBM fraction is "
                        "reported because the contract asks for it, DRAM traffic ~= algorithmic bytes (no re-reads)",
                "kernel_ms": float(stage[dom]),
                "issue_roofline": issue_roofline(names[dom].split("(")[0], P, float(stage[dom]), (clocks or {}).get("sm_mhz")),
                "stage_ms": {n: float(s) for n, s in zip(names, stage)},
                "pipeline_achieved_GBps": BYTES_PER_FRAME * 2 * P / (ms_step * 1e-3) / 1e9,
                "pipeline_frac": BYTES_PER_FRAME * 2 * P / (ms_step * 1e-3) / 1e9 / peak}

    # ---- end to end through the host-buffer C-ABI (pinned host images in, key-points/descriptors/matches out)
    outL = adb.orb.HostResults(P, cap, pinned=True); outR = adb.orb.HostResults(P, cap, pinned=True)
    outS = adb.orb.HostStereo(P, cap, pinned=True)
    npL, npR = hostL.numpy(), hostR.numpy()

    def e2e_step():
        # the stereo Frame constructor's hot part (src/Frame.cc:80-100: both extractions, then ComputeStereoMatches) through the one
        # C-ABI call that stands for it: host images in, key-points / descriptors / stereo matches out, chunk-pipelined inside
        adb.stereo_frames_batch(exL, exR, npL, npR, mb, mbf, out_left=outL, out_right=outR, out_stereo=outS)
        return int(outL.counts.sum() + outR.counts.sum())

    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    tot = 0
    for _ in range(K):
        tot += e2e_step()
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    tot_t = torch.tensor([float(tot)], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_t, op=dist.ReduceOp.SUM)
    e2e = {"value": float(tot_t.item()) / float(t_e2e.item()), "unit": "keypoints/s",
           "h2d_bytes_per_step": int(2 * P * W * H),
           "d2h_bytes_per_step": int(2 * (P * cap * (24 + 32) + P * 4) + 4 * P * cap * 4 + P * 4)}

    # ---- masked extraction = the reference's shipped configuration (System.IsMask: 1, Examples/Stereo/config/tartanair.yaml:73;
    #      Frame::ExtractORB passes a person mask with every image, src/Frame.cc:551-571): erosion 10x10 + mask pyramid + masked FAST
    masked = None
    if world == 1:
        mbase = np.stack([np.stack([synth.make_human_mask(7000 + 2 * i + side, W, H, 3) for side in range(2)]) for i in range(16)])
        mhost = np.concatenate([mbase] * ((P + 15) // 16))[:P]
        mLh = torch.from_numpy(np.ascontiguousarray(mhost[:, 0])).pin_memory(); mRh = torch.from_numpy(np.ascontiguousarray(mhost[:, 1])).pin_memory()
        dmL, dmR = mLh.to(dev), mRh.to(dev)

        def mstep():
            exL.extract_batch_device(dL.data_ptr(), P, d_masks=dmL.data_ptr())
            exR.extract_batch_device(dR.data_ptr(), P, d_masks=dmR.data_ptr())
            adb.orb.stereo_match_device(exL, exR, P, mb, mbf)

        for _ in range(3):
            mstep()
        sync_all()
        n_kp_m = int(cntL_t.sum().item() + cntR_t.sum().item())
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(sL)
        for _ in range(K):
            mstep()
        m1.record(sL)
        sync_all()
        ms_m = m0.elapsed_time(m1) / K
        # where the masked step's extra time goes: the five pipeline stages of one handle with the mask riding along, and the erosion
        # in front of them (it runs before the first stage event) as the remainder of the handle's whole call
        exL.profile(True)
        mstage = np.zeros(5); mcall = 0.0
        for _ in range(K):
            m0.record(sL)
            exL.extract_batch_device(dL.data_ptr(), P, d_masks=dmL.data_ptr())
            m1.record(sL); exL.sync()
            mstage += np.array(exL.stage_ms()); mcall += m0.elapsed_time(m1)
        mstage /= K; mcall /= K
        exL.profile(False)
        npmL, npmR = mLh.numpy(), mRh.numpy()

        def e2e_mstep():
            adb.stereo_frames_batch(exL, exR, npL, npR, mb, mbf, npmL, npmR, out_left=outL, out_right=outR, out_stereo=outS)
            return int(outL.counts.sum() + outR.counts.sum())

        for _ in range(2):
            e2e_mstep()
        t0 = time.perf_counter()
        totm = 0
        for _ in range(K):
            totm += e2e_mstep()
        t_m = time.perf_counter() - t0
        masked = {"value": n_kp_m / (ms_m * 1e-3), "unit": "keypoints/s", "ms_per_step": ms_m, "frames_per_s": 2 * P / (ms_m * 1e-3),
                  "keypoints_per_step": n_kp_m, "masked_pixel_fraction": float((mhost == 0).mean()),
                  "time_vs_unmasked": ms_m / ms_step, "e2e": {"value": totm / t_m, "unit": "keypoints/s", "h2d_bytes_per_step": int(4 * P * W * H)},
                  "stage_ms_one_handle": dict({"erode10_tile_kernel": float(mcall - mstage.sum())}, **{n: float(v) for n, v in zip(names, mstage)}),
                  "note": "same stream with a person-shaped mask per image: cv::erode 10x10 (separable, one HBM pass; tiles whose input is constant "
                          "-- most of a segmentation mask -- are stored without the two minimum passes), mask pyramid, masked FAST"}
        # leave the handles in the unmasked state for the sections below
        step(); sync_all()

    # ---- drop-in latency: one stereo pair per call through the host-buffer C-ABI, as Frame::Frame would call it
    lat_ms = None
    if rank == 0:
        e1L = adb.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, W, H, max_batch=1, device=local)
        e1R = adb.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, W, H, max_batch=1, device=local)
        o1L = adb.orb.HostResults(1, cap, pinned=True); o1R = adb.orb.HostResults(1, cap, pinned=True); o1S = adb.orb.HostStereo(1, cap, pinned=True)

        def one_pair(i):
            tR = threading.Thread(target=e1R.extract_batch, args=(npR[i:i + 1],), kwargs={"out": o1R})
            tR.start()
            e1L.extract_batch(npL[i:i + 1], out=o1L)
            tR.join()
            adb.compute_stereo_matches(e1L, e1R, 1, mb, mbf, out=o1S)

        for i in range(5):
            one_pair(i % P)
        t0 = time.perf_counter()
        for i in range(50):
            one_pair(i % P)
        lat_ms = (time.perf_counter() - t0) / 50 * 1e3
        e1L.close(); e1R.close()

    out = {
        "metric": "orb_keypoints_per_s", "value": value, "unit": "keypoints/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "640x480 stereo stream, 8-level pyramid, 2000 feat/frame, ORB extract L+R + stereo match"
                               + (" + all-gather of descriptor records" if world > 1 else ""),
                   "pairs_per_step_per_gpu": P, "gather": gather_mode, "gather_payload_checked": symm is not None,
                   "nvlink_bytes_sent_per_step_per_gpu": nvlink_bytes, "gathered_record_bytes_per_step_per_gpu": (int(cntL_t.sum().item()) * 56 if world > 1 else None), "frames_per_step": 2 * P * world, "keypoints_per_step": int(nkp.item()),
                   "frames_per_s": 2 * P * world / (ms_step * 1e-3), "single_pair_latency_ms_host_api": lat_ms, "cpu_affinity": numa,
                   "l2_policy": "inputs larger than L2: %.0f MB of images + %.0f MB of pyramid per step vs 126 MB L2"
                                % (2 * P * W * H / 1e6, 2 * P * (PYR_PX - W * H) / 1e6)},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if masked is not None:
        out["masked"] = masked
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        n_s = max(8, threads)
        sample = synth.make_stereo_batch(n_s)
        v, ms, kpf = cpu_reference_run(sample, threads, 3, 1)
        out["cpu_baseline"] = {"value": v, "unit": "keypoints/s", "cores": threads, "kind": "port",
                               "sample": f"3 timed passes (1 warm) over {n_s} stereo pairs ({2 * n_s} frames) of the same workload, oracle port, {threads} host threads "
                                         "(the routine `--impl reference` runs)"}
        sec = cv2_primitive_baseline()
        if "ms_per_frame_one_thread" in sec:   # what a real OpenCV build could reach at best on these cores, in the metric's unit
            sec["keypoints_per_s_upper_bound_all_cores"] = kpf / (2 * n_s) * threads * sec["frames_per_s_per_core_upper_bound"]
        out["cpu_baseline"]["cv2_primitives"] = sec
    if not args.no_ba and rank == 0 and world == 1:   # single-GPU sections (BA stays single-GPU; replicas only)
        try:
            import bench_ba
            out["ba"] = bench_ba.run(local, steps=max(3, K // 2), with_cpu=not args.no_cpu_baseline)
        except ImportError:
            pass
        except Exception as e:   # the headline metric must still print
            out["ba"] = {"error": repr(e)}
        try:
            import bench_ba
            out["ba_dynamic"] = bench_ba.run_dynamic(local, steps=3, with_cpu=not args.no_cpu_baseline)
        except ImportError:
            pass
        except Exception as e:
            out["ba_dynamic"] = {"error": repr(e)}
    if not args.no_ba and rank == 0 and world == 1:
        try:
            import bench_search
            out["search"] = bench_search.run(local, steps=max(3, K // 2))
        except Exception as e:
            out["search"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out))
    exL.close(); exR.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
