"""Pins the LocalBundleAdjustment schedule of oracle/ba_oracle.cpp (ba_oracle_solve) to the LITERAL reference: tests/golden/lba_ref.npz
holds what the reference's own Optimizer::LocalBundleAdjustment (src/Optimizer.cc:431-731, the whole function compiled from
/root/reference: oracle/ref_lba.cpp) does on five seeded covisibility windows -- which key-frames it frees and fixes, the edges it
builds (information from mvInvLevelSigma2[octave], Huber deltas sqrt(5.991) / sqrt(7.815) as floats), optimize(5) with the kernel,
the chi2 / depth gate with the reference's own edge types, optimize(10) without kernel and without the gated edges, the erase list and
the poses / positions written back through Converter.  Its LM control is the reference's too (oracle/ref_lm.cpp); its solver steps are
the oracle's.  ba_oracle_solve on the problem the function built must give every LM trial, the final estimates and the erase list
bit for bit; oracle/gen_ref_lba_golden.py wrote the fixture."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "lba_ref.npz")
REF = os.path.join(ROOT, "oracle", "_ref")
HAVE_REF = os.path.exists(os.path.join(REF, "libref_lba.so")) and os.path.isdir("/root/reference")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_ref_lba_golden", os.path.join(ROOT, "oracle", "gen_ref_lba_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    return g


def fixture_problem(gold, i):
    return {k[len(f"w{i}_p_"):]: (gold[k].item() if gold[k].ndim == 0 else gold[k]) for k in gold.files if k.startswith(f"w{i}_p_")}


def erase_list_from_flags(gold, i, flags):
    """(key-frame index, map-point index) pairs in the reference's order: monocular edges first, then stereo (src/Optimizer.cc:672-700)."""
    p = fixture_problem(gold, i)
    g = _gen()
    w = g.make_window(i)
    kf_of = {int(v): j for j, v in enumerate(w["kf_id"])}; mp_of = {int(v): j for j, v in enumerate(w["mp_id"])}
    max_kf = int(w["kf_id"].max())
    mono = p["edge_obs"][:, 2] < 0
    pairs = []
    for sel in (mono, ~mono):
        for e in np.flatnonzero(sel & (flags != 0)):
            pairs.append((kf_of[int(gold[f"w{i}_pose_id"][p["edge_pose"][e]])], mp_of[int(gold[f"w{i}_point_id"][p["edge_point"][e]]) - max_kf - 1]))
    return np.array(pairs, np.int32).reshape(-1, 2)


def test_oracle_schedule_equals_the_reference_function(oracle_mod):
    gold = np.load(GOLD)
    g = _gen()
    for i in range(len(g.CASES)):
        prob = fixture_problem(gold, i)
        o = oracle_mod.ba_default_options()
        assert (o.huber_mono, o.huber_stereo) == tuple(gold[f"w{i}_huber"])           # thHuberMono / thHuberStereo (:538-539)
        assert [o.robust[0], o.robust[1]] == list(gold[f"w{i}_round_robust"])          # kernel in round 1, setRobustKernel(0) before round 2
        p, res, st = oracle_mod.ba_solve(prob, o)
        assert st == 0
        assert [res.c.iterations_run[0], res.c.iterations_run[1]] == list(gold[f"w{i}_round_iterations"]), i
        tr = res.trace_rows[:, [0, 1, 2, 4]]
        assert tr.shape == gold[f"w{i}_rows"].shape and (tr == gold[f"w{i}_rows"]).all(), i
        state = np.concatenate([p["pose_q"].ravel(), p["pose_t"].ravel(), p["points"].ravel()])
        assert (state == gold[f"w{i}_final_state"]).all(), i
        assert (erase_list_from_flags(gold, i, res.edge_outlier) == gold[f"w{i}_erased"]).all(), i


def test_written_back_poses_and_positions(oracle_mod):
    """KeyFrame::SetPose(Converter::toCvMat(SE3Quat)) and MapPoint::SetWorldPos(Converter::toCvMat(Vector3d)) of the function against the
    oracle's converters on the oracle's result; key-frames the function fixed keep their pose, every local point is updated once."""
    import ctypes as C
    gold = np.load(GOLD)
    g = _gen()
    lib = oracle_mod.ba_lib()
    lib.ba_oracle_pose_to_tcw.argtypes = [C.c_void_p] * 3
    for i in range(len(g.CASES)):
        w = g.make_window(i)
        prob = fixture_problem(gold, i)
        p, res, st = oracle_mod.ba_solve(prob)
        kf_of = {int(v): j for j, v in enumerate(w["kf_id"])}
        n_local = 1 + len(w["covisible"])
        for j, vid in enumerate(gold[f"w{i}_pose_id"]):
            k = kf_of[int(vid)]
            T = np.zeros(16, np.float32)
            q = np.ascontiguousarray(p["pose_q"][j]); t = np.ascontiguousarray(p["pose_t"][j])
            lib.ba_oracle_pose_to_tcw(q.ctypes.data, t.ctypes.data, T.ctypes.data)
            if k < n_local:
                assert (T.reshape(4, 4) == gold[f"w{i}_kf_tcw"][k]).all(), (i, k)
            else:                                                      # fixed cameras are never written back (:712-718 walks lLocalKeyFrames)
                assert (gold[f"w{i}_kf_tcw"][k] == w["kf_tcw"][k]).all()
        mp_of = {int(v): j for j, v in enumerate(w["mp_id"])}
        max_kf = int(w["kf_id"].max())
        seen = np.zeros(len(w["mp_id"]), bool)
        for l, vid in enumerate(gold[f"w{i}_point_id"]):
            m = mp_of[int(vid) - max_kf - 1]
            seen[m] = True
            assert (p["points"][l].astype(np.float32) == gold[f"w{i}_mp_pos"][m]).all(), (i, m)
        assert (gold[f"w{i}_mp_updates"] == seen.astype(np.int32)).all()              # UpdateNormalAndDepth once per local point
        assert (gold[f"w{i}_mp_pos"][~seen] == w["mp_pos"][~seen]).all()


def test_problem_the_function_built(oracle_mod):
    """Vertex order (ascending id), the fixed flags (mnId == 0 and the cameras outside the covisibility list), edge information and the
    stereo / mono split of the recorded problem follow from the window."""
    gold = np.load(GOLD)
    g = _gen()
    for i in range(len(g.CASES)):
        w = g.make_window(i)
        p = fixture_problem(gold, i)
        pid = gold[f"w{i}_pose_id"]
        assert (np.diff(pid) > 0).all() and (np.diff(gold[f"w{i}_point_id"]) > 0).all()
        kf_of = {int(v): j for j, v in enumerate(w["kf_id"])}
        n_local = 1 + len(w["covisible"])
        want_fixed = np.array([int(v) == 0 or kf_of[int(v)] >= n_local for v in pid])
        assert (p["pose_fixed"].astype(bool) == want_fixed).all(), i
        assert set(np.unique(p["edge_info"].astype(np.float32))) <= set(w["inv_level_sigma2"])
        assert len(p["edge_pose"]) == len(w["obs_kf"])                                # every observation of a local point became an edge
        assert int((p["edge_obs"][:, 2] < 0).sum()) == int((w["obs_uvr"][:, 2] < 0).sum())   # mvuRight < 0 -> EdgeSE3ProjectXYZ


@pytest.mark.skipif(not HAVE_REF, reason="reference tree / oracle/_ref not present (GPU box)")
def test_fixture_is_what_the_reference_library_computes_now(oracle_mod):
    import ctypes as C
    g = _gen()
    gold = np.load(GOLD)
    LM = C.CDLL(os.path.join(REF, "libref_lm.so")); LBA = C.CDLL(os.path.join(REF, "libref_lba.so"))
    for i in (1, 4):
        r = oracle_mod.ref_local_bundle_adjustment(LBA, LM, g.make_window(i))
        for k in ("kf_tcw", "mp_pos", "erased", "rows", "final_state"):
            assert (r[k] == gold[f"w{i}_{k}"]).all(), (i, k)
    j = 2
    r = oracle_mod.ref_local_bundle_adjustment(LBA, LM, g.make_gba_window(j), global_ba=g.GBA_CASES[j][1:])
    assert (r["rows"] == gold[f"g{j}_rows"]).all() and (r["final_state"] == gold[f"g{j}_final_state"]).all() and (r["kf_tcw"] == gold[f"g{j}_kf_tcw"]).all()
    c = 1
    w, hum = g.make_human_window(c)
    r = oracle_mod.ref_local_bundle_adjustment(LBA, LM, w, humans=hum)
    assert (r["rows"] == gold[f"h{c}_rows"]).all() and (r["final_state"] == gold[f"h{c}_final_state"]).all()
    assert (r["edge_final_chi2"] == gold[f"h{c}_edge_final_chi2"]).all() and (r["humans"]["key_flags"] == gold[f"h{c}_hum_key_flags"]).all()


def gba_problem(gold, j):
    return {k[len(f"g{j}_p_"):]: (gold[k].item() if gold[k].ndim == 0 else gold[k]) for k in gold.files if k.startswith(f"g{j}_p_")}


def test_oracle_global_ba_equals_the_reference_function(oracle_mod):
    """Optimizer::BundleAdjustment (src/Optimizer.cc:60-230, what GlobalBundleAdjustemnt calls), the whole function compiled from
    /root/reference: one round of nIterations with thHuber2D = sqrt(5.99) (not 5.991) / thHuber3D = sqrt(7.815) or no kernel, all key-frames,
    mnId 0 fixed, unobserved points removed.  ba_oracle_solve with ba_global_options on the problem it built: trials and final state
    bit for bit; results land in the poses (nLoopKF == 0) or in mTcwGBA / mPosGBA with mnBAGlobalForKF = nLoopKF."""
    import ctypes as C
    gold = np.load(GOLD)
    g = _gen()
    lib = oracle_mod.ba_lib()
    lib.ba_oracle_pose_to_tcw.argtypes = [C.c_void_p] * 3
    for j, (c, its, loop_kf, robust) in enumerate(g.GBA_CASES):
        w = g.make_gba_window(j)
        prob = gba_problem(gold, j)
        o = oracle_mod.ba_global_options(its, robust)
        if robust:
            assert (o.huber_mono, o.huber_stereo) == tuple(gold[f"g{j}_huber"])
            assert o.huber_mono == float(np.float32(np.sqrt(5.99)))
        p, res, st = oracle_mod.ba_solve(prob, o)
        assert st == 0 and [res.c.iterations_run[0]] == list(gold[f"g{j}_round_iterations"]) and res.c.iterations_run[1] == 0
        assert list(gold[f"g{j}_round_robust"]) == [int(robust)]
        tr = res.trace_rows[:, [0, 1, 2, 4]]
        assert tr.shape == gold[f"g{j}_rows"].shape and (tr == gold[f"g{j}_rows"]).all(), j
        assert (np.concatenate([p["pose_q"].ravel(), p["pose_t"].ravel(), p["points"].ravel()]) == gold[f"g{j}_final_state"]).all(), j
        # every key-frame is a vertex, only mnId 0 is fixed; the unobserved point is not in the problem
        assert len(gold[f"g{j}_pose_id"]) == len(w["kf_id"]) and (prob["pose_fixed"].astype(bool) == (gold[f"g{j}_pose_id"] == 0)).all()
        assert len(gold[f"g{j}_point_id"]) == len(w["mp_id"]) - 1
        kf_of = {int(v): k for k, v in enumerate(w["kf_id"])}
        for i, vid in enumerate(gold[f"g{j}_pose_id"]):
            T = np.zeros(16, np.float32)
            q = np.ascontiguousarray(p["pose_q"][i]); t = np.ascontiguousarray(p["pose_t"][i])
            lib.ba_oracle_pose_to_tcw(q.ctypes.data, t.ctypes.data, T.ctypes.data)
            assert (T.reshape(4, 4) == gold[f"g{j}_kf_tcw"][kf_of[int(vid)]]).all(), (j, i)
        upd = gold[f"g{j}_mp_updates"]
        assert (upd[:-1] == (loop_kf if loop_kf else 1)).all() and upd[-1] == 0          # mnBAGlobalForKF / UpdateNormalAndDepth; removed point untouched


def hba_problem(gold, c):
    return {k[len(f"h{c}_p_"):]: (gold[k].item() if gold[k].ndim == 0 else gold[k]) for k in gold.files if k.startswith(f"h{c}_p_")}


def test_oracle_dynamic_ba_equals_the_reference_function(oracle_mod):
    """Optimizer::LocalBundleAdjustmentHumanTrajactory (src/Optimizer.cc:1496-2222), the whole function compiled from /root/reference with
    the AirDOS vertex / edge types (VertexDistanceDouble, VertexSE3, EdgeRigidBodyDouble, LandmarkMotionTernaryEdge: their computeError and
    chi2() decide the gates) and stand-ins of MapHumanTrajectory / MapHumanPose / MapHumanKey / Rigidbody / Map: which trajectories and
    poses enter (thLongTrajectory, reference key-frame inside the window), vertex ids by kind, the joint / rigidity / motion edges with
    SigmaHuman / SigmaRigidity / SigmaMotion and the kernels sqrt(7.815) / thRanSacRigidity / sqrt(thRanSacMotion), optimize(5), the four
    gates + setRobustKernel(0), optimize(10), and the epilogue.  The Jacobians of the rigidity and motion edges are the oracle's
    (conventions D.4 / D.6: the reference's are undefined); everything else of the run is the reference's.  ba_oracle_solve on the
    problem the function built: every LM trial, the final estimates and all four outlier sets bit for bit."""
    gold = np.load(GOLD)
    g = _gen()
    seen = set()
    for c in range(len(g.HBA_CASES)):
        prob = hba_problem(gold, c)
        o = oracle_mod.ba_default_options()
        hub = gold[f"h{c}_huber"]
        assert hub[1] == o.huber_stereo
        if len(prob["redge_i"]):
            assert (hub[2], hub[3]) == (o.huber_rigid, o.huber_motion) == (1.0, 2.0)      # rk->setDelta(thRanSacRigidity), sqrt(thRanSacMotion) as float
            assert (o.chi2_rigid, o.chi2_motion) == (g.HBA_SIGMAS["th_rigidity"], g.HBA_SIGMAS["th_motion"])
        p, res, st = oracle_mod.ba_solve(prob, o)
        assert st == 0 and [res.c.iterations_run[0], res.c.iterations_run[1]] == list(gold[f"h{c}_round_iterations"])
        tr = res.trace_rows[:, [0, 1, 2, 4]]
        assert tr.shape == gold[f"h{c}_rows"].shape and (tr == gold[f"h{c}_rows"]).all(), c
        state = np.concatenate([p[k].ravel() for k in ("pose_q", "pose_t", "points", "joints", "dists", "motion_q", "motion_t")])
        assert state.shape == gold[f"h{c}_final_state"].shape and (state == gold[f"h{c}_final_state"]).all(), c
        kind, chi, dep = gold[f"h{c}_edge_kind"], gold[f"h{c}_edge_final_chi2"], gold[f"h{c}_edge_final_depth_positive"]
        assert ((kind == 1) | (kind == 2) | (kind == 3) | (kind == 4)).all()             # the function builds stereo edges only
        assert (res.edge_outlier == ((chi > 7.815) | (dep == 0))[kind == 1]).all(), c
        assert int(res.edge_outlier.sum()) == len(gold[f"h{c}_erased"])
        assert (res.jedge_outlier == ((chi > 7.815) | (dep == 0))[kind == 2]).all(), c
        assert (res.redge_outlier == (chi > g.HBA_SIGMAS["th_rigidity"])[kind == 3]).all(), c
        assert (res.medge_outlier == (chi > g.HBA_SIGMAS["th_motion"])[kind == 4]).all(), c
        # the epilogue's bookkeeping follows the same flags: one HumanKeyPair per rigidity edge, mnBadTrack per motion outlier
        pf = gold[f"h{c}_hum_pair_flags"]
        assert int(pf[:, :, 0].sum()) == int(res.redge_outlier.sum()) and int(pf[:, :, 1].sum()) == int((res.redge_outlier == 0).sum())
        assert int(gold[f"h{c}_hum_traj_out"][:, 0].sum()) == int(res.medge_outlier.sum())
        if len(prob["redge_i"]):
            assert (gold[f"h{c}_hum_traj_out"][:, 1] == 1).all() and int(gold[f"h{c}_hum_n_optimized_tracks"]) == len(gold[f"h{c}_motion_id"])
            seen.add("articulated")
        else:
            seen.add("short trajectories left out")
        if res.redge_outlier.any(): seen.add("rigidity outliers")
        if res.medge_outlier.any(): seen.add("motion outliers")
        if (tr[:, 3] == 0).any(): seen.add("rejected trial")
    assert seen >= {"articulated", "short trajectories left out", "rigidity outliers", "motion outliers", "rejected trial"}


def test_dynamic_ba_writes_back_joints_and_motion(oracle_mod):
    """MapHumanPose::SetHumanKeyPos(Converter::toCvMat(vertex estimate)) for every key whose vertex exists, mTMotion = the motion vertex."""
    gold = np.load(GOLD)
    g = _gen()
    for c in range(len(g.HBA_CASES)):
        prob = hba_problem(gold, c)
        if not len(prob["redge_i"]):
            continue
        w, hum = g.make_human_window(c)
        p, res, st = oracle_mod.ba_solve(prob, oracle_mod.ba_default_options())
        key_of = {int(v): i for i, v in enumerate(hum["key_id"].ravel())}
        max_mt = int(gold[f"h{c}_motion_id"].max())
        out = gold[f"h{c}_hum_key_pos"].reshape(-1, 3)
        touched = np.zeros(len(out), bool)
        for l, vid in enumerate(gold[f"h{c}_joint_id"]):
            i = key_of[int(vid) - max_mt - 1]
            touched[i] = True
            assert (p["joints"][l].astype(np.float32) == out[i]).all(), (c, l)
        assert (out[~touched] == hum["key_pos"].reshape(-1, 3)[~touched]).all()           # keys without a vertex keep their position
        for m in range(len(gold[f"h{c}_motion_id"])):
            T = gold[f"h{c}_hum_traj_motion"][m]
            assert np.abs(T[:3, 3] - p["motion_t"][m]).max() < 1e-6 and T[3, 3] == 1


@pytest.mark.skipif(not HAVE_REF, reason="reference tree / oracle/_ref not present (GPU box)")
def test_reference_function_at_baseline_config4_size(oracle_mod):
    """BASELINE.json configs[3] (50 key-frames, 20 000 map points, 120 000 stereo edges): the reference's own LocalBundleAdjustment run live
    (nothing this large is stored) against ba_oracle_solve on the problem it built: every trial, the final state and the erase list."""
    import ctypes as C
    g = _gen()
    LM = C.CDLL(os.path.join(REF, "libref_lm.so")); LBA = C.CDLL(os.path.join(REF, "libref_lba.so"))
    from airdos_b200 import synth
    d = synth.make_ba_problem(50, 20000, 6, seed=4000)
    assert len(d["edge_pose"]) == 120000
    K = len(d["pose_t"]); cur = K - 1
    order = [cur] + [k for k in range(K) if k != cur]
    pos = {k: j for j, k in enumerate(order)}
    lib = oracle_mod.ba_lib()
    lib.ba_oracle_pose_to_tcw.argtypes = [C.c_void_p] * 3
    tcw = np.zeros((K, 4, 4), np.float32)
    for j, k in enumerate(order):
        q = np.ascontiguousarray(d["pose_q"][k]); t = np.ascontiguousarray(d["pose_t"][k]); T = np.zeros(16, np.float32)
        lib.ba_oracle_pose_to_tcw(q.ctypes.data, t.ctypes.data, T.ctypes.data)
        tcw[j] = T.reshape(4, 4)
    sig = oracle_mod.orb_tables(2000, 1.2, 8)["inv_sigma2"]
    octave = np.argmin(np.abs(sig[None, :] - d["edge_info"].astype(np.float32)[:, None]), axis=1).astype(np.int32)
    w = dict(kf_id=np.array([k * 2 for k in order], np.int32), kf_tcw=tcw, covisible=np.arange(1, K, dtype=np.int32), fx=d["fx"], fy=d["fy"], cx=d["cx"], cy=d["cy"],
             bf=d["bf"], inv_level_sigma2=sig, mp_id=np.arange(len(d["points"]), dtype=np.int32) * 3 + 1, mp_pos=d["points"].astype(np.float32),
             obs_kf=np.array([pos[k] for k in d["edge_pose"]], np.int32), obs_mp=d["edge_point"].astype(np.int32), obs_uvr=d["edge_obs"].astype(np.float32),
             obs_octave=octave)
    r = oracle_mod.ref_local_bundle_adjustment(LBA, LM, w)
    prob = r["problem"]
    assert len(prob["edge_pose"]) == 120000 and len(r["pose_id"]) == 50 and len(r["point_id"]) == 20000
    p, res, st = oracle_mod.ba_solve(prob)
    tr = res.trace_rows[:, [0, 1, 2, 4]]
    assert st == 0 and tr.shape == r["rows"].shape and (tr == r["rows"]).all()
    assert (np.concatenate([p["pose_q"].ravel(), p["pose_t"].ravel(), p["points"].ravel()]) == r["final_state"]).all()
    assert int(res.edge_outlier.sum()) == len(r["erased"]) > 1000


@pytest.mark.skipif(not HAVE_REF, reason="reference tree / oracle/_ref not present (GPU box)")
def test_reference_dynamic_function_at_baseline_config5_size(oracle_mod):
    """BASELINE.json configs[4] (80 key-frames, 30 000 map points, 180 000 stereo edges, 16 skeletons = 4 trajectories x 4 poses: 224 joints,
    56 bone lengths, 4 motions): the reference's own LocalBundleAdjustmentHumanTrajactory run live against ba_oracle_solve."""
    import ctypes as C
    g = _gen()
    LM = C.CDLL(os.path.join(REF, "libref_lm.so")); LBA = C.CDLL(os.path.join(REF, "libref_lba.so"))
    g.HBA_CASES.append((5000, 80, 30000, 0, 4, 4, 3, False, 6))
    w, hum = g.make_human_window(len(g.HBA_CASES) - 1)
    r = oracle_mod.ref_local_bundle_adjustment(LBA, LM, w, humans=hum)
    prob = r["problem"]
    assert (len(prob["edge_pose"]), len(r["joint_id"]), len(r["dist_id"]), len(r["motion_id"])) == (180000, 224, 56, 4)
    assert (len(prob["jedge_pose"]), len(prob["redge_i"]), len(prob["medge_p1"])) == (224, 224, 60)
    p, res, st = oracle_mod.ba_solve(prob, g.hba_options(oracle_mod, r))
    tr = res.trace_rows[:, [0, 1, 2, 4]]
    assert st == 0 and tr.shape == r["rows"].shape and (tr == r["rows"]).all()
    state = np.concatenate([p[k].ravel() for k in ("pose_q", "pose_t", "points", "joints", "dists", "motion_q", "motion_t")])
    assert (state == r["final_state"]).all()
    kind, chi, dep = r["edge_kind"], r["edge_final_chi2"], r["edge_final_depth_positive"]
    assert (res.edge_outlier == ((chi > 7.815) | (dep == 0))[kind == 1]).all() and (res.jedge_outlier == ((chi > 7.815) | (dep == 0))[kind == 2]).all()
    assert (res.redge_outlier == (chi > 1.0)[kind == 3]).all() and (res.medge_outlier == (chi > 4.0)[kind == 4]).all()
