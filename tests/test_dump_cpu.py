"""Host logic: the reference's map dump (Tracking::SaveMap, src/Tracking.cc:1745-1838) <-> BA problem, including the
Match.txt records that run together because of the missing newline at src/Tracking.cc:1806-1807."""
import numpy as np
import pytest

from airdos_b200 import dump, synth


def _problem():
    d = synth.make_ba_problem(6, 200, 4, seed=31)
    table, _ = dump.inv_sigma2_table()
    lv = np.random.default_rng(0).integers(0, 8, len(d["edge_info"]))
    d["edge_info"] = table[lv].astype(np.float64)
    return d


def test_match_token_split():
    _, printed = dump.inv_sigma2_table()
    assert printed[:3] == ["1", "0.694444", "0.482253"]
    text = "0 57 12.5 100.25 -1 112 58 3 4 5 0.69444413 59 1 2 3 " + printed[7]
    recs = dump._split_match_tokens(text, printed)
    assert [r[0] for r in recs] == ["0", "12", "13"] and [r[5] for r in recs] == ["1", "0.694444", printed[7]]
    with pytest.raises(ValueError):
        dump._split_match_tokens("0 1 2 3 4 0.5", printed)


@pytest.mark.parametrize("newlines", [False, True])
def test_dump_round_trip(tmp_path, newlines):
    d = _problem()
    cam = {k: d[k] for k in ("fx", "fy", "cx", "cy", "bf")}
    dump.save_map_dump(str(tmp_path), d, match_newlines=newlines)
    if not newlines:
        assert "\n" not in open(tmp_path / "Match.txt").read()
    g = dump.load_map_dump(str(tmp_path), cam)
    assert (g["edge_pose"] == d["edge_pose"]).all()
    # points without observations are dropped; the rest keep their order
    used = np.unique(d["edge_point"])
    assert (g["mp_ids"] - (len(d["pose_q"]) - 1) - 1 == used).all()
    assert (used[g["edge_point"]] == d["edge_point"]).all()
    assert np.allclose(g["points"], d["points"][used], rtol=1e-5, atol=1e-6)
    assert np.allclose(g["edge_obs"], d["edge_obs"], rtol=1e-5, atol=1e-5)
    assert (g["edge_info"] == d["edge_info"].astype(np.float32)).all() or np.allclose(g["edge_info"], d["edge_info"], rtol=1e-5)
    assert g["pose_fixed"][0] == 1 and g["pose_fixed"][1:].sum() == 0
    # same rotation up to the quaternion sign, translation to print precision
    dots = np.abs((g["pose_q"] * d["pose_q"]).sum(1))
    assert (dots > 1 - 1e-9).all() and np.allclose(g["pose_t"], d["pose_t"], atol=1e-4)


def test_match_split_is_linear_and_skips_unknown_points(tmp_path):
    """one pass over the tokens (no list.insert per record), and a Match.txt id that is missing from MP.txt is skipped."""
    import time
    _, printed = dump.inv_sigma2_table()
    n = 200000
    # the reference writes `kf mp u v ur invSigma2` and no separator after it: the next record's KF id is glued to invSigma2
    glued = "".join(f"{i % 50} {60 + i % 1000} 1.5 2.5 -1 {printed[i % 8]}" for i in range(n))
    t0 = time.perf_counter()
    recs = dump._split_match_tokens(glued, printed)
    assert time.perf_counter() - t0 < 5.0 and len(recs) == n
    assert [r[0] for r in recs[:3]] == ["0", "1", "2"] and recs[-1][5] == printed[(n - 1) % 8] and recs[57][:2] == ["7", "117"]
    d = _problem()
    dump.save_map_dump(str(tmp_path), d)
    lines = open(tmp_path / "MP.txt").read().splitlines()
    open(tmp_path / "MP.txt", "w").write("\n".join(lines[:-5]) + "\n")       # the last five points vanish from MP.txt
    g = dump.load_map_dump(str(tmp_path), {k: d[k] for k in ("fx", "fy", "cx", "cy", "bf")})
    assert len(g["edge_pose"]) < len(d["edge_pose"]) and g["edge_point"].max() < len(g["points"])


def test_human_dump_round_trip(tmp_path):
    """HMTraj.txt / Motion.txt (src/Tracking.cc:1812-1830): joints, flags and motions survive; rigidity and motion edges are rebuilt
    with the topology of include/Map.h:49-56; edges that touch a bad / lost joint are left out."""
    d = synth.make_ba_problem(n_kf=10, n_points=150, seed=33, humans=3, human_poses=4)
    table, _ = dump.inv_sigma2_table()
    d["edge_info"] = table[np.random.default_rng(2).integers(0, 8, len(d["edge_info"]))].astype(np.float64)
    flags = np.zeros((len(d["joints"]), 2), np.uint8)
    flags[5, 0] = 1; flags[40, 1] = 1
    dump.save_map_dump(str(tmp_path), d, human_poses=4, joint_flags=flags)
    g = dump.load_map_dump(str(tmp_path), {k: d[k] for k in ("fx", "fy", "cx", "cy", "bf")}, humans=True)
    assert g["joints"].shape == d["joints"].shape and np.allclose(g["joints"], d["joints"], rtol=1e-5, atol=1e-6)
    assert (g["joint_bad"] == flags[:, 0]).all() and (g["joint_lost"] == flags[:, 1]).all()
    assert len(g["motion_t"]) == 3 and np.allclose(g["motion_t"], d["motion_t"], atol=1e-6) and np.allclose(np.abs(g["motion_q"]), np.abs(d["motion_q"]), atol=1e-6)
    # same edges as the generator's, minus those on the flagged joints
    keep_r = [k for k in range(len(d["redge_i"])) if d["redge_i"][k] not in (5, 40) and d["redge_j"][k] not in (5, 40)]
    assert (g["redge_i"] == d["redge_i"][keep_r]).all() and (g["redge_j"] == d["redge_j"][keep_r]).all() and (g["redge_dist"] == d["redge_dist"][keep_r]).all()
    keep_m = [k for k in range(len(d["medge_p1"])) if d["medge_p1"][k] not in (5, 40) and d["medge_p2"][k] not in (5, 40)]
    assert (g["medge_p1"] == d["medge_p1"][keep_m]).all() and (g["medge_p2"] == d["medge_p2"][keep_m]).all() and (g["medge_motion"] == d["medge_motion"][keep_m]).all()
    assert len(g["dists"]) == 3 * 14 and (g["dists"] > 0.05).all() and len(g["jedge_pose"]) == 0
