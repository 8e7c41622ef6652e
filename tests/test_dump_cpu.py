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
