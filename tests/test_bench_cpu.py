"""Host logic of bench.py that needs no GPU: the reference arm (`--impl reference`) runs the CPU oracle port on the host
cores and prints one JSON line with the contract's keys; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run()
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "orb_keypoints_per_s" and d["unit"] == "keypoints/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 1e4 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "640x480" in d["config"]["workload"]


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_roofline_traffic_comes_from_the_unmasked_capture():
    """`roofline.traffic` is scaled from the committed ncu capture of the UNMASKED extractor kernels (profiles/r2c_dram_traffic.json),
    never from whichever *_dram_traffic.json happens to be newest on disk (profiles/r2j_masked_* holds the masked instances under the same
    kernel names: twice the DRAM bytes)."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.dram_traffic("fast_cells_warp_kernel", 3072)
    cap = json.load(open(os.path.join(ROOT, "profiles", bench.TRAFFIC_CAPTURE)))
    assert src == "profiles/r2c_dram_traffic.json"
    assert t == sum(e["dram_bytes"] for e in cap["fast_cells_warp_kernel"][:2]) / 128.0 * 3072
    alg = 3072 * (950532 + 102400)          # one pyramid read + the candidate slots written, per frame: same order
    assert 0.5 * alg < t < 1.5 * alg
    assert bench.dram_traffic("pyr_resize_strip_kernel", 3072) == (None, None)      # seven launches per call: not a per-launch figure
    assert bench.dram_traffic("no_such_kernel", 3072) == (None, None)


def test_issue_roofline_of_the_fast_kernel():
    """The secondary roofline of the dominant kernel (issue slots) is computed from the committed ncu capture and the
    live stage time; a missing kernel or a zero time gives None, never an exception inside the bench."""
    sys.path.insert(0, ROOT)
    import bench
    r = bench.issue_roofline("fast_cells_warp_kernel", 3072, 11.2, 1965.0)
    assert r["bound"] == "issue" and abs(r["peak"] - 148 * 4 * 1.965) < 1e-6
    assert 0.6 < r["frac"] < 0.8 and abs(r["achieved"] / r["peak"] - r["frac"]) < 1e-12      # ncu: 70.7 % issue-active on the same kernel
    assert bench.issue_roofline("fast_cells_warp_kernel", 3072, 11.2, None)["peak"] == r["peak"]        # nominal clock when nvidia-smi gave none
    assert bench.issue_roofline("no_such_kernel", 3072, 11.2, 1965.0) is None
    assert bench.issue_roofline("fast_cells_warp_kernel", 3072, 0.0, 1965.0) is None
