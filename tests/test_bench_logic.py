"""CPU: the host logic of bench.py that needs no GPU - configuration table vs BASELINE.json, pyramid sizes, the algorithmic
byte formulas of SURVEY.md §8d, the affinity parser, and the reference CPU arm (a few pairs, both kinds)."""
import json
from pathlib import Path

import numpy as np

import bench

ROOT = Path(__file__).resolve().parents[1]


def test_configs_cover_baseline_json():
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert len(base["configs"]) == 5
    c = bench.CONFIGS
    assert set(c) == {"kitti1000", "kitti2000", "euroc", "hd1080", "shard8000"}
    assert c["kitti1000"]["params"] == (1000, 1.2, 8, 20, 5) and (c["kitti1000"]["W"], c["kitti1000"]["H"]) == (1242, 375)
    assert c["kitti2000"]["params"] == (2000, 1.2, 8, 20, 7)
    assert (c["euroc"]["W"], c["euroc"]["H"], c["euroc"]["params"][0]) == (752, 480, 1200)
    assert (c["hd1080"]["W"], c["hd1080"]["H"], c["hd1080"]["params"][0]) == (1920, 1080, 4000)
    assert c["shard8000"]["total_pairs"] == 8000 and c["shard8000"]["params"] == c["kitti2000"]["params"]
    assert bench.metric_name(c["kitti1000"]).startswith(base["metric"][:60])


def test_level_sizes_match_survey_appendix_b():
    assert sum(bench.level_sizes(bench.CONFIGS["kitti1000"])) == 1441432
    assert sum(bench.level_sizes(bench.CONFIGS["euroc"])) == 1117367
    assert sum(bench.level_sizes(bench.CONFIGS["hd1080"])) == 6419321
    assert bench.level_sizes(bench.CONFIGS["kitti1000"])[0] == 465750


def test_stage_bytes_nominal_values():
    """SURVEY §8d nominal KITTI counts -> B_img = 9,530,975 and B_st = 2,040,000 (our K4-6 is fused: it is charged its
    external bytes only, so the describe row is smaller than the survey's unfused K4 + K5 + K6)."""
    cfg = bench.CONFIGS["kitti2000"]
    tot = dict(pairs=1, candidates=2 * 10000, keypoints=2 * 2000, sad_evals=600, matches=400, row_candidates=2000 * 58,
               hamming_evals=2000 * 10)
    b, counts = bench.stage_bytes_per_pair(cfg, tot)
    assert counts["row_candidates_per_left_keypoint"] == 58 and counts["hamming_evals_per_left_keypoint"] == 10
    st = 60 * 4000 + 2000 * (8 * 58 + 32 * 10) + 352 * 600 + 8 * 2000 + 8 * 600
    assert b["stereo_match"] == st == 2040000
    SP, P0, PL = 1441432, 465750, 36435
    assert b["pyramid"] == 2 * ((SP - PL) + (SP - P0)) and b["fast"] == 2 * (SP + 80000)
    assert b["describe"] == 2 * 2000 * 1913


def test_affinity_parser(monkeypatch):
    class R:
        stdout = "\t\x1b[4mGPU0\tGPU1\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\nGPU0\t X \tNV18\t0-15,64-79\t0\t\tN/A\nGPU1\tNV18\t X \t16-31\t1\t\tN/A\n"
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: R())
    assert bench.gpu_cpu_affinity(0) == set(range(0, 16)) | set(range(64, 80))
    assert bench.gpu_cpu_affinity(1) == set(range(16, 32))
    assert bench.gpu_cpu_affinity(5) is None


def test_cpu_arm_runs_reference_and_port():
    from pointslot_b200.synth import synth_pair
    cfg = dict(bench.CONFIGS["kitti1000"]); cfg["W"], cfg["H"] = 640, 360
    L, R = synth_pair(640, 360, 4)
    for pref in (True, False):
        arm = bench.CpuArm(cfg, prefer_reference=pref)
        assert arm.kind in ("reference", "port") and (pref or arm.kind == "port")
        fn = arm.make_pair_fn()
        assert fn(L, R) > 100 and fn(L, R, True) > 100
        v, done, el = arm.pairs_per_sec([L], [R], 2, 0.5)
        assert done >= 1 and v > 0


def test_proportional_shards():
    """End-to-end shards sized by each GPU's upload rate: same total, every rank at least one sub-batch."""
    a = bench.proportional_shards(1280, [23.3] * 4 + [36.0] * 4)
    assert sum(a) == 1280 and a[:4] == [a[0]] * 4 and a[4] > a[0] and abs(a[4] / a[0] - 36.0 / 23.3) < 0.02
    assert bench.proportional_shards(160, [55.5]) == [160]
    b = bench.proportional_shards(10, [1, 1, 1000])
    assert sum(b) == 10 and min(b) >= 1
    assert bench.proportional_shards(8, [1.0] * 8) == [1] * 8


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line with the GPU arm's metric / unit / static config keys,
    "impl": "reference", a cpu_baseline describing the run and an e2e object without copies."""
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    cfg = bench.CONFIGS["kitti1000"]
    assert line["impl"] == "reference" and line["metric"] == bench.metric_name(cfg) and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 0 and line["value"] > 0
    assert line["config"]["workload"] == cfg["workload"] and line["config"]["params"] == list(cfg["params"])
    assert line["config"]["maxD"] == bench.max_disparity(cfg)
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["cores"] >= 1 and line["sample"]["pairs_per_step"] >= 64
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
