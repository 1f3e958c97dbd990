"""GPU parity of the feature grid: pslot_assign_grid / pslot_features_in_area /
pslot_batch_assign_grid against the oracle (Frame.cc:1636-1655, 2027-2038, 1808-1861).  Index work: bit-exact."""
import numpy as np
import pytest

from test_oracle_grid import COLS, ROWS, random_keypoints

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def B():
    from pointslot_b200 import binding
    assert binding.device_count() > 0
    return binding


@pytest.mark.parametrize("n,seed", [(0, 1), (1, 2), (1500, 3), (9000, 4), (100000, 5)])  # the last exceeds the on-chip cell cache
def test_assign_grid_random(n, seed, B, oracle_mod):
    W, H = 1241, 376
    e = B.OrbExtractor(1000, 1.2, 8, 20, 7)
    spec = B.GridSpec.for_image(W, H)
    k = random_keypoints(oracle_mod, n, W, H, seed)
    cs, idx = B.assign_grid(e, k, spec)
    ocs, oidx = oracle_mod.assign_grid(k, spec.min_x, spec.min_y, spec.inv_w, spec.inv_h, COLS, ROWS)
    assert np.array_equal(cs, ocs) and np.array_equal(idx, oidx)


def test_assign_grid_crowded_and_offset(B, oracle_mod):
    """All keypoints in a few cells (long per-cell lists: exercises the ordering pass) and non-zero bounds."""
    rng = np.random.default_rng(8)
    k = np.zeros(4000, oracle_mod.KEYPOINT_DTYPE)
    k["x"] = rng.uniform(100, 160, 4000).astype(F); k["y"] = rng.uniform(50, 70, 4000).astype(F)
    e = B.OrbExtractor(1000, 1.2, 8, 20, 7)
    spec = B.GridSpec(-11.3, -7.9, float(F(64) / F(781.5)), float(F(48) / F(497.5)), 64, 48)
    cs, idx = B.assign_grid(e, k, spec)
    ocs, oidx = oracle_mod.assign_grid(k, spec.min_x, spec.min_y, spec.inv_w, spec.inv_h, 64, 48)
    assert np.array_equal(cs, ocs) and np.array_equal(idx, oidx)
    assert np.diff(cs).max() > 100
    # other grid shapes
    spec2 = B.GridSpec(0.0, 0.0, float(F(20) / F(300)), float(F(10) / F(100)), 20, 10)
    cs, idx = B.assign_grid(e, k, spec2)
    ocs, oidx = oracle_mod.assign_grid(k, 0.0, 0.0, spec2.inv_w, spec2.inv_h, 20, 10)
    assert np.array_equal(cs, ocs) and np.array_equal(idx, oidx)


def test_features_in_area(B, oracle_mod):
    from pointslot_b200.synth import synth_pair
    W, H = 1242, 375
    L, _ = synth_pair(W, H, 21)
    e = B.OrbExtractor(2000, 1.2, 8, 20, 7)
    k, _ = e.extract(L)
    spec = B.GridSpec.for_image(W, H)
    cs, idx = B.assign_grid(e, k, spec)
    ocs, oidx = oracle_mod.assign_grid(k, 0.0, 0.0, spec.inv_w, spec.inv_h, COLS, ROWS)
    assert np.array_equal(cs, ocs) and np.array_equal(idx, oidx)
    rng = np.random.default_rng(22)
    nq = 2000
    q = np.zeros(nq, B.AREA_QUERY_DTYPE)
    q["x"] = rng.uniform(-80, W + 80, nq); q["y"] = rng.uniform(-80, H + 80, nq)
    q["r"] = rng.choice([2.0, 10.0, 35.0, 150.0, 2000.0], nq)
    lv = np.array([(-1, -1), (0, -1), (2, -1), (-1, 3), (1, 4), (5, 2)])[rng.integers(0, 6, nq)]
    q["min_level"], q["max_level"] = lv[:, 0], lv[:, 1]
    got = B.features_in_area(e, q, cap=len(k))
    for i in range(nq):
        ref = oracle_mod.features_in_area(k, ocs, oidx, 0.0, 0.0, spec.inv_w, spec.inv_h, COLS, ROWS, float(q["x"][i]),
                                          float(q["y"][i]), float(q["r"][i]), int(q["min_level"][i]), int(q["max_level"][i]))
        assert np.array_equal(got[i], ref), f"query {i}"
    assert sum(len(g) for g in got) > nq
    with pytest.raises(B.PslotError):
        B.features_in_area(e, q[q["r"] > 1000][:4], cap=8)  # capacity overflow is reported, not truncated silently


def test_batch_assign_grid(B, oracle_mod):
    from pointslot_b200.synth import synth_batch
    params = (1000, 1.2, 8, 20, 5)
    W, H, n = 1242, 375, 5
    Ls, Rs = synth_batch(W, H, range(60, 60 + n))
    sb = B.StereoBatch(params, W, H, 8)
    out = sb.alloc_out(n)
    sb.submit(Ls, Rs, 384.38148, 721.5377, out)
    spec = B.GridSpec.for_image(W, H)
    cs, idx = sb.assign_grid(spec, n)
    sb.wait()
    for i in range(n):
        k = out["kps"][i][: out["counts"][i]]
        ocs, oidx = oracle_mod.assign_grid(k, 0.0, 0.0, spec.inv_w, spec.inv_h, COLS, ROWS)
        assert np.array_equal(cs[i], ocs) and np.array_equal(idx[i][: len(oidx)], oidx)
