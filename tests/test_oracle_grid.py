"""Oracle for the feature grid against a literal Python transcription of
Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (reference src/Frame.cc:1636-1655, 2027-2038, 1808-1861)."""
import math

import numpy as np

F = np.float32
COLS, ROWS = 64, 48


def c_round(v):  # C round(): half away from zero
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)


def literal_grid(kps, min_x, min_y, inv_w, inv_h):
    grid = [[[] for _ in range(ROWS)] for _ in range(COLS)]
    for i, kp in enumerate(kps):
        px = c_round(float(F(F(kp["x"] - F(min_x)) * F(inv_w))))
        py = c_round(float(F(F(kp["y"] - F(min_y)) * F(inv_h))))
        if px < 0 or px >= COLS or py < 0 or py >= ROWS:
            continue
        grid[px][py].append(i)
    return grid


def literal_area(kps, grid, min_x, min_y, inv_w, inv_h, x, y, r, min_level, max_level):
    x, y, r, min_x, min_y, inv_w, inv_h = map(F, (x, y, r, min_x, min_y, inv_w, inv_h))
    out = []
    a = max(0, int(math.floor(float(F(F(F(x - min_x) - r) * inv_w)))))
    if a >= COLS:
        return out
    b = min(COLS - 1, int(math.ceil(float(F(F(F(x - min_x) + r) * inv_w)))))
    if b < 0:
        return out
    c = max(0, int(math.floor(float(F(F(F(y - min_y) - r) * inv_h)))))
    if c >= ROWS:
        return out
    d = min(ROWS - 1, int(math.ceil(float(F(F(F(y - min_y) + r) * inv_h)))))
    if d < 0:
        return out
    check = min_level > 0 or max_level >= 0
    for ix in range(a, b + 1):
        for iy in range(c, d + 1):
            for j in grid[ix][iy]:
                kp = kps[j]
                if check:
                    if kp["octave"] < min_level:
                        continue
                    if max_level >= 0 and kp["octave"] > max_level:
                        continue
                if abs(F(kp["x"] - x)) < r and abs(F(kp["y"] - y)) < r:
                    out.append(j)
    return out


def random_keypoints(oracle_mod, n, W, H, seed, spill=30.0):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, oracle_mod.KEYPOINT_DTYPE)
    k["x"] = rng.uniform(-spill, W + spill, n).astype(np.float32)
    k["y"] = rng.uniform(-spill, H + spill, n).astype(np.float32)
    k["octave"] = rng.integers(0, 8, n)
    # exact cell-boundary cases: (x - min) * inv lands on k + 0.5 for some entries
    k["x"][: n // 8] = (np.arange(n // 8) % 70 + 0.5).astype(np.float32) * np.float32(W) / np.float32(COLS)
    return k


def test_assign_grid_matches_literal(oracle_mod):
    W, H = 1241, 376
    inv_w, inv_h = float(F(COLS) / F(W)), float(F(ROWS) / F(H))
    for seed, n in ((1, 0), (2, 1), (3, 1500), (4, 6000)):
        k = random_keypoints(oracle_mod, n, W, H, seed)
        cs, idx = oracle_mod.assign_grid(k, 0.0, 0.0, inv_w, inv_h)
        grid = literal_grid(k, 0.0, 0.0, inv_w, inv_h)
        assert cs[0] == 0 and cs[-1] == len(idx) == sum(len(c) for col in grid for c in col)
        for i in range(COLS):
            for j in range(ROWS):
                c = i * ROWS + j
                assert list(idx[cs[c]:cs[c + 1]]) == grid[i][j]


def test_undistorted_bounds_offset(oracle_mod):
    """Non-zero mnMinX / mnMinY (distorted cameras): same arithmetic with an offset."""
    W, H = 752, 480
    min_x, min_y, max_x, max_y = -11.3, -7.9, 770.2, 489.6
    inv_w, inv_h = float(F(COLS) / F(F(max_x) - F(min_x))), float(F(ROWS) / F(F(max_y) - F(min_y)))
    k = random_keypoints(oracle_mod, 2000, W, H, 9)
    cs, idx = oracle_mod.assign_grid(k, min_x, min_y, inv_w, inv_h)
    grid = literal_grid(k, min_x, min_y, inv_w, inv_h)
    for i in range(COLS):
        for j in range(ROWS):
            c = i * ROWS + j
            assert list(idx[cs[c]:cs[c + 1]]) == grid[i][j]


def test_features_in_area_matches_literal(oracle_mod):
    W, H = 1241, 376
    inv_w, inv_h = float(F(COLS) / F(W)), float(F(ROWS) / F(H))
    k = random_keypoints(oracle_mod, 3000, W, H, 5)
    cs, idx = oracle_mod.assign_grid(k, 0.0, 0.0, inv_w, inv_h)
    grid = literal_grid(k, 0.0, 0.0, inv_w, inv_h)
    rng = np.random.default_rng(6)
    for _ in range(300):
        x, y = rng.uniform(-60, W + 60), rng.uniform(-60, H + 60)
        r = float(rng.choice([3.0, 15.0, 40.0, 200.0]))
        lv = [(-1, -1), (0, -1), (2, -1), (-1, 3), (1, 4), (5, 2)][int(rng.integers(0, 6))]
        got = oracle_mod.features_in_area(k, cs, idx, 0.0, 0.0, inv_w, inv_h, COLS, ROWS, x, y, r, *lv)
        assert list(got) == literal_area(k, grid, 0.0, 0.0, inv_w, inv_h, x, y, r, *lv)


def test_assign_features_against_a_numpy_model(oracle_mod):
    """orc_assign_features (Frame::AssignFeatures, Frame.cc:759-977) against a few lines of numpy on random label images:
    label 0 -> background, 255 -> dropped, L -> first detection j with (id_j > 255 ? id_j - 255 : id_j) == L - 1; every class in
    ascending feature order.  (tests/test_oracle_reference.py pins the same routine to the reference's own function.)"""
    rng = np.random.default_rng(17)
    for trial in range(20):
        H, W = int(rng.integers(40, 200)), int(rng.integers(40, 300))
        n, J = int(rng.integers(0, 600)), int(rng.integers(0, 60))
        mask = rng.choice(np.array([0, 0, 0, 255, 1, 2, 3, 7, 100, 254], np.uint8), size=(H, W))
        k = np.zeros(n, oracle_mod.KEYPOINT_DTYPE)
        k["x"] = rng.uniform(0, W - 0.001, n).astype(np.float32); k["y"] = rng.uniform(0, H - 0.001, n).astype(np.float32)
        ids = rng.choice(np.array([-1, 0, 1, 2, 6, 99, 253, 255 + 2, 255 + 6, 400], np.int32), size=J)
        depth = rng.uniform(-5, 100, n).astype(np.float32)
        cls, groups, eff = oracle_mod.assign_features(k, mask, ids, depth=depth, depth_limit=50.0)
        lab = mask[k["y"].astype(np.int32), k["x"].astype(np.int32)].astype(np.int32) if n else np.zeros(0, np.int32)
        eff_id = np.where(ids > 255, ids - 255, ids)
        want = np.full(n, -1, np.int32)
        for i in range(n):
            if lab[i] == 0:
                want[i] = 0
            elif lab[i] != 255:
                hit = np.nonzero((ids >= 0) & (eff_id == lab[i] - 1))[0]
                if len(hit):
                    want[i] = 1 + hit[0]
        assert np.array_equal(cls, want)
        assert len(groups) == J + 1
        for c in range(J + 1):
            assert np.array_equal(groups[c], np.nonzero(want == c)[0])
        for j in range(J):
            sel = want == 1 + j
            assert eff[j] == int(((depth[sel] > 0) & (depth[sel] < 50.0)).sum())
