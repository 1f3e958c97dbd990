"""Randomised parity sweep of the oracle's "next rows" against the reference's own functions lifted verbatim into
oracle/_ref/libref_orb.so: Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea
(src/Frame.cc:1636-1655, 2027-2038, 1808-1861), the five ORBmatcher::SearchByProjection overloads (src/ORBmatcher.cc:68-155,
157-246, 413-527, 1613-1756, 2529-2654) and Frame::AssignFeatures (src/Frame.cc:759-977).  CPU only; test infrastructure.

Every case draws a seed, feature / map-point counts, grid origin offsets, search radii, level windows, nnratio, and - for the
partition - a random label image and detection list; the frame scenario builder is the one of tests/ (make_case).

    python tools/ref_fuzz_frame.py [cases=150] [seed=0] > profiles/r02_ref_fuzz_frame.txt
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle as O  # noqa: E402
from oracle import ref_path as RP  # noqa: E402
import test_oracle_projection as TP  # noqa: E402
import test_oracle_reference as TR  # noqa: E402

F = np.float32


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if not RP.available("stock"):
        sys.exit("oracle/_ref/libref_orb.so is not built (make -C oracle ref REF=/root/reference)")
    O.build()
    rng = np.random.default_rng(seed0)
    t0 = time.time()
    n_grid = n_area = n_proj = n_match = n_part = n_part_feat = n_other = n_other_match = bad = 0
    for case in range(cases):
        seed = int(rng.integers(0, 1 << 30))
        n_feat = int(rng.choice([200, 500, 800, 1200, 2000]))
        n_mp = int(rng.choice([50, 400, 900, 2500]))
        c = TP.make_case(O, seed, n_feat=n_feat, n_mp=n_mp)
        k = c["k"].copy()
        m = min(40, len(k) // 4)  # push some keypoints across the grid borders so that PosInGrid's reject branch runs
        k["x"][:m] = rng.uniform(-30, 5, m).astype(F); k["y"][m:2 * m] = rng.uniform(c["H"] - 5, c["H"] + 30, m).astype(F)
        mx, my = (0.0, 0.0) if case % 2 else (float(F(rng.uniform(-20, 20))), float(F(rng.uniform(-20, 20))))
        # ---- grid + area lookups (on the perturbed keypoints)
        cg = dict(c); cg["k"] = k
        fr = TR._ref_frame(RP, cg, mx, my)
        rcs, ridx = fr.grid()
        ocs, oidx = O.assign_grid(k, mx, my, c["inv_w"], c["inv_h"])
        n_grid += 1
        if not (np.array_equal(rcs, ocs) and np.array_equal(ridx, oidx)):
            bad += 1; print(f"MISMATCH case {case} (seed {seed}): grid", flush=True)
        for _ in range(100):
            x, y = float(rng.uniform(-40, c["W"] + 40)), float(rng.uniform(-40, c["H"] + 40))
            r = float(rng.choice([0.5, 3.0, 9.5, 25.0, 80.0, 400.0, float(rng.uniform(0.1, 200))]))
            lo, hi = [(-1, -1), (0, 3), (2, -1), (5, 7), (3, 2), (0, 0), (7, 7)][int(rng.integers(0, 7))]
            got = O.features_in_area(k, ocs, oidx, mx, my, c["inv_w"], c["inv_h"], 64, 48, x, y, r, lo, hi)
            n_area += 1
            if not np.array_equal(fr.features_in_area(x, y, r, lo, hi), got):
                bad += 1; print(f"MISMATCH case {case} (seed {seed}): area ({x}, {y}, {r}, {lo}, {hi})", flush=True)
        # ---- local-map SearchByProjection (on the frame as extracted)
        th = float(rng.choice([1.0, 3.0, 5.0, float(F(rng.uniform(0.5, 8)))]))
        nn = float(rng.choice([0.6, 0.7, 0.8, 0.9, float(F(rng.uniform(0.5, 1.0)))]))
        got, n = TP.run_oracle(O, c, th, nn)
        fr2 = TR._ref_frame(RP, c)
        rn, rslot = fr2.search_by_projection(c["mps"], c["md"], c["occupied"], th, nn)
        n_proj += 1; n_match += int(n)
        if rn != n or not np.array_equal(rslot, TR._replay_slots(len(c["k"]), c["occupied"], got)):
            bad += 1; print(f"MISMATCH case {case} (seed {seed}): search_by_projection th={th} nn={nn}: {n} vs {rn}", flush=True)
        # ---- the other SearchByProjection overloads (check functions of tests/test_oracle_reference.py, random arguments)
        sd = seed % 1000000
        for name, fn, args in (
                ("last-frame", TR._check_projection_last, (sd, float(rng.choice([0.0, 2.0, -2.0, 0.3, -0.3])), int(rng.random() < 0.3),
                                                           int(rng.random() < 0.7), float(rng.choice([7.0, 15.0, 3.0, 30.0])))),
                ("relocalisation", TR._check_projection_kf, (sd, int(rng.random() < 0.7), float(rng.choice([3.0, 10.0, 15.0, 25.0])),
                                                               int(rng.choice([50, 64, 80, 100])))),
                ("loop-closing", TR._check_projection_sim3, (sd, int(rng.choice([3, 4, 10, 15, 25])))),
                ("object-point", TR._check_projection_obj, (sd, float(rng.choice([1.0, 3.0, 5.0])), float(rng.choice([0.6, 0.8, 0.9]))))):
            try:
                r = fn(O, RP, *args)
                n_other += 1; n_other_match += int(r[0] if isinstance(r, tuple) else r)
            except AssertionError as e:
                bad += 1; print(f"MISMATCH case {case}: {name} SearchByProjection{args}: {e}", flush=True)
        # ---- Frame::AssignFeatures
        if case % 3 == 0:
            ok, nf = partition_case(seed % 100000, rng)
            n_part += 1; n_part_feat += nf
            if not ok:
                bad += 1; print(f"MISMATCH case {case} (seed {seed}): assign_features", flush=True)
    print(f"oracle vs the reference's lifted functions, randomised: {cases} frames (seed {seed0}): {n_grid} grids, {n_area} area lookups, "
          f"{n_proj} local-map projection searches with {n_match} matches, {n_other} searches of the four other SearchByProjection overloads "
          f"with {n_other_match} matches, {n_part} feature partitions over {n_part_feat} features, "
          f"{bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


def partition_case(seed, rng):
    """One Frame::AssignFeatures comparison on a random label image (rectangles of random labels incl. 255 = ignored) and a random
    detection list (ids above 255, null detections, duplicates, ids without a region); -> (equal, features)."""
    W, H = 1242, 375
    r = np.random.default_rng(seed)

    def feats(n):
        k = np.zeros(n, O.KEYPOINT_DTYPE)
        k["x"] = r.uniform(0, W - 0.01, n).astype(F); k["y"] = r.uniform(0, H - 0.01, n).astype(F)
        k["octave"] = r.integers(0, 8, n); k["size"] = 31; k["angle"] = r.uniform(0, 360, n).astype(F)
        d = r.integers(0, 256, (n, 32), dtype=np.uint8)
        ur = np.where(r.random(n) < 0.6, k["x"] - r.uniform(1, 60, n), -1).astype(F)
        dp = np.where(ur >= 0, r.uniform(0.5, 120, n), -1).astype(F)
        return k, d, ur, dp
    n_orb = int(rng.choice([0, 1, 300, 1500, 3000])); n_dense = int(rng.choice([0, 1, 200, 1000]))
    (k, d, ur, dp), (ok, od, our, odp) = feats(n_orb), feats(n_dense)
    mask = np.zeros((H, W), np.uint8)
    n_reg = int(r.integers(0, 14))
    labels = r.integers(1, 256, n_reg)
    for lab in labels:
        w, h = int(r.integers(10, 400)), int(r.integers(10, 300))
        x, y = int(r.integers(0, W - w)), int(r.integers(0, H - h))
        mask[y:y + h, x:x + w] = lab
    n_det = int(r.integers(0, 12))
    ids = []
    for _ in range(n_det):
        u = r.random()
        if u < 0.5 and n_reg:
            lab = int(r.choice(labels)) - 1           # the id whose label is in the image
            ids.append(lab + 255 if r.random() < 0.2 and lab > 0 else lab)   # mnObjectID > 255 -> id - 255
        elif u < 0.6:
            ids.append(-1)                            # null detection
        else:
            ids.append(int(r.integers(0, 600)))
    ids = np.array(ids, np.int32)
    th_depth = float(r.choice([35.0, 10.0, 60.0]))
    inv_w, inv_h = F(64) / F(W), F(48) / F(H)
    ref = RP.assign_features(k, d, ur, dp, ok, od, our, odp, mask, ids, th_depth, 0.0, float(W), 0.0, float(H), float(inv_w),
                             float(inv_h))
    cls, groups, _ = O.assign_features(k, mask, ids)
    ocls, ogroups, eff = O.assign_features(ok, mask, ids, depth=odp, depth_limit=2 * th_depth)
    ok_ = np.array_equal(ref["background"], groups[0])
    kept = [j for j in range(len(ids)) if ids[j] >= 0]
    ok_ &= ref["det_of"].tolist() == kept
    if ok_:
        for kk, j in enumerate(kept):
            ok_ &= np.array_equal(ref["ori"][kk], groups[1 + j]) and np.array_equal(ref["obj"][kk], ogroups[1 + j])
            ok_ &= ref["init_n"][kk] == len(ogroups[1 + j])
            cs, idx = O.assign_grid(ok[ogroups[1 + j]], 0.0, 0.0, inv_w, inv_h)
            ok_ &= np.array_equal(ref["grid"][kk][0], cs) and np.array_equal(ref["grid"][kk][1], idx)
            ok_ &= eff[j] == int(((odp[ogroups[1 + j]] > 0) & (odp[ogroups[1 + j]] < 2 * th_depth)).sum())
    return bool(ok_), n_orb + n_dense


if __name__ == "__main__":
    main()
