"""Oracle for ORBmatcher::SearchByProjection (local-map variant) against a literal Python
transcription of reference src/ORBmatcher.cc:68-155 (+ RadiusByViewingCos :252-258) on synthetic tracking scenarios."""
import numpy as np

from test_oracle_grid import COLS, ROWS, literal_area, literal_grid

F = np.float32
TH_HIGH = 100


def _flip_bits(rng, desc, flips):
    """Descriptors with about flips[i] random bits of row i inverted."""
    mask = rng.random((len(desc), 256)) < (np.asarray(flips, np.float64) / 256.0)[:, None]
    return desc ^ np.packbits(mask, axis=1)


def make_case(oracle_mod, seed, n_feat=1200, n_mp=1800):
    """A frame (extracted by the oracle, with stereo) and map points that re-observe its features with noise:
    duplicated sources (conflicts), perturbed descriptors, unrelated points, disabled points, pre-occupied features."""
    from pointslot_b200.synth import synth_pair
    W, H = 1242, 375
    rng = np.random.default_rng(seed)
    L, R = synth_pair(W, H, 300 + seed)
    params = (n_feat, 1.2, 8, 20, 7)
    oL, oR = oracle_mod.OracleExtractor(*params), oracle_mod.OracleExtractor(*params)
    k, d = oL.extract(L); kR, dR = oR.extract(R)
    ur, _ = oracle_mod.stereo_match(oL, oR, k, d, kR, dR, 384.38148, 721.5377)
    n = len(k)
    sf = (F(1.2) ** np.arange(8, dtype=F)).astype(F)
    src = rng.integers(0, n, n_mp)
    mps = np.zeros(n_mp, oracle_mod.MAP_POINT_DTYPE)
    mps["proj_x"] = k["x"][src] + rng.normal(0, 2.0, n_mp).astype(F)
    mps["proj_y"] = k["y"][src] + rng.normal(0, 2.0, n_mp).astype(F)
    mps["proj_xr"] = np.where(ur[src] > 0, ur[src] + rng.normal(0, 2.0, n_mp), mps["proj_x"] - 8).astype(F)
    mps["view_cos"] = rng.choice(np.array([0.9995, 0.9981, 0.998, 0.97, 0.4], F), n_mp)
    mps["level"] = np.clip(k["octave"][src] + rng.integers(0, 2, n_mp), 0, 7)
    mps["flags"] = (rng.random(n_mp) < 0.92) * 1 + (rng.random(n_mp) < 0.8) * 2
    md = _flip_bits(rng, d[src], rng.integers(0, 60, n_mp))
    stray = rng.random(n_mp) < 0.12  # unrelated map points
    mps["proj_x"][stray] = rng.uniform(-20, W + 20, stray.sum()); mps["proj_y"][stray] = rng.uniform(-20, H + 20, stray.sum())
    md[stray] = rng.integers(0, 256, (int(stray.sum()), 32), dtype=np.uint8)
    occupied = (rng.random(n) < 0.1).astype(np.uint8)
    inv_w, inv_h = float(F(COLS) / F(W)), float(F(ROWS) / F(H))
    return dict(k=k, d=d, ur=ur, sf=sf, mps=mps, md=md, occupied=occupied, inv_w=inv_w, inv_h=inv_h, W=W, H=H)


def literal_search(c, th, nn_ratio):
    k, d, ur, sf, mps, md = c["k"], c["d"], c["ur"], c["sf"], c["mps"], c["md"]
    grid = literal_grid(k, 0.0, 0.0, c["inv_w"], c["inv_h"])
    blocked = c["occupied"].astype(bool).copy()
    pc = np.unpackbits(d, axis=1)
    out = np.full(len(mps), -1, np.int32)
    nmatches = 0
    for i, p in enumerate(mps):
        if not (p["flags"] & 1):
            continue
        lvl = int(p["level"])
        r = F(2.5) if float(p["view_cos"]) > 0.998 else F(4.0)
        if th != 1.0:
            r = F(r * F(th))
        rad = F(r * sf[lvl])
        cand = literal_area(k, grid, 0.0, 0.0, c["inv_w"], c["inv_h"], p["proj_x"], p["proj_y"], rad, lvl - 1, lvl)
        if not cand:
            continue
        mbits = np.unpackbits(md[i])
        best, best_lv, best2, best_lv2, best_idx = 256, -1, 256, -1, -1
        for idx in cand:
            if blocked[idx]:
                continue
            if ur[idx] > 0:
                if abs(F(p["proj_xr"] - ur[idx])) > F(r * sf[lvl]):
                    continue
            dist = int((mbits != pc[idx]).sum())
            if dist < best:
                best2, best, best_lv2, best_lv, best_idx = best, dist, best_lv, int(k["octave"][idx]), idx
            elif dist < best2:
                best_lv2, best2 = int(k["octave"][idx]), dist
        if best <= TH_HIGH:
            if best_lv == best_lv2 and F(best) > F(F(nn_ratio) * F(best2)):
                continue
            out[i] = best_idx
            if p["flags"] & 2:
                blocked[best_idx] = True
            nmatches += 1
    return out, nmatches


def run_oracle(oracle_mod, c, th, nn_ratio):
    cs, idx = oracle_mod.assign_grid(c["k"], 0.0, 0.0, c["inv_w"], c["inv_h"])
    return oracle_mod.search_by_projection(c["k"], c["d"], c["ur"], cs, idx, 0.0, 0.0, c["inv_w"], c["inv_h"], COLS, ROWS,
                                           c["mps"], c["md"], c["occupied"], c["sf"], th, nn_ratio)


def test_search_by_projection_matches_literal(oracle_mod):
    for seed, th, nn in ((1, 1.0, 0.8), (2, 3.0, 0.8), (3, 5.0, 0.6)):
        c = make_case(oracle_mod, seed, n_feat=800, n_mp=900)
        got, n = run_oracle(oracle_mod, c, th, nn)
        ref, nref = literal_search(c, th, nn)
        assert np.array_equal(got, ref) and n == nref
        assert n > 100
        # the order dependence is exercised: some feature was wanted by more than one map point
        m = got[got >= 0]
        assert len(np.unique(m)) <= len(m)


# ---- last-frame variant (ORBmatcher.cc:1613-1756, from the GetFeaturesInArea call on) -------------------------------
def make_last_frame_case(oracle_mod, seed, n_feat=1200, mode="neutral"):
    """Current frame = oracle features of a synthetic image; 'last frame' points = the same features moved by a small
    flow, with the level window the reference picks for forward / backward / neutral motion."""
    c = make_case(oracle_mod, seed, n_feat=n_feat, n_mp=10)
    rng = np.random.default_rng(1000 + seed)
    k, d, ur = c["k"], c["d"], c["ur"]
    n = len(k)
    n_pts = n + n // 3
    src = np.concatenate([np.arange(n), rng.integers(0, n, n_pts - n)])  # every feature once, a third twice (conflicts)
    rng.shuffle(src)
    pts = np.zeros(n_pts, oracle_mod.PROJ_POINT_DTYPE)
    pts["u"] = k["x"][src] + rng.normal(0, 3.0, n_pts).astype(F)
    pts["v"] = k["y"][src] + rng.normal(0, 3.0, n_pts).astype(F)
    pts["ur"] = np.where(ur[src] > 0, ur[src] + rng.normal(0, 3.0, n_pts), pts["u"] - 12).astype(F)
    octv = k["octave"][src]
    th = F(7.0 if seed % 2 else 15.0)
    pts["radius"] = (th * c["sf"][octv]).astype(F)
    if mode == "forward":
        pts["min_level"], pts["max_level"] = octv, -1
    elif mode == "backward":
        pts["min_level"], pts["max_level"] = 0, octv
    else:
        pts["min_level"], pts["max_level"] = octv - 1, octv + 1
    # last-frame angles: mostly a common rotation of the current ones (one dominant histogram bin), some random
    ang = (k["angle"][src] + F(11.0)) % F(360.0)
    rnd = rng.random(n_pts) < 0.25
    ang[rnd] = rng.uniform(0, 360, rnd.sum()).astype(F)
    pts["angle"] = ang.astype(F)
    pts["flags"] = (rng.random(n_pts) < 0.9) * 1 + (rng.random(n_pts) < 0.85) * 2
    pd = _flip_bits(rng, d[src], rng.integers(0, 50, n_pts))
    c.update(pts=pts, pd=pd)
    return c


def literal_search_last(c, check_orientation=True):
    k, d, ur, pts, pd = c["k"], c["d"], c["ur"], c["pts"], c["pd"]
    grid = literal_grid(k, 0.0, 0.0, c["inv_w"], c["inv_h"])
    blocked = c["occupied"].astype(bool).copy()
    pc = np.unpackbits(d, axis=1)
    HISTO = 30
    factor = F(HISTO / F(360.0))
    hist = [[] for _ in range(HISTO)]
    match = np.full(len(pts), -1, np.int32)
    rejected = np.zeros(len(pts), np.uint8)
    nmatches = 0
    for i, p in enumerate(pts):
        if not (p["flags"] & 1):
            continue
        cand = literal_area(k, grid, 0.0, 0.0, c["inv_w"], c["inv_h"], p["u"], p["v"], p["radius"], int(p["min_level"]),
                            int(p["max_level"]))
        if not cand:
            continue
        bits = np.unpackbits(pd[i])
        best, best_idx = 256, -1
        for i2 in cand:
            if blocked[i2]:
                continue
            if ur[i2] > 0 and abs(F(p["ur"] - ur[i2])) > p["radius"]:
                continue
            dist = int((bits != pc[i2]).sum())
            if dist < best:
                best, best_idx = dist, i2
        if best <= TH_HIGH:
            match[i] = best_idx
            if p["flags"] & 2:
                blocked[best_idx] = True
            nmatches += 1
            if check_orientation:
                rot = F(p["angle"] - k["angle"][best_idx])
                if rot < 0.0:
                    rot = F(rot + F(360.0))
                v = float(F(rot * factor))
                b = int(np.floor(v + 0.5))  # C round() for v >= 0
                if b == HISTO:
                    b = 0
                hist[b].append(i)
    if check_orientation:
        max1 = max2 = max3 = 0
        ind1 = ind2 = ind3 = -1
        for i in range(HISTO):
            s = len(hist[i])
            if s > max1:
                max3, max2, max1, ind3, ind2, ind1 = max2, max1, s, ind2, ind1, i
            elif s > max2:
                max3, max2, ind3, ind2 = max2, s, ind2, i
            elif s > max3:
                max3, ind3 = s, i
        if max2 < F(0.1) * F(max1):
            ind2 = ind3 = -1
        elif max3 < F(0.1) * F(max1):
            ind3 = -1
        for i in range(HISTO):
            if i not in (ind1, ind2, ind3):
                for j in hist[i]:
                    rejected[j] = 1
                    nmatches -= 1
    return match, rejected, nmatches


def run_oracle_last(oracle_mod, c, check_orientation=True):
    cs, idx = oracle_mod.assign_grid(c["k"], 0.0, 0.0, c["inv_w"], c["inv_h"])
    return oracle_mod.search_by_projection_last(c["k"], c["d"], c["ur"], cs, idx, 0.0, 0.0, c["inv_w"], c["inv_h"], COLS,
                                                ROWS, c["pts"], c["pd"], c["occupied"], check_orientation)


def test_search_by_projection_last_matches_literal(oracle_mod):
    for seed, mode, ori in ((1, "neutral", True), (2, "forward", True), (3, "backward", True), (4, "neutral", False)):
        c = make_last_frame_case(oracle_mod, seed, n_feat=700, mode=mode)
        m, rej, n = run_oracle_last(oracle_mod, c, ori)
        rm, rrej, rn = literal_search_last(c, ori)
        assert np.array_equal(m, rm) and np.array_equal(rej, rrej) and n == rn
        assert n > 100 and (not ori or rej.sum() > 0)
