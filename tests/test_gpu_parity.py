"""GPU parity tests (run with -m gpu on the B200 box): CUDA path through the C ABI vs the oracle and
the golden vectors.  Bar (north_star): keypoints / octaves / angles / descriptors bit-exact; matched
pair set identical; sub-pixel uRight within 1e-3 px (TOL_PX below; in practice bit-equal)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
TOL_PX = 1e-3  # north_star: "stereo depth must fall within 1e-3 px of the reference's sub-pixel disparity"


@pytest.fixture(scope="module")
def B():
    from pointslot_b200 import binding
    assert binding.device_count() > 0
    return binding


def _assert_stereo(ur, dp, ur_ref, dp_ref):
    assert np.array_equal(ur >= 0, ur_ref >= 0), "matched-pair set differs"
    m = ur_ref >= 0
    assert np.all(np.abs(ur[m] - ur_ref[m]) <= TOL_PX)
    rel = np.abs(dp[m] - dp_ref[m]) / np.maximum(np.abs(dp_ref[m]), 1e-9)
    assert np.all(rel <= 1e-5)
    assert np.all(ur[~m] == -1) and np.all(dp[~m] == -1)


@pytest.mark.parametrize("name", golden_names())
def test_golden_end_to_end(name, B):
    g = load_golden(name)
    eL, eR = B.OrbExtractor(*g["params_t"]), B.OrbExtractor(*g["params_t"])
    kL, dL = eL.extract(g["left"]); kR, dR = eR.extract(g["right"])
    assert kL.tobytes() == g["kL"].tobytes() and kR.tobytes() == g["kR"].tobytes()
    assert np.array_equal(dL, g["dL"]) and np.array_equal(dR, g["dR"])
    ur, dp = B.stereo_match(eL, eR, kL, dL, kR, dR, float(g["bf"]), float(g["maxD"]))
    _assert_stereo(ur, dp, g["uright"], g["depth"])


@pytest.mark.parametrize("name", ["odd_synth_s5_500", "kitti_natural_2000"])
def test_stage_parity(name, B, oracle_mod):
    """pyramid -> FAST candidates -> quadtree keypoints -> angles, each against the oracle."""
    g = load_golden(name)
    o = oracle_mod.OracleExtractor(*g["params_t"]); o.extract(g["left"])
    e = B.OrbExtractor(*g["params_t"]); e.extract(g["left"])
    for l in range(o.nlevels):
        assert np.array_equal(e.level(l), o.level(l)), f"pyramid level {l}"
        for a, b in zip(e.debug_candidates(l), o.level_candidates(l)):
            assert np.array_equal(a, b), f"candidates level {l}"
        for a, b in zip(e.debug_level_keypoints(l), o.level_keypoints(l)):
            assert np.array_equal(a, b), f"keypoints level {l}"


@pytest.mark.parametrize("shape,params", [((375, 1242), (1000, 1.2, 8, 20, 5)), ((480, 752), (1200, 1.2, 8, 20, 7)),
                                         ((376, 1241), (2000, 1.2, 8, 20, 7)), ((300, 400), (800, 1.3, 6, 15, 5))])
def test_random_seeds_vs_oracle(shape, params, B, oracle_mod):
    from pointslot_b200.synth import synth_pair
    H, W = shape
    bf, fx = 384.38148, 721.5377
    maxD = B.max_disparity(bf, fx)
    eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
    oL, oR = oracle_mod.OracleExtractor(*params), oracle_mod.OracleExtractor(*params)
    for seed in (100, 101, 102):
        L, R = synth_pair(W, H, seed)
        kL, dL = eL.extract(L); kR, dR = eR.extract(R)
        okL, odL = oL.extract(L); okR, odR = oR.extract(R)
        assert kL.tobytes() == okL.tobytes() and kR.tobytes() == okR.tobytes()
        assert np.array_equal(dL, odL) and np.array_equal(dR, odR)
        ur, dp = B.stereo_match(eL, eR, kL, dL, kR, dR, bf, maxD)
        our, odp = oracle_mod.stereo_match(oL, oR, okL, odL, okR, odR, bf, maxD)
        _assert_stereo(ur, dp, our, odp)
        ur2, dp2 = B.stereo_match(eL, eR, kL, dL, kR, dR, bf, float("inf"))  # the reference's mb == 0 case
        our2, odp2 = oracle_mod.stereo_match(oL, oR, okL, odL, okR, odR, bf, float("inf"))
        _assert_stereo(ur2, dp2, our2, odp2)


def test_batch_equals_single_and_oracle(B, oracle_mod):
    from pointslot_b200.synth import synth_batch
    params = (1000, 1.2, 8, 20, 5)
    W, H, n = 1242, 375, 6
    bf, fx = 384.38148, 721.5377
    maxD = B.max_disparity(bf, fx)
    Ls, Rs = synth_batch(W, H, range(40, 40 + n))
    sb = B.StereoBatch(params, W, H, 8)
    res = sb.run(Ls, Rs, bf, maxD)
    res_again = sb.run(Ls[:3], Rs[:3], bf, maxD)  # smaller batch through the same context
    oL, oR = oracle_mod.OracleExtractor(*params), oracle_mod.OracleExtractor(*params)
    for i in range(n):
        okL, odL = oL.extract(Ls[i]); okR, odR = oR.extract(Rs[i])
        our, odp = oracle_mod.stereo_match(oL, oR, okL, odL, okR, odR, bf, maxD)
        r = res[i]
        assert r["kL"].tobytes() == okL.tobytes() and r["kR"].tobytes() == okR.tobytes()
        assert np.array_equal(r["dL"], odL) and np.array_equal(r["dR"], odR)
        _assert_stereo(r["u_right"], r["depth"], our, odp)
        if i < 3:
            for k in r:
                assert r[k].tobytes() == res_again[i][k].tobytes()


def test_edge_cases(B, oracle_mod):
    e = B.OrbExtractor(500, 1.2, 8, 20, 7)
    k, d = e.extract(np.zeros((0, 0), np.uint8))  # empty image: reference returns without touching outputs
    assert len(k) == 0
    k, d = e.extract(np.full((400, 600), 77, np.uint8))  # featureless
    assert len(k) == 0 and d.shape == (0, 32)
    ur, dp = B.stereo_match(e, e, k, d, k, d, 40.0, 300.0)
    assert len(ur) == 0
    with pytest.raises(B.PslotError) as ei:
        e.extract(np.zeros((120, 160), np.uint8))  # coarsest level < 62 px: reference divides by zero
    assert ei.value.code == -4
    # non-contiguous rows (honour `pitch`)
    from pointslot_b200.synth import synth_pair
    L, _ = synth_pair(500, 320, 9)
    big = np.zeros((320, 640), np.uint8); big[:, :500] = L
    k1, d1 = e.extract(L); k2, d2 = e.extract(big[:, :500])
    assert k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)
    o = oracle_mod.OracleExtractor(500, 1.2, 8, 20, 7)
    ok, od = o.extract(L)
    assert k1.tobytes() == ok.tobytes() and np.array_equal(d1, od)
    # capacity error reports the needed count
    import ctypes as C
    n = C.c_int(0)
    kp = np.zeros(4, B.KEYPOINT_DTYPE); ds = np.zeros((4, 32), np.uint8)
    rc = B.lib().pslot_orb_extract(e._h, L.ctypes.data_as(C.c_void_p), 500, 320, 500, kp.ctypes.data_as(C.c_void_p),
                                   ds.ctypes.data_as(C.c_void_p), 4, C.byref(n))
    assert rc == -2 and n.value == len(k1)


def test_noise_image_many_candidates(B, oracle_mod):
    """Uniform noise: FAST fires almost everywhere -> stresses cell capacity, quadtree spill, ties."""
    img = np.random.default_rng(5).integers(0, 256, (240, 320), dtype=np.uint8)
    params = (1500, 1.2, 4, 20, 7)
    e = B.OrbExtractor(*params); o = oracle_mod.OracleExtractor(*params)
    k, d = e.extract(img); ok, od = o.extract(img)
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)


@pytest.mark.parametrize("params", [(7000, 1.2, 8, 12, 5),    # right set doubled to > 12.4 k keys: no room for row-bin lists
                                    (3000, 1.44, 7, 20, 7)])  # bands up to 2 * 2 * 1.44^6 rows: many bins per key
def test_stereo_search_fallbacks(params, B, oracle_mod):
    """The stereo search buckets right keys by 16-row bins in shared memory; both escape hatches of that scheme
    (lists dropped, lists overflowing) must give the same matches as the oracle's plain row table."""
    from pointslot_b200.synth import synth_pair
    rng = np.random.default_rng(11)
    W, H = (1241, 376) if params[0] > 5000 else (1241, 752)
    L, R = synth_pair(W, H, 77)
    if params[0] > 5000:  # add texture everywhere so that the quota is actually reached
        n = rng.integers(-40, 40, (H, W + 16))
        L = np.clip(L.astype(int) + n[:, 16:], 0, 255).astype(np.uint8)
        R = np.clip(R.astype(int) + n[:, 8:W + 8], 0, 255).astype(np.uint8)
    bf, maxD = 384.38148, B.max_disparity(384.38148, 721.5377)
    eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
    oL, oR = oracle_mod.OracleExtractor(*params), oracle_mod.OracleExtractor(*params)
    kL, dL = eL.extract(L); kR, dR = eR.extract(R)
    okL, odL = oL.extract(L); okR, odR = oR.extract(R)
    assert kL.tobytes() == okL.tobytes() and kR.tobytes() == okR.tobytes()
    if params[0] > 5000:
        # the matcher takes plain arrays: double the right set (copies shifted by a quarter pixel, same descriptors,
        # so every Hamming minimum is a tie that the first index must win)
        k2 = kR.copy(); k2["x"] += 0.25
        kR = okR = np.concatenate([kR, k2]); dR = odR = np.concatenate([dR, dR])
        assert len(kR) > 12500
    ur, dp = B.stereo_match(eL, eR, kL, dL, kR, dR, bf, maxD)
    our, odp = oracle_mod.stereo_match(oL, oR, okL, odL, okR, odR, bf, maxD)
    _assert_stereo(ur, dp, our, odp)
    assert (our > 0).sum() > 50
    if params[0] <= 5000:
        # same keypoints relabelled to the two coarsest octaves: every right band now spans 3-4 bins, which overflows the
        # 3-entries-per-key lists; the matcher takes keypoints as plain arrays, so both sides accept the relabelling
        # (border keypoints are dropped: at a coarser level their 11x11 SAD window would leave the image, which the
        # reference does not guard against because real keypoints never get there)
        def inner(k):
            return (k["x"] > 100) & (k["x"] < W - 100) & (k["y"] > 64) & (k["y"] < H - 64)
        mL, mR = inner(kL), inner(kR)
        kL2, kR2, dL, dR, odL, odR = kL[mL].copy(), kR[mR].copy(), dL[mL], dR[mR], odL[mL], odR[mR]
        kL2["octave"] = 5 + (np.arange(len(kL2)) & 1); kR2["octave"] = 6
        ur, dp = B.stereo_match(eL, eR, kL2, dL, kR2, dR, bf, maxD)
        our, odp = oracle_mod.stereo_match(oL, oR, kL2, odL, kR2, odR, bf, maxD)
        _assert_stereo(ur, dp, our, odp)
        assert (our > 0).sum() > 20


def test_steering_coefficients_vs_libm(B, oracle_mod):
    """a=(float)cos, b=(float)sin of angle*pi/180 (double trig, rounded): every 509th float in [0,360] + specials.
    (tools/steer_sweep.py runs all 1.14e9 angles; result recorded in profiles/.)"""
    hi = int(np.float32(360.0).view(np.uint32)) + 1
    bits = np.arange(0, hi, 509, dtype=np.uint32)
    ang = np.concatenate([bits.view(np.float32), np.arange(0, 361, dtype=np.float32),
                          np.float32([0.5, 44.99999, 45.0, 89.99999, 90.00001, 179.9999, 270.0001, 359.9999])])
    ga, gb = B.debug_steer(ang)
    oa, ob = oracle_mod.steer_coeffs(ang)
    assert np.array_equal(ga.view(np.uint32), oa.view(np.uint32))
    assert np.array_equal(gb.view(np.uint32), ob.view(np.uint32))


@pytest.mark.parametrize("channels,rgb", [(3, True), (3, False), (4, True), (4, False)])
def test_color_ingest(channels, rgb, B, oracle_mod):
    """cvtColor fused with the upload == oracle cvtColor followed by the gray path."""
    rng = np.random.default_rng(channels * 2 + rgb)
    from pointslot_b200.synth import synth_pair
    W, H = 641, 361  # odd width: the last 4-pixel group is partial, image starts are not word aligned
    base, _ = synth_pair(W, H, 31)
    col = np.stack([np.clip(base.astype(np.int32) + rng.integers(-30, 31, base.shape), 0, 255).astype(np.uint8)
                    for _ in range(channels)], axis=2)
    gray = oracle_mod.cvt_gray(col, rgb)
    params = (700, 1.2, 8, 20, 7)
    e = B.OrbExtractor(*params)
    k, d = e.extract(col, rgb=rgb)
    assert np.array_equal(e.level(0), gray)
    o = oracle_mod.OracleExtractor(*params)
    ok, od = o.extract(gray)
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    # batch path
    sb = B.StereoBatch(params, W, H, 2)
    out = sb.alloc_out(2)
    cols = np.ascontiguousarray(np.stack([col, col[:, ::-1].copy()]))
    sb.submit_color(cols, cols, rgb, 100.0, 300.0, out); sb.wait()
    n0 = int(out["counts"][0])
    assert out["kps"][0, :n0].tobytes() == ok.tobytes() and np.array_equal(out["desc"][0, :n0], od)
    ok1, od1 = o.extract(oracle_mod.cvt_gray(cols[1], rgb))
    n1 = int(out["counts"][1])
    assert out["kps"][1, :n1].tobytes() == ok1.tobytes() and np.array_equal(out["desc"][1, :n1], od1)


def test_object_subset_stereo(B, oracle_mod):
    """Frame::ComputeObjStereoMatches (Frame.cc:2318-2503) is ComputeStereoMatches on a keypoint
    subset reading the same pyramids -> pslot_stereo_match with explicit feature sets covers it."""
    g = load_golden("kitti_natural_2000")
    eL, eR = B.OrbExtractor(*g["params_t"]), B.OrbExtractor(*g["params_t"])
    oL, oR = oracle_mod.OracleExtractor(*g["params_t"]), oracle_mod.OracleExtractor(*g["params_t"])
    kL, dL = eL.extract(g["left"]); kR, dR = eR.extract(g["right"])
    oL.extract(g["left"]); oR.extract(g["right"])
    rng = np.random.default_rng(4)
    selL = np.sort(rng.choice(len(kL), 300, replace=False)); selR = np.sort(rng.choice(len(kR), 700, replace=False))
    ur, dp = B.stereo_match(eL, eR, kL[selL], dL[selL], kR[selR], dR[selR], float(g["bf"]), float(g["maxD"]))
    our, odp = oracle_mod.stereo_match(oL, oR, kL[selL], dL[selL], kR[selR], dR[selR], float(g["bf"]), float(g["maxD"]))
    _assert_stereo(ur, dp, our, odp)
    assert (ur >= 0).sum() > 20


def test_stereo_rejects_instead_of_truncating(B, oracle_mod):
    """Limits of the stereo kernel are errors, never silent clamps: too many right keypoints (the kernel stages them in
    shared memory and names them with 16 bits), out-of-range octaves on either side, non-finite coordinates."""
    from pointslot_b200.synth import synth_pair
    params = (800, 1.2, 8, 20, 7)
    L, R = synth_pair(640, 360, 9)
    eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
    kL, dL = eL.extract(L); kR, dR = eR.extract(R)
    bf, maxD = 200.0, B.max_disparity(200.0, 500.0)
    reps = 16001 // len(kR) + 1
    with pytest.raises(B.PslotError):
        B.stereo_match(eL, eR, kL, dL, np.tile(kR, reps), np.tile(dR, (reps, 1)), bf, maxD)
    bad = kR.copy(); bad["octave"][3] = 8
    with pytest.raises(B.PslotError):
        B.stereo_match(eL, eR, kL, dL, bad, dR, bf, maxD)
    bad = kR.copy(); bad["octave"][0] = -1
    with pytest.raises(B.PslotError):
        B.stereo_match(eL, eR, kL, dL, bad, dR, bf, maxD)
    bad = kL.copy(); bad["x"][1] = np.nan
    with pytest.raises(B.PslotError):
        B.stereo_match(eL, eR, bad, dL, kR, dR, bf, maxD)
    bad = kR.copy(); bad["y"][1] = np.inf
    with pytest.raises(B.PslotError):
        B.stereo_match(eL, eR, kL, dL, bad, dR, bf, maxD)
    # the context is still usable afterwards and gives the oracle's answer
    oL, oR = oracle_mod.OracleExtractor(*params), oracle_mod.OracleExtractor(*params)
    oL.extract(L); oR.extract(R)
    ur, dp = B.stereo_match(eL, eR, kL, dL, kR, dR, bf, maxD)
    our, odp = oracle_mod.stereo_match(oL, oR, kL, dL, kR, dR, bf, maxD)
    _assert_stereo(ur, dp, our, odp)


def test_stereo_work_counters_match_the_row_table(B):
    """pslot_batch_stereo_counters: c_row = sum of |vRowIndices[row]| over the left keypoints that search, c-bar = the
    descriptor comparisons (Frame.cc:2188-2217), recomputed here from the returned keypoints."""
    from pointslot_b200.synth import synth_pair
    params = (1000, 1.2, 8, 20, 5)
    W, H, n = 1242, 375, 3
    Ls, Rs = zip(*[synth_pair(W, H, 40 + i) for i in range(n)])
    bf, maxD = 384.38148, B.max_disparity(384.38148, 721.5377)
    sb = B.StereoBatch(params, W, H, n)
    sb.set_profiling(True)
    res = sb.run(np.stack(Ls), np.stack(Rs), bf, maxD)
    got = sb.stereo_counters()
    sf = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    sfs = [np.float32(1.0)]
    for _ in range(7):
        sfs.append(np.float32(np.float64(sfs[-1]) * np.float64(np.float32(1.2))))
    sf = np.array(sfs, np.float32)
    rows = cmps = 0
    for r in res:
        kL, kR = r["kL"], r["kR"]
        rad = np.float32(2.0) * sf[kR["octave"]]
        maxr = np.ceil(kR["y"] + rad).astype(int); minr = np.floor(kR["y"] - rad).astype(int)
        for k in kL:
            row = int(k["y"])
            band = (minr <= row) & (row <= maxr)
            if k["x"] < 0:
                continue
            rows += int(band.sum())
            minU = np.float32(k["x"]) - np.float32(maxD)
            ok = band & (np.abs(kR["octave"] - k["octave"]) <= 1) & (kR["x"] >= minU) & (kR["x"] <= k["x"])
            cmps += int(ok.sum())
    assert got["row_candidates"] == rows and got["hamming_evals"] == cmps
    assert sb.stereo_counters() == {"row_candidates": 0, "hamming_evals": 0}  # reset by the read


def test_extract_in_two_halves(oracle_mod):
    """pslot_orb_extract_begin / _end: both views in flight from one host thread, frame after frame (the second frame of a
    geometry onwards replays a CUDA graph), other sizes in between, an empty image - always the bytes of the oracle."""
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_pair
    params = (1000, 1.2, 8, 20, 7)
    eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
    oe = oracle_mod.OracleExtractor(*params)
    for W, H, seed in ((640, 360, 1), (640, 360, 2), (640, 360, 3), (403, 301, 4), (640, 360, 5)):
        L, R = synth_pair(W, H, seed)
        eL.extract_begin(L); eR.extract_begin(R)
        kL, dL = eL.extract_end(); kR, dR = eR.extract_end()
        okL, odL = oe.extract(L); okR, odR = oe.extract(R)
        assert kL.tobytes() == okL.tobytes() and np.array_equal(dL, odL), (W, H, seed)
        assert kR.tobytes() == okR.tobytes() and np.array_equal(dR, odR), (W, H, seed)
        k2, d2 = eL.extract(L)  # the one-call form on the same context
        assert k2.tobytes() == okL.tobytes() and np.array_equal(d2, odL)
    big = np.zeros((360, 768), np.uint8); big[:, :640] = L   # a pitched view takes the 2-D upload
    eL.extract_begin(big[:, :640])
    k3, d3 = eL.extract_end()
    assert k3.tobytes() == okL.tobytes() and np.array_equal(d3, odL)
    eL.extract_begin(np.zeros((0, 0), np.uint8))
    k0, d0 = eL.extract_end()
    assert len(k0) == 0 and len(d0) == 0


def test_stereo_match_on_resident_feature_sets(oracle_mod):
    """pslot_stereo_match_resident: the matcher on the feature sets the two extractions left on the device equals the general
    form with the same arrays uploaded (and the oracle); stale tokens are refused."""
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_pair
    params = (1500, 1.2, 8, 20, 7)
    eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
    oL, oR = oracle_mod.OracleExtractor(*params), oracle_mod.OracleExtractor(*params)
    bf, maxD = 386.1448, B.max_disparity(386.1448, 718.856)
    for seed in (11, 12):
        L, R = synth_pair(1242, 375, seed)
        eL.extract_begin(L); eR.extract_begin(R)
        kL, dL = eL.extract_end(); kR, dR = eR.extract_end()
        tl, tr = B.extract_token(eL), B.extract_token(eR)
        assert tl and tr and tl != tr
        ur, dp = B.stereo_match_resident(eL, eR, tl, tr, bf, maxD)
        ur2, dp2 = B.stereo_match(eL, eR, kL, dL, kR, dR, bf, maxD)
        assert len(ur) == len(kL) and np.array_equal(ur.view(np.uint32), ur2.view(np.uint32)) and np.array_equal(dp.view(np.uint32), dp2.view(np.uint32))
        okL, odL = oL.extract(L); okR, odR = oR.extract(R)
        our, odp = oracle_mod.stereo_match(oL, oR, okL, odL, okR, odR, bf, maxD)
        assert np.array_equal(ur >= 0, our >= 0) and np.all(np.abs(ur - our) <= 1e-3) and (ur >= 0).sum() > 50
    eR.extract(R)  # the right extractor moves on: the old token no longer names what is on the device
    with pytest.raises(B.PslotError):
        B.stereo_match_resident(eL, eR, tl, tr, bf, maxD)
