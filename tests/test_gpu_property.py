"""GPU property tests (hypothesis): random small images, shapes, and extractor parameters — the CUDA path must
equal the oracle byte for byte on every draw (ragged sizes, tiny levels, saturated / flat / noisy content)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu


def _image(kind, h, w, seed):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "checker":
        p = int(rng.integers(3, 17))
        yy, xx = np.mgrid[:h, :w]
        img = (((yy // p) + (xx // p)) % 2 * int(rng.integers(40, 255))).astype(np.uint8)
        return np.clip(img.astype(np.int32) + rng.integers(-3, 4, (h, w)), 0, 255).astype(np.uint8)
    if kind == "blobs":
        img = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        for _ in range(int(rng.integers(5, 60))):
            y, x = int(rng.integers(0, h - 8)), int(rng.integers(0, w - 8))
            img[y:y + int(rng.integers(2, 40)), x:x + int(rng.integers(2, 40))] = int(rng.integers(0, 256))
        return img
    if kind == "saturated":
        img = rng.integers(0, 2, (h, w), dtype=np.uint8) * 255
        img[: h // 2] = 255
        return img
    g = np.linspace(0, 255, w)[None, :] * np.ones((h, 1))  # gradient + sparse spikes
    img = g.astype(np.uint8)
    ys, xs = rng.integers(0, h, 200), rng.integers(0, w, 200)
    img[ys, xs] = rng.integers(0, 256, 200)
    return img


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(kind=st.sampled_from(["noise", "checker", "blobs", "saturated", "gradient"]),
       h=st.integers(180, 330), w=st.integers(180, 420), seed=st.integers(0, 10 ** 6),
       nfeatures=st.integers(50, 1500), nlevels=st.integers(1, 5), scale=st.sampled_from([1.1, 1.2, 1.25, 1.3, 1.4]),
       ini=st.integers(8, 40), dmin=st.integers(1, 7))
def test_extract_matches_oracle(kind, h, w, seed, nfeatures, nlevels, scale, ini, dmin):
    import oracle as O
    from pointslot_b200 import binding as B

    min_th = max(1, ini - dmin)
    # coarsest level must keep one 30-px cell per axis (62 px); skip draws the reference would divide by zero on
    s = scale ** (nlevels - 1)
    if min(h, w) / s < 64:
        nlevels = 1
    params = (nfeatures, scale, nlevels, ini, min_th)
    img = _image(kind, h, w, seed)
    e = B.OrbExtractor(*params)
    o = O.OracleExtractor(*params)
    try:
        k, d = e.extract(img)
    except B.PslotError as err:
        # shapes the reference itself cannot process (a level with < 1 cell per axis, or a portrait level whose
        # width/height rounds to 0 quadtree roots: division by zero in ORBextractor.cc:543-545, 784-787)
        assert err.code == -4
        with pytest.raises(RuntimeError):
            o.extract(img)
        e.close()
        return
    ok, od = o.extract(img)
    assert k.tobytes() == ok.tobytes(), (kind, h, w, seed, params, len(k), len(ok))
    assert np.array_equal(d, od)
    e.close()


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 10 ** 6), shift=st.integers(0, 40), nfeatures=st.integers(200, 1200),
       maxd=st.sampled_from([30.0, 100.0, 400.0, float("inf")]))
def test_stereo_matches_oracle(seed, shift, nfeatures, maxd):
    import oracle as O
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_pair

    L, _ = synth_pair(420, 260, seed % 1000)
    R = np.roll(L, -shift, axis=1)
    R = np.clip(R.astype(np.int32) + np.random.default_rng(seed).integers(-2, 3, R.shape), 0, 255).astype(np.uint8)
    params = (nfeatures, 1.2, 5, 20, 7)
    eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
    oL, oR = O.OracleExtractor(*params), O.OracleExtractor(*params)
    kL, dL = eL.extract(L); kR, dR = eR.extract(R)
    okL, odL = oL.extract(L); okR, odR = oR.extract(R)
    assert kL.tobytes() == okL.tobytes() and kR.tobytes() == okR.tobytes()
    ur, dp, bd, bi, bs = B.stereo_match(eL, eR, kL, dL, kR, dR, 120.0, maxd, debug=True)
    our, odp, obd, obi, obs = O.stereo_match(oL, oR, okL, odL, okR, odR, 120.0, maxd, debug=True)
    assert np.array_equal(bd, obd) and np.array_equal(bi, obi) and np.array_equal(bs, obs)
    assert np.array_equal(ur >= 0, our >= 0) and np.all(np.abs(ur - our) <= 1e-3)
    assert ur.tobytes() == our.tobytes() and dp.tobytes() == odp.tobytes()  # in practice bit-equal
    eL.close(); eR.close()


def test_cross_device_determinism(oracle_mod):
    """the same pairs give the same bytes on every GPU of the box (bench.py checks it across ranks; this is the
    in-process form).  Skips on a single-GPU box."""
    import hashlib
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_pair
    ndev = B.device_count()
    if ndev < 2:
        pytest.skip("needs at least two CUDA devices")
    W, H, P = 640, 360, 6
    params = (800, 1.2, 8, 20, 7)
    pairs = [synth_pair(W, H, 90 + s) for s in range(P)]
    Ls = np.stack([p[0] for p in pairs]); Rs = np.stack([p[1] for p in pairs])
    digests = []
    for dev in range(ndev):
        sb = B.StereoBatch(params, W, H, P, device=dev)
        res = sb.run(Ls, Rs, 200.0, B.max_disparity(200.0, 500.0))
        h = hashlib.sha256()
        for r in res:
            for key in ("kL", "dL", "kR", "dR", "u_right", "depth"):
                h.update(np.ascontiguousarray(r[key]).tobytes())
        digests.append(h.hexdigest())
        if dev == 0:
            o = oracle_mod.OracleExtractor(*params)
            ok, od = o.extract(Ls[0])
            assert res[0]["kL"].tobytes() == ok.tobytes() and np.array_equal(res[0]["dL"], od)
        # the single-frame path and the cv::ORB row on that device as well
        e = B.OrbExtractor(*params, device=dev)
        k, d = e.extract(Ls[1])
        assert k.tobytes() == res[1]["kL"].tobytes() and np.array_equal(d, res[1]["dL"])
        c = B.CvOrb(500, 1.2, 8, 19, device=dev)
        ck, cd = c.detect_and_compute(Ls[2])
        if dev == 0:
            first_cv = (ck.tobytes(), cd.tobytes())
        else:
            assert (ck.tobytes(), cd.tobytes()) == first_cv
    assert len(set(digests)) == 1, digests
