"""CPU known-answer tests: the oracle's restated OpenCV primitives vs OpenCV 4.13 itself (cv2),
and the whole C oracle vs the cv2-in-the-loop path (oracle/cv2_path.py) on a small image."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _chain_sizes(w, h):
    out, sf = [], np.float32(1.0)
    for _ in range(7):
        sf = np.float32(float(sf) * float(np.float32(1.2)))
        inv = np.float32(1.0) / sf
        out.append((int(np.rint(np.float64(np.float32(w) * inv))), int(np.rint(np.float64(np.float32(h) * inv)))))
    return out


@pytest.mark.parametrize("wh", [(1242, 375), (752, 480), (1241, 376), (333, 222), (640, 480)])
def test_resize_linear(wh, oracle_mod):
    w, h = wh
    cur = np.random.default_rng(w).integers(0, 256, (h, w), dtype=np.uint8)
    for dw, dh in _chain_sizes(w, h):
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, oracle_mod.resize_linear(cur, dw, dh))
        cur = ref


@pytest.mark.parametrize("wh", [(1242, 375), (347, 105), (64, 9), (7, 7), (13, 200)])
def test_gaussian_blur(wh, oracle_mod):
    w, h = wh
    img = np.random.default_rng(h).integers(0, 256, (h, w), dtype=np.uint8)
    ref = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, oracle_mod.gaussian7x7(img))


@pytest.mark.parametrize("seed", range(8))
def test_fast_9_16(seed, oracle_mod):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (70, 90), dtype=np.uint8)
    if seed % 2:
        img = cv2.GaussianBlur(img, (0, 0), 1.0 + 0.3 * seed)
        img = np.clip(img.astype(np.int32) * 3 - 256, 0, 255).astype(np.uint8)
    for th in (20, 7, 5):
        det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True)
        ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)]
        x, y, r = oracle_mod.fast9_16(img, th)
        assert ref == list(zip(x.tolist(), y.tolist(), r.tolist()))


def test_fast_atan2(oracle_mod):
    rng = np.random.default_rng(3)
    ys = rng.integers(-70000, 70000, 20000).astype(np.float32)
    xs = rng.integers(-70000, 70000, 20000).astype(np.float32)
    for yv, xv in list(zip(ys, xs)) + [(0, 0), (1, 0), (-1, 0), (0, 1), (0, -1), (5, 5), (-5, 5)]:
        assert cv2.fastAtan2(float(yv), float(xv)) == oracle_mod.fast_atan2(yv, xv)


def test_cv_round(oracle_mod):
    for v in (0.5, 1.5, 2.5, -0.5, -1.5, 2.4999, 1e6 + 0.5, -2.5):
        assert oracle_mod.cv_round(v) == int(np.rint(v))


def test_c_oracle_equals_cv2_path(oracle_mod):
    """Full path, small image: C restatement == cv2 primitives + literal Python glue."""
    from oracle.cv2_path import Cv2Extractor, stereo_match_cv2
    from pointslot_b200.synth import synth_pair

    L, R = synth_pair(400, 240, 11)
    params = (600, 1.2, 8, 20, 7)
    eL, eR = oracle_mod.OracleExtractor(*params), oracle_mod.OracleExtractor(*params)
    cL, cR = Cv2Extractor(*params), Cv2Extractor(*params)
    kL, dL = eL.extract(L); kR, dR = eR.extract(R)
    kL2, dL2 = cL.extract(L); kR2, dR2 = cR.extract(R)
    assert kL.tobytes() == kL2.tobytes() and kR.tobytes() == kR2.tobytes()
    assert np.array_equal(dL, dL2) and np.array_equal(dR, dR2)
    for maxD in (300.0, float("inf")):
        ur, dp = oracle_mod.stereo_match(eL, eR, kL, dL, kR, dR, 40.0, maxD)
        ur2, dp2 = stereo_match_cv2(cL, cR, kL2, dL2, kR2, dR2, 40.0, maxD)
        assert ur.tobytes() == ur2.tobytes() and dp.tobytes() == dp2.tobytes()
        assert (ur >= 0).sum() > 10


def test_cvt_gray_exhaustive(oracle_mod):
    """All 2^24 colours, both channel orders and the alpha variants, vs cv2.cvtColor (OpenCV 4.13 arithmetic)."""
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(v & 0xFF), (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(oracle_mod.cvt_gray(img, True), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(oracle_mod.cvt_gray(img, False), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    img4 = np.concatenate([img[:512], np.full((512, 4096, 1), 200, np.uint8)], axis=2)
    assert np.array_equal(oracle_mod.cvt_gray(img4, True), cv2.cvtColor(img4, cv2.COLOR_RGBA2GRAY))
    assert np.array_equal(oracle_mod.cvt_gray(img4, False), cv2.cvtColor(img4, cv2.COLOR_BGRA2GRAY))


def test_glibc_sincosf_restatement_equals_libm(oracle_mod):
    """orc_glibc_cosf / orc_glibc_sinf (the steering trigonometry, ORBextractor.cc:112-113 -> std::cos(float) = glibc cosf)
    against this host's libm.  The exhaustive sweep over every float in [0, 2 pi] is profiles/r02_sincosf_sweep.txt."""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(0, 2 * np.pi, 60000), np.deg2rad(np.arange(0, 360, 0.25)),
                        [0.0, 1e-5, 2.4e-4, 0.7853981, 0.7853982, 0.78539824, 1.5707963, 3.1415927, 4.712389, 6.2831855]]).astype(np.float32)
    c, s = oracle_mod.glibc_sincosf(x)
    lc, ls = oracle_mod.libm_sincosf(x)
    assert c.tobytes() == lc.tobytes() and s.tobytes() == ls.tobytes()
    # and the angles the path really feeds it: fastAtan2 output (degrees) times the float pi/180
    ang = rng.uniform(0, 360, 20000).astype(np.float32) * np.float32(np.pi / 180.0)
    c, s = oracle_mod.glibc_sincosf(ang); lc, ls = oracle_mod.libm_sincosf(ang)
    assert c.tobytes() == lc.tobytes() and s.tobytes() == ls.tobytes()
