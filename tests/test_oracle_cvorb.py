"""oracle/cvorb_path.py (restated cv::ORB::create(1000, 1.2, 8, 19)->detectAndCompute with a mask, reference
src/Frame.cc:2623-2627) against the real OpenCV 4.13 implementation behind cv2.ORB_create: bit-exact keypoints and
descriptors, compared as sets ordered by (octave, y, x)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _cv2_sorted(image, mask):
    kps, desc = cv2.ORB_create(1000, 1.2, 8, 19).detectAndCompute(image, mask)
    rec = [(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave) for k in kps]
    order = sorted(range(len(rec)), key=lambda i: (rec[i][5], rec[i][1], rec[i][0]))
    return [rec[i] for i in order], (desc[order] if len(rec) else np.zeros((0, 32), np.uint8))


@pytest.mark.parametrize("shape,seed,masked", [((375, 1242), 3, False), ((375, 1242), 3, True), ((480, 752), 9, True),
                                               ((222, 333), 5, False)])
def test_cvorb_path_equals_cv2_orb(shape, seed, masked):
    from oracle import cvorb_path
    from pointslot_b200.synth import synth_pair
    H, W = shape
    L, _ = synth_pair(W, H, seed)
    mask = None
    if masked:  # object masks as Frame::ExtractObjORB builds them: 255 inside the instance regions, 0 elsewhere
        mask = np.zeros_like(L)
        mask[H // 4:3 * H // 4, W // 6:W // 2] = 255
        mask[H // 8:H // 2, 2 * W // 3:W - 40] = 255
    ref, rdesc = _cv2_sorted(L, mask)
    k, d = cvorb_path.detect_and_compute(L, mask)
    assert len(k) == len(ref) > 100
    for name, col in (("x", 0), ("y", 1), ("size", 2), ("angle", 3), ("response", 4)):
        assert np.array_equal(k[name], np.array([r[col] for r in ref], np.float32)), name
    assert np.array_equal(k["octave"], np.array([r[5] for r in ref], np.int32))
    assert np.array_equal(d, rdesc)


def test_resize_linear_exact_restated():
    from oracle import cvorb_path
    from pointslot_b200.synth import synth_pair
    L, _ = synth_pair(1242, 375, 3)
    noise = np.random.default_rng(1).integers(0, 256, (240, 333), dtype=np.uint8)
    for img, sizes in ((L, ((1035, 313), (863, 260), (620, 187))), (noise, ((278, 200), (231, 167), (100, 70)))):
        for (dw, dh) in sizes:
            assert np.array_equal(cvorb_path.resize_linear_exact(img, dw, dh),
                                  cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)), (dw, dh)


def test_submatrix_gaussian_blur_restated():
    """The float separable blur cv::ORB applies to its pyramid levels; fused multiply-adds on CPUs with FMA dispatch."""
    from oracle import cvorb_path
    kk = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    rng = np.random.default_rng(2)
    ok = {True: True, False: True}
    for _ in range(3):
        img = rng.integers(0, 256, (300, 420), dtype=np.uint8)
        ref = cv2.sepFilter2D(img, cv2.CV_8U, kk, kk, borderType=cv2.BORDER_REFLECT_101)
        for fma in (True, False):
            ok[fma] &= bool(np.array_equal(cvorb_path.blur_float_separable(img, fma), ref))
        # either way it is NOT the fixed-point blur the ORB-SLAM extractor gets on its cloned level
        assert not np.array_equal(ref, cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101))
    assert ok[True] or ok[False], "neither the fused nor the unfused evaluation order reproduces cv2.sepFilter2D"


@pytest.mark.parametrize("shape,seed,masked", [((375, 1242), 3, False), ((375, 1242), 4, True), ((480, 752), 9, True),
                                               ((222, 333), 5, False), ((120, 200), 6, False)])
def test_c_oracle_equals_cv2_orb(shape, seed, masked, oracle_mod):
    """The C restatement (orc_cvorb_detect_and_compute) against cv2.ORB_create itself."""
    from pointslot_b200.synth import synth_pair
    H, W = shape
    L, _ = synth_pair(W, H, seed)
    mask = None
    if masked:
        mask = np.zeros_like(L)
        mask[H // 4:3 * H // 4, W // 6:W // 2] = 255
        mask[H // 8:H // 2, 2 * W // 3:W - 40] = 255
    ref, rdesc = _cv2_sorted(L, mask)
    k, d = oracle_mod.cvorb_detect_and_compute(L, mask)
    assert len(k) == len(ref)
    for name, col in (("x", 0), ("y", 1), ("size", 2), ("angle", 3), ("response", 4)):
        assert np.array_equal(k[name], np.array([r[col] for r in ref], np.float32)), name
    assert np.array_equal(k["octave"], np.array([r[5] for r in ref], np.int32))
    assert np.array_equal(d, rdesc)


def test_c_oracle_equals_cv2_orb_on_varied_content(oracle_mod):
    """Noise, flat, saturated and checkerboard content at odd sizes: ties in the scores and sparse levels."""
    rng = np.random.default_rng(7)
    cases = []
    for (h, w) in ((97, 131), (150, 150), (241, 322), (64, 500)):
        cases.append(rng.integers(0, 256, (h, w), dtype=np.uint8))                      # noise: every quota overflows
        yy, xx = np.mgrid[0:h, 0:w]
        cases.append((((xx // 9 + yy // 7) & 1) * 200 + 20).astype(np.uint8))           # checkerboard: massive score ties
        cases.append(np.clip(rng.normal(128, 60, (h, w)), 0, 255).astype(np.uint8) // 64 * 64)  # posterised
    cases.append(np.full((120, 160), 77, np.uint8))                                     # flat: no keypoints at all
    for img in cases:
        ref, rdesc = _cv2_sorted(img, None)
        k, d = oracle_mod.cvorb_detect_and_compute(img)
        assert len(k) == len(ref), img.shape
        if len(ref):
            got = np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"]], 1)
            want = np.array([r[:5] for r in ref], np.float32)
            assert np.array_equal(got, want), img.shape
            assert np.array_equal(k["octave"], np.array([r[5] for r in ref], np.int32))
            assert np.array_equal(d, rdesc), img.shape


def _equal_to_cv2(k, d, ref, rdesc, what):
    assert len(k) == len(ref), f"{what}: {len(k)} vs {len(ref)} keypoints"
    if len(ref):
        got = np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"]], 1)
        assert np.array_equal(got, np.array([r[:5] for r in ref], np.float32)), what
        assert np.array_equal(k["octave"], np.array([r[5] for r in ref], np.int32)), what
        assert np.array_equal(d, rdesc), what


@pytest.mark.parametrize("shape", [(465, 750), (585, 1123), (832, 609), (312, 561), (390, 513)])
def test_level_sizes_round_like_opencv(shape, oracle_mod):
    """A dimension with dim / 1.2 = x.5: OpenCV 4.13 is observed to multiply by one float reciprocal, cvRound(dim * (1.f / scale)),
    not to divide (465 / 1.2f = 387.49997 -> 387, 465 * (1 / 1.2f) = 387.5 -> 388), and every level from there on is a
    different image.  Found by tools/cvorb_fuzz.py; both restatements follow the library."""
    from oracle import cvorb_path
    from pointslot_b200.synth import synth_pair
    H, W = shape
    img = synth_pair(W, H, 17)[0]
    ref, rdesc = _cv2_sorted(img, None)
    _equal_to_cv2(*oracle_mod.cvorb_detect_and_compute(img), ref, rdesc, f"C oracle {shape}")
    if shape == (465, 750):
        _equal_to_cv2(*cvorb_path.detect_and_compute(img, None), ref, rdesc, f"python path {shape}")


def test_mask_is_zero_or_nonzero(oracle_mod):
    """cv::ORB looks at its mask as zero / non-zero: a region filled with 1, 100 or 254 keeps its keypoints on EVERY level
    (not only on level 0, as resizing the raw mask and thresholding at 254 would give).  The reference itself only builds
    0 / 255 masks (Frame.cc:2637-2646)."""
    from oracle import cvorb_path
    from pointslot_b200.synth import synth_pair
    img = synth_pair(871, 409, 7)[0]
    m = np.zeros_like(img)
    m[50:300, 100:500] = 255
    m[100:380, 450:800] = 100
    m[20:90, 600:860] = 1
    ref, rdesc = _cv2_sorted(img, m)
    ref255, _ = _cv2_sorted(img, np.where(m != 0, 255, 0).astype(np.uint8))
    assert ref == ref255 and len(ref) > 500
    _equal_to_cv2(*oracle_mod.cvorb_detect_and_compute(img, m), ref, rdesc, "C oracle")
    _equal_to_cv2(*cvorb_path.detect_and_compute(img, m), ref, rdesc, "python path")


def test_c_oracle_equals_cv2_orb_randomised(oracle_mod):
    """A bounded run of tools/cvorb_fuzz.py (random shapes, parameters, content classes, object masks with arbitrary
    non-zero values); the long sweeps are in profiles/r02_cvorb_fuzz.txt."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("cvorb_fuzz", Path(__file__).resolve().parents[1] / "tools" / "cvorb_fuzz.py")
    fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
    for flags in ({"ref"}, {"anymask"}):
        s = fz.run(25, 1234, flags)
        assert s["cases"] == 25 and s["mismatches"] == 0 and s["keypoints"] > 3000, s


def test_c_oracle_refuses_what_the_abi_refuses(oracle_mod):
    """edgeThreshold < 19 lets the steered pattern (radius 18.4) leave the level; neither the device nor the restatement
    models the unblurred frame cv::ORB samples there, so both refuse (pslot_cvorb_create: PSLOT_ERR_ARG)."""
    img = np.zeros((200, 300), np.uint8)
    for bad in (dict(edge_threshold=18), dict(scale_factor=1.0), dict(fast_threshold=0), dict(nfeatures=0)):
        with pytest.raises(ValueError):
            oracle_mod.cvorb_detect_and_compute(img, None, **bad)
