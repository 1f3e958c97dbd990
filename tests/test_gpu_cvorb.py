"""GPU: pslot_cvorb_detect_and_compute (cvorb.cu) - OpenCV's own ORB for the dense object features, reference
src/Frame.cc:2623-2627 `cv::ORB::create(1000, 1.2, 8, 19)->detectAndCompute(im, ObjMask, kp, descriptor)` - compared
DIRECTLY with cv2.ORB_create (OpenCV 4.13 itself) and with the C oracle: the keypoint set and every field / descriptor
bit, ordered by (octave, y, x) (cv::ORB's own order is what std::nth_element leaves)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def B():
    from pointslot_b200 import binding
    return binding


def _cv2_sorted(image, mask, *args):
    kps, desc = cv2.ORB_create(*(args or (1000, 1.2, 8, 19))).detectAndCompute(image, mask)
    rec = [(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave) for k in kps]
    order = sorted(range(len(rec)), key=lambda i: (rec[i][5], rec[i][1], rec[i][0]))
    return [rec[i] for i in order], (desc[order] if len(rec) else np.zeros((0, 32), np.uint8))


def _object_mask(H, W):
    """Object masks as Frame::ExtractObjORB builds them (Frame.cc:2634-2645): 255 inside the instance regions."""
    mask = np.zeros((H, W), np.uint8)
    mask[H // 4:3 * H // 4, W // 6:W // 2] = 255
    mask[H // 8:H // 2, 2 * W // 3:W - 40] = 255
    return mask


def _assert_equal(k, d, ref, rdesc, what):
    assert len(k) == len(ref), f"{what}: {len(k)} vs {len(ref)} keypoints"
    if not len(ref):
        return
    got = np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"]], 1)
    want = np.array([r[:5] for r in ref], np.float32)
    for col, name in enumerate(("x", "y", "size", "angle", "response")):
        assert np.array_equal(got[:, col].view(np.uint32), want[:, col].view(np.uint32)), f"{what}: {name}"
    assert np.array_equal(k["octave"], np.array([r[5] for r in ref], np.int32)), f"{what}: octave"
    assert np.all(k["class_id"] == -1)
    assert np.array_equal(d, rdesc), f"{what}: descriptors"


@pytest.mark.parametrize("shape,seed,masked", [((375, 1242), 3, False), ((375, 1242), 4, True), ((480, 752), 9, True),
                                               ((480, 752), 10, False), ((222, 333), 5, False), ((120, 200), 6, False)])
def test_cvorb_equals_cv2_orb(shape, seed, masked, B, oracle_mod):
    from pointslot_b200.synth import synth_pair
    H, W = shape
    L, R = synth_pair(W, H, seed)
    mask = _object_mask(H, W) if masked else None
    orb = B.CvOrb(1000, 1.2, 8, 19)
    for name, img in (("left", L), ("right", R)):  # the same context serves image after image
        ref, rdesc = _cv2_sorted(img, mask)
        k, d = orb.detect_and_compute(img, mask)
        _assert_equal(k, d, ref, rdesc, f"{shape} {name} vs cv2")
        ok, od = oracle_mod.cvorb_detect_and_compute(img, mask)
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    assert len(k) > (100 if H > 150 else 20)


def test_cvorb_natural_texture_with_mask(B):
    """The reference's only image asset (via the golden fixture) with an object mask."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "kitti_natural_2000.npz")
    img = np.ascontiguousarray(g["left"])
    mask = _object_mask(*img.shape)
    orb = B.CvOrb()
    for m in (mask, None):
        ref, rdesc = _cv2_sorted(img, m)
        k, d = orb.detect_and_compute(img, m)
        _assert_equal(k, d, ref, rdesc, "natural texture")
    assert len(k) >= 1000


def test_cvorb_varied_content_and_parameters(B):
    """Noise (every quota overflows), checkerboards (massive score ties, which retainBest keeps), posterised and flat
    content at odd sizes; other feature counts / level counts / thresholds; pitched input."""
    rng = np.random.default_rng(7)
    cases = []
    for (h, w) in ((97, 131), (150, 150), (241, 322), (64, 500)):
        cases.append(rng.integers(0, 256, (h, w), dtype=np.uint8))
        yy, xx = np.mgrid[0:h, 0:w]
        cases.append((((xx // 9 + yy // 7) & 1) * 200 + 20).astype(np.uint8))
        cases.append((np.clip(rng.normal(128, 60, (h, w)), 0, 255).astype(np.uint8) // 64 * 64).astype(np.uint8))
    cases.append(np.full((120, 160), 77, np.uint8))
    orb = B.CvOrb()
    for img in cases:
        ref, rdesc = _cv2_sorted(img, None)
        k, d = orb.detect_and_compute(img)
        _assert_equal(k, d, ref, rdesc, f"varied {img.shape}")
    from pointslot_b200.synth import synth_pair
    L, _ = synth_pair(640, 400, 12)
    for args in ((500, 1.2, 8, 19), (2000, 1.2, 4, 19), (800, 1.5, 5, 31), (1000, 1.2, 8, 19, 0, 2, 0, 31, 10)):
        ref, rdesc = _cv2_sorted(L, None, *args)
        fast_th = args[8] if len(args) > 8 else 20
        k, d = B.CvOrb(args[0], args[1], args[2], args[3], fast_th).detect_and_compute(L)
        _assert_equal(k, d, ref, rdesc, f"params {args}")
    big = np.zeros((400, 768), np.uint8); big[:, :640] = L   # a view with a row pitch
    ref, rdesc = _cv2_sorted(L, None)
    k, d = orb.detect_and_compute(big[:, :640])
    _assert_equal(k, d, ref, rdesc, "pitched input")


def test_cvorb_batch_equals_cv2_orb(B):
    """pslot_cvorb_detect_and_compute_batch (the two views of Frame::ExtractObjORB, Frame.cc:2648-2651, or more frames in one
    call): every image of the batch equals cv2.ORB_create on that image - with masks, with a mask on some images only,
    without masks - and the same context then serves other batch sizes and the single-image call."""
    from pointslot_b200.synth import synth_pair
    H, W = 375, 1242
    imgs = []
    for seed in (31, 32, 33):
        imgs.extend(synth_pair(W, H, seed))
    m1 = _object_mask(H, W)
    m2 = np.zeros((H, W), np.uint8); m2[60:300, 500:1100] = 255
    orb = B.CvOrb(1000, 1.2, 8, 19)
    want = {}

    def ref(i, m):
        key = (i, None if m is None else id(m))
        if key not in want:
            want[key] = _cv2_sorted(imgs[i], m)
        return want[key]

    # left + right with their object masks: the reference's call pattern
    out = orb.detect_and_compute_batch(imgs[:2], [m1, m2])
    for i, m in ((0, m1), (1, m2)):
        _assert_equal(*out[i], *ref(i, m), f"pair, image {i}")
    # six images, masks on some
    masks = [m1, None, m2, None, None, m1]
    out = orb.detect_and_compute_batch(imgs, masks)
    assert len(out) == 6
    for i, m in enumerate(masks):
        _assert_equal(*out[i], *ref(i, m), f"batch of 6, image {i}")
    # no masks at all, a smaller batch on the same context, then the single-image entry point
    out = orb.detect_and_compute_batch(imgs[:3])
    for i in range(3):
        _assert_equal(*out[i], *ref(i, None), f"batch of 3, image {i}")
        k, d = orb.detect_and_compute(imgs[i])
        assert k.tobytes() == out[i][0].tobytes() and np.array_equal(d, out[i][1])
    assert orb.detect_and_compute_batch([]) == []
    # another size re-configures the context; noise overflows every quota and keeps response ties
    rng = np.random.default_rng(5)
    noise = [rng.integers(0, 256, (150, 210), dtype=np.uint8) for _ in range(4)]
    out = orb.detect_and_compute_batch(noise)
    for i in range(4):
        _assert_equal(*out[i], *_cv2_sorted(noise[i], None), f"noise batch, image {i}")


def test_cvorb_large_batches_equal_single_calls(B):
    """Batches of 16 images and more travel in chunks (uploads run ahead on a second stream; four chunks from pageable
    memory, two from page-locked memory): every image's result must equal the single-image call's, which the tests above
    tie to cv2.ORB_create.  KITTI-sized frames, masks on some images only."""
    from pointslot_b200.synth import synth_pair
    H, W = 375, 1242
    base = []
    for seed in (51, 52, 53, 54):
        base.extend(synth_pair(W, H, seed))
    mask = _object_mask(H, W)
    orb = B.CvOrb(1000, 1.2, 8, 19)
    single = {}
    for i, img in enumerate(base):
        for m in (None, mask):
            single[(i, m is not None)] = orb.detect_and_compute(img, m)
    for n in (16, 37, 64):
        imgs = [base[i % 8] for i in range(n)]
        masks = [mask if (i * 7) % 3 == 0 else None for i in range(n)]
        for ms in (None, masks):
            out = orb.detect_and_compute_batch(imgs, ms)
            assert len(out) == n
            for i, (k, d) in enumerate(out):
                sk, sd = single[(i % 8, ms is not None and ms[i] is not None)]
                assert k.tobytes() == sk.tobytes() and np.array_equal(d, sd), f"batch of {n}, image {i}"
    # views with a row pitch (cv::Mat ROIs): the 2-D upload path, un-chunked and chunked
    wide = np.zeros((20, H, 1280), np.uint8); wmask = np.zeros((H, 1300), np.uint8)
    for i in range(20):
        wide[i, :, :W] = base[i % 8]
    wmask[:, :W] = mask
    for n in (3, 20):
        out = orb.detect_and_compute_batch([wide[i, :, :W] for i in range(n)], [wmask[:, :W] if i % 2 else None for i in range(n)])
        for i, (k, d) in enumerate(out):
            sk, sd = single[(i % 8, i % 2 == 1)]
            assert k.tobytes() == sk.tobytes() and np.array_equal(d, sd), f"pitched batch of {n}, image {i}"
    # page-locked input: two chunks from 32 images on
    pin = B.PinnedBuffer((40, H, W))
    for i in range(40):
        pin.array[i] = base[i % 8]
    out = orb.detect_and_compute_batch([pin.array[i] for i in range(40)])
    for i, (k, d) in enumerate(out):
        sk, sd = single[(i % 8, False)]
        assert k.tobytes() == sk.tobytes() and np.array_equal(d, sd), f"pinned batch, image {i}"
    del out
    pin.close()


def test_cvorb_on_the_images_of_a_batch(B):
    """pslot_cvorb_detect_and_compute_resident: cv::ORB on the images a StereoBatch still holds on the device (level 0 of its
    pyramids) - left views, right views, with masks, a sub-range, 40 images (two chunks) - equals the calls fed from the host."""
    from pointslot_b200.synth import synth_pair
    W, H, P = 1242, 375, 20
    pairs = [synth_pair(W, H, 70 + s) for s in range(4)]
    Ls = np.stack([pairs[i % 4][0] for i in range(P)]); Rs = np.stack([pairs[i % 4][1] for i in range(P)])
    sb = B.StereoBatch((1000, 1.2, 8, 20, 7), W, H, P)
    sb.run(Ls, Rs, 386.1448, B.max_disparity(386.1448, 718.856))
    mask = _object_mask(H, W)
    orb = B.CvOrb(1000, 1.2, 8, 19)
    want = {}
    for i in range(4):
        for view, img in enumerate(pairs[i]):
            for m in (None, mask):
                want[(i, view, m is not None)] = orb.detect_and_compute(img, m)
    out = orb.detect_and_compute_resident(sb, 0, 2 * P)                       # all 40 images, no masks
    for s, (k, d) in enumerate(out):
        wk, wd = want[(s % P % 4, s // P, False)]
        assert k.tobytes() == wk.tobytes() and np.array_equal(d, wd), f"image {s}"
    # ComputeObjStereoMatches for all 20 pairs at once on those resident features == the per-pair call with the same arrays
    bf, maxD = 386.1448, B.max_disparity(386.1448, 718.856)
    ur, dp = orb.stereo_match_resident(sb, P, bf, maxD)
    eL, eR = B.OrbExtractor(1000, 1.2, 8, 20, 7), B.OrbExtractor(1000, 1.2, 8, 20, 7)
    for p_ in (0, 1, 2, 3, 19):
        eL.extract(Ls[p_]); eR.extract(Rs[p_])
        (kl, dl), (kr, dr) = out[p_], out[P + p_]
        wu, wd = B.stereo_match(eL, eR, kl, dl, kr, dr, bf, maxD)
        n_ = len(kl)
        assert np.array_equal(ur[p_, :n_].view(np.uint32), wu.view(np.uint32)) and np.array_equal(dp[p_, :n_].view(np.uint32), wd.view(np.uint32)), p_
        assert (wu >= 0).sum() > 30
    masks = [mask if s % 3 else None for s in range(2 * P)]
    out = orb.detect_and_compute_resident(sb, 0, 2 * P, masks)                # masks on two thirds of them
    for s, (k, d) in enumerate(out):
        wk, wd = want[(s % P % 4, s // P, masks[s] is not None)]
        assert k.tobytes() == wk.tobytes() and np.array_equal(d, wd), f"masked image {s}"
    out = orb.detect_and_compute_resident(sb, P + 3, 2, [mask, mask])         # the right views of pairs 3 and 4
    for j, (k, d) in enumerate(out):
        wk, wd = want[((3 + j) % 4, 1, True)]
        assert k.tobytes() == wk.tobytes() and np.array_equal(d, wd)
    with pytest.raises(B.PslotError):
        orb.stereo_match_resident(sb, P, bf, maxD)   # the context now holds a sub-range, not the whole batch
    with pytest.raises(B.PslotError):
        orb.detect_and_compute_resident(sb, 2 * P - 1, 2)


def test_cvorb_full_hd_pair(B):
    """BASELINE.json's largest shape (1920x1080, 4000 features) as a left + right batch with object masks."""
    from pointslot_b200.synth import synth_pair
    H, W = 1080, 1920
    L, R = synth_pair(W, H, 41)
    mask = _object_mask(H, W)
    orb = B.CvOrb(4000, 1.2, 8, 19)
    out = orb.detect_and_compute_batch([L, R], [mask, mask])
    for name, img, (k, d) in (("left", L, out[0]), ("right", R, out[1])):
        ref, rdesc = _cv2_sorted(img, mask, 4000, 1.2, 8, 19)
        _assert_equal(k, d, ref, rdesc, f"1080p {name}")
        assert len(k) >= 3500


def test_cvorb_argument_errors(B):
    with pytest.raises(B.PslotError):
        B.CvOrb(1000, 1.0, 8, 19)
    with pytest.raises(B.PslotError):
        B.CvOrb(1000, 1.2, 8, 5)
    orb = B.CvOrb()
    k, d = orb.detect_and_compute(np.zeros((30, 30), np.uint8))  # smaller than 2 * edgeThreshold + 1: cv::ORB returns nothing
    assert len(k) == 0


def test_cvorb_stages_against_the_oracle(B, oracle_mod):
    """Stage accessors: every pyramid level equals the INTER_LINEAR_EXACT chain (cv2), and the per-level candidate set
    (FAST score + whole-level maxima test + border filter) equals the oracle's cv::FAST on that level."""
    from pointslot_b200.synth import synth_pair
    L, _ = synth_pair(752, 480, 21)
    orb = B.CvOrb(1200, 1.2, 8, 19)
    orb.detect_and_compute(L)
    prev = L
    for l in range(8):
        img = orb.stage(0, l)
        if l:
            inv = np.float32(1) / np.float32(np.float64(np.float32(1.2)) ** l)  # one float reciprocal, as OpenCV 4.13 is observed to do
            w, h = int(np.rint(np.float32(752) * inv)), int(np.rint(np.float32(480) * inv))
            assert img.shape == (h, w)
            assert np.array_equal(img, cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR_EXACT)), f"level {l}"
        prev = img
        fx, fy, fr = oracle_mod.fast9_16(img, 20)
        h, w = img.shape
        keep = (fx >= 19) & (fx < w - 19) & (fy >= 19) & (fy < h - 19)
        want = set(zip(fx[keep].tolist(), fy[keep].tolist(), fr[keep].tolist()))
        cx, cy, cs = orb.stage(2, l)
        assert set(zip(cx.tolist(), cy.tolist(), cs.tolist())) == want and len(cx) == len(want), f"candidates of level {l}"
        kx, ky, ks, kr = orb.stage(3, l)
        # retainBest(2 * quota) by the FAST score keeps a score-threshold set, ties included
        kept = set(zip(kx.tolist(), ky.tolist(), ks.tolist()))
        assert kept <= want and all(s < ks.min() for (_, _, s) in want - kept)
        assert np.all(np.isfinite(kr))


@pytest.mark.parametrize("shape", [(465, 750), (585, 1123), (312, 561)])
def test_cvorb_level_sizes_round_like_opencv(shape, B, oracle_mod):
    """A dimension with dim / 1.2 = x.5 (465 * (1 / 1.2f) rounds to 388, 465 / 1.2f to 387): OpenCV 4.13 is observed to multiply by one
    float reciprocal, and every level from there on is a different image.  Found by tools/cvorb_fuzz.py."""
    from pointslot_b200.synth import synth_pair
    H, W = shape
    img = synth_pair(W, H, 17)[0]
    mask = _object_mask(H, W)
    orb = B.CvOrb(1000, 1.2, 8, 19)
    for m in (None, mask):
        ref, rdesc = _cv2_sorted(img, m)
        k, d = orb.detect_and_compute(img, m)
        _assert_equal(k, d, ref, rdesc, f"{shape} mask={m is not None} vs cv2")
        ok, od = oracle_mod.cvorb_detect_and_compute(img, m)
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)


def test_cvorb_mask_is_zero_or_nonzero(B):
    """cv::ORB looks at its mask as zero / non-zero: regions filled with 1 or 100 keep their keypoints on every level."""
    from pointslot_b200.synth import synth_pair
    img = synth_pair(871, 409, 7)[0]
    m = np.zeros_like(img)
    m[50:300, 100:500] = 255
    m[100:380, 450:800] = 100
    m[20:90, 600:860] = 1
    ref, rdesc = _cv2_sorted(img, m)
    orb = B.CvOrb(1000, 1.2, 8, 19)
    k, d = orb.detect_and_compute(img, m)
    _assert_equal(k, d, ref, rdesc, "mask values 255 / 100 / 1 vs cv2")
    (k2, d2), (k3, d3) = orb.detect_and_compute_batch([img, img], [m, None])   # the batched path reads the same mask plane
    assert k2.tobytes() == k.tobytes() and np.array_equal(d2, d)
    ref3, rdesc3 = _cv2_sorted(img, None)
    _assert_equal(k3, d3, ref3, rdesc3, "unmasked image of the same batch vs cv2")


def test_cvorb_randomised_against_cv2(B):
    """Cases drawn by tools/cvorb_fuzz.py (random shapes, parameters, content classes, object masks with arbitrary non-zero
    values): the device against cv2.ORB_create itself."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("cvorb_fuzz", Path(__file__).resolve().parents[1] / "tools" / "cvorb_fuzz.py")
    fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
    rng = np.random.default_rng(99)
    done = n = 0
    while done < 16:
        n += 1
        W, H, p = fz.draw_case(rng)
        if n % 2:
            p = (1000, 1.2, 8, 19, 20)
        kind = fz.CONTENT[n % len(fz.CONTENT)]
        if kind == "checker" or p[0] > 2000:
            continue  # thousands of exactly tied responses on one level exceed the on-chip sort (PSLOT_ERR_CAPACITY, by design)
        img = fz.make_image(kind, W, H, rng, 990000 + n)
        mask = fz.make_mask(W, H, rng, True) if n % 3 else None
        try:
            orb = B.CvOrb(*p)
        except B.PslotError:
            continue  # edgeThreshold < 19
        rec, octv, rdesc = fz.cv2_sorted(img, mask, p)
        k, d = orb.detect_and_compute(img, mask)
        what = f"{W}x{H} {p} mask={mask is not None}"
        assert len(k) == len(rec), f"{what}: {len(k)} vs {len(rec)} keypoints"
        if len(rec):
            got = np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"]], 1)
            assert np.array_equal(got.view(np.uint32), rec.view(np.uint32)), what
            assert np.array_equal(k["octave"], octv) and np.array_equal(d, rdesc), what
        orb.close()
        done += 1
