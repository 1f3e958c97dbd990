"""GPU: pslot_assign_features / pslot_batch_assign_features (partition.cu) - Frame::AssignFeatures, reference
src/Frame.cc:759-977: classification of the features by the instance mask and the push_back order of every set - against the
oracle, which tests/test_oracle_reference.py pins to the reference's own function."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def B():
    from pointslot_b200 import binding
    return binding


def _mask_and_ids(H, W):
    mask = np.zeros((H, W), np.uint8)
    labels = [1, 4, 9, 30, 200, 254, 77]
    for i, lab in enumerate(labels):
        x0 = 40 + i * (W - 120) // 7
        mask[H // 6 + 5 * i:H - H // 5 - 4 * i, x0:x0 + (W - 120) // 8] = lab
    mask[:, W - 30:] = 255
    mask[H - 50:, W // 3:2 * W // 3] = 255
    ids = np.array([0, 3 + 255, -1, 8, 8, 29, 199, 253, 500, 120], np.int32)
    return mask, ids


def _same(got, want):
    gc, gg, ge = got
    wc, wg, we = want
    assert np.array_equal(gc, wc), "classes"
    assert len(gg) == len(wg)
    for c, (a, b) in enumerate(zip(gg, wg)):
        assert np.array_equal(a, b), f"order of class {c}"
    if we is not None:
        assert np.array_equal(ge, we), "effective depth counts"


def test_assign_features_equals_oracle(B, oracle_mod):
    """The frame's own ORB features and the dense cv::ORB object features of an extracted pair, one shared mask."""
    from pointslot_b200.synth import synth_pair
    W, H = 1242, 375
    L, R = synth_pair(W, H, 5)
    mask, ids = _mask_and_ids(H, W)
    eL, eR = B.OrbExtractor(2000, 1.2, 8, 20, 7), B.OrbExtractor(2000, 1.2, 8, 20, 7)
    kL, dL = eL.extract(L); kR, dR = eR.extract(R)
    bf = 386.1448
    ur, dp = B.stereo_match(eL, eR, kL, dL, kR, dR, bf, B.max_disparity(bf, 718.856))
    got = B.assign_features(eL, kL, mask, ids)
    _same(got, oracle_mod.assign_features(kL, mask, ids))
    assert len(got[1][0]) > 200 and len(got[1][1]) > 20 and (got[0] == -1).sum() > 20
    # the dense object features: cv::ORB inside the object regions, stereo depth, the mask of the previous call reused
    obj_mask = ((mask != 0) & (mask != 255)).astype(np.uint8) * 255  # Frame.cc:2634-2645
    orb = B.CvOrb(1000, 1.2, 8, 19)
    (okL, odL), (okR, odR) = orb.detect_and_compute_batch([L, R], [obj_mask, obj_mask])
    our, odp = B.stereo_match(eL, eR, okL, odL, okR, odR, bf, B.max_disparity(bf, 718.856))
    th_depth = F(35.0)
    got = B.assign_features(eL, okL, None, ids, depth=odp, depth_limit=float(F(2) * th_depth))
    want = oracle_mod.assign_features(okL, mask, ids, depth=odp, depth_limit=float(F(2) * th_depth))
    _same(got, want)
    # (a coarse-level keypoint scaled back to level 0 may land just outside its object region: background, as in the reference)
    assert len(got[1][0]) < 20 and sum(len(g) for g in got[1][1:]) > 500 and got[2].sum() > 0


def test_assign_features_edge_cases(B, oracle_mod):
    rng = np.random.default_rng(3)
    H, W = 120, 200
    mask = rng.integers(0, 256, (H, W), dtype=np.uint8)  # every label, pixel by pixel
    ext = B.OrbExtractor(500, 1.2, 8, 20, 7)
    for n, J in ((0, 3), (1, 0), (33, 1), (4097, 256), (700, 40)):
        k = np.zeros(n, B.KEYPOINT_DTYPE)
        k["x"] = rng.uniform(0, W - 0.01, n).astype(F); k["y"] = rng.uniform(0, H - 0.01, n).astype(F)
        ids = rng.integers(-3, 520, J).astype(np.int32)
        dp = rng.uniform(-2, 90, n).astype(F)
        _same(B.assign_features(ext, k, mask, ids, depth=dp, depth_limit=40.0),
              oracle_mod.assign_features(k, mask, ids, depth=dp, depth_limit=40.0))
    # a pitched mask view, and positions outside the image (the reference would read out of bounds: dropped here)
    big = np.zeros((H, 256), np.uint8); big[:, :W] = mask
    k = np.zeros(5, B.KEYPOINT_DTYPE); k["x"] = [3, 199.9, 250, -4, 17]; k["y"] = [5, 119.5, 10, 10, 300]
    ids = np.arange(0, 254, dtype=np.int32)
    got = B.assign_features(ext, k, big[:, :W], ids)
    _same(got, oracle_mod.assign_features(k, mask, ids))
    assert got[0][2] == -1 and got[0][3] == -1 and got[0][4] == -1
    with pytest.raises(B.PslotError):
        B.assign_features(ext, k, mask, np.zeros(257, np.int32))
    with pytest.raises(B.PslotError):
        B.assign_features(B.OrbExtractor(500, 1.2, 8, 20, 7), k, None, ids)  # no mask uploaded yet on this context


def test_batch_assign_features(B, oracle_mod):
    """The left keypoints of every pair of a batch, partitioned where they are (on the device)."""
    from pointslot_b200.synth import synth_pair
    W, H, P = 640, 360, 4
    pairs = [synth_pair(W, H, 60 + s) for s in range(P)]
    Ls = np.stack([p[0] for p in pairs]); Rs = np.stack([p[1] for p in pairs])
    sb = B.StereoBatch((800, 1.2, 8, 20, 7), W, H, P)
    res = sb.run(Ls, Rs, 200.0, B.max_disparity(200.0, 500.0))
    masks, idsl = [], []
    for s in range(P):
        m, ids = _mask_and_ids(H, W)
        masks.append(np.roll(m, 37 * s, axis=1)); idsl.append(np.roll(ids, s))
    cls, order, cs = sb.assign_features(np.stack(masks), np.stack(idsl))
    sb.wait()
    for s in range(P):
        k = res[s]["kL"]
        wc, wg, _ = oracle_mod.assign_features(k, masks[s], idsl[s])
        assert np.array_equal(cls[s, :len(k)], wc)
        for c, g in enumerate(wg):
            assert np.array_equal(order[s, cs[s, c]:cs[s, c + 1]], g), f"pair {s}, class {c}"
        assert cs[s, -1] == sum(len(g) for g in wg) and len(wg[0]) > 50
