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
