"""GPU: the C++ drop-in classes (ORB_SLAM2::ORBextractor, Frame stereo constructor, ORBmatcher::
DescriptorDistance) driven exactly as Tracking::GrabImageStereo drives the reference's, vs the oracle."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
O_MP_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"),
                       ("level", "<i4"), ("flags", "<i4")])
ROOT = Path(__file__).resolve().parents[1]
SHIM = ROOT / "pointslot_b200" / "cpp"


def test_cpp_shim_matches_oracle(tmp_path, oracle_mod):
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_pair

    subprocess.check_call(["make", "-C", str(SHIM), "-s"])
    W, H = 640, 400
    params = (800, 1.2, 8, 20, 7)
    bf, fx = 300.0, 600.0
    L, R = synth_pair(W, H, 21)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<6i3f", W, H, params[0], params[2], params[3], params[4], params[1], bf, fx))
        f.write(L.tobytes()); f.write(R.tobytes())
    from pointslot_b200.vocab import random_vocabulary
    voc = random_vocabulary(10, 4, seed=2)
    voc.to_text(tmp_path / "voc.txt")
    # instance mask + detections for Frame::AssignFeatures (Frame.cc:759-977): ids above 255, a null detection, a repeated
    # id, ids without a region, a region without a detection, ignored (255) areas
    seg = np.zeros((H, W), np.uint8)
    for i, lab in enumerate([1, 4, 9, 30, 200, 254, 77]):
        x0 = 20 + i * 85
        seg[50 + 6 * i:340 - 5 * i, x0:x0 + 70] = lab
    seg[:, W - 25:] = 255
    seg[360:, 200:420] = 255
    det_ids = np.array([0, 3 + 255, -1, 8, 8, 29, 199, 253, 500, 120], np.int32)
    with open(tmp_path / "seg.bin", "wb") as f:
        f.write(struct.pack("<i", len(det_ids))); f.write(det_ids.tobytes()); f.write(seg.tobytes())
    subprocess.check_call([str(SHIM / "test_shim"), str(inp), str(out), str(tmp_path / "voc.txt"), str(tmp_path / "seg.bin")])
    raw = out.read_bytes()
    N, Nr = struct.unpack_from("<2i", raw, 0)
    o = 8
    kL = np.frombuffer(raw, B.KEYPOINT_DTYPE, N, o); o += 28 * N
    dL = np.frombuffer(raw, np.uint8, 32 * N, o).reshape(N, 32); o += 32 * N
    kR = np.frombuffer(raw, B.KEYPOINT_DTYPE, Nr, o); o += 28 * Nr
    dR = np.frombuffer(raw, np.uint8, 32 * Nr, o).reshape(Nr, 32); o += 32 * Nr
    ur = np.frombuffer(raw, np.float32, N, o); o += 4 * N
    dp = np.frombuffer(raw, np.float32, N, o); o += 4 * N
    pw, ph = struct.unpack_from("<2i", raw, o); o += 8
    p1 = np.frombuffer(raw, np.uint8, pw * ph, o).reshape(ph, pw); o += pw * ph
    (dd,) = struct.unpack_from("<i", raw, o); o += 4

    def read_list():
        nonlocal o
        (n,) = struct.unpack_from("<i", raw, o); o += 4
        v = np.frombuffer(raw, np.int32, n, o); o += 4 * n
        return v
    grid = [read_list() for _ in range(64 * 48)]
    areas = [read_list() for _ in range(6)]
    (n_mp,) = struct.unpack_from("<i", raw, o); o += 4
    mps = np.frombuffer(raw, O_MP_DTYPE, n_mp, o); o += 24 * n_mp
    md = np.frombuffer(raw, np.uint8, 32 * n_mp, o).reshape(n_mp, 32); o += 32 * n_mp
    (nmatches,) = struct.unpack_from("<i", raw, o); o += 4
    holder = np.frombuffer(raw, np.int32, N, o); o += 4 * N
    n_obj_l, n_obj_r = struct.unpack_from("<2i", raw, o); o += 8
    obj_ur = np.frombuffer(raw, np.float32, n_obj_l, o); o += 4 * n_obj_l
    obj_dp = np.frombuffer(raw, np.float32, n_obj_l, o); o += 4 * n_obj_l
    (n_other,) = struct.unpack_from("<i", raw, o); o += 4
    n_ol, n_or = struct.unpack_from("<2i", raw, o); o += 8
    okl = np.frombuffer(raw, B.KEYPOINT_DTYPE, n_ol, o); o += 28 * n_ol
    odl = np.frombuffer(raw, np.uint8, 32 * n_ol, o).reshape(n_ol, 32); o += 32 * n_ol
    okr = np.frombuffer(raw, B.KEYPOINT_DTYPE, n_or, o); o += 28 * n_or
    odr = np.frombuffer(raw, np.uint8, 32 * n_or, o).reshape(n_or, 32); o += 32 * n_or
    n_words, n_fv = struct.unpack_from("<2i", raw, o); o += 8
    bow = np.frombuffer(raw, np.dtype([("id", "<i4"), ("val", "<f8")]), n_words, o); o += 12 * n_words
    featvec = []
    for _ in range(n_fv):
        node, cnt = struct.unpack_from("<2i", raw, o); o += 8
        featvec.append((node, np.frombuffer(raw, np.uint32, cnt, o))); o += 4 * cnt
    # ---- Frame::AssignFeatures section
    (n_ol2,) = struct.unpack_from("<i", raw, o); o += 4
    f_our = np.frombuffer(raw, np.float32, n_ol2, o); o += 4 * n_ol2
    f_odp = np.frombuffer(raw, np.float32, n_ol2, o); o += 4 * n_ol2
    n_kept, n_bg = struct.unpack_from("<2i", raw, o); o += 8

    def read_set(n):
        nonlocal o
        k = np.frombuffer(raw, B.KEYPOINT_DTYPE, n, o); o += 28 * n
        kun = np.frombuffer(raw, B.KEYPOINT_DTYPE, n, o); o += 28 * n
        d = np.frombuffer(raw, np.uint8, 32 * n, o).reshape(n, 32); o += 32 * n
        u = np.frombuffer(raw, np.float32, n, o); o += 4 * n
        z = np.frombuffer(raw, np.float32, n, o); o += 4 * n
        return k, kun, d, u, z
    bg_set = read_set(n_bg)
    det_sets = []
    for _ in range(n_kept):
        head = struct.unpack_from("<7i", raw, o); o += 28
        det_sets.append((head, read_set(head[1]), read_set(head[2])))
    assert o == len(raw)

    O = oracle_mod
    oL, oR = O.OracleExtractor(*params), O.OracleExtractor(*params)
    okL, odL = oL.extract(L); okR, odR = oR.extract(R)
    maxD = B.max_disparity(bf, fx)  # the shim sets mb = mbf/fx before matching (parity rule 2)
    our, odp = O.stereo_match(oL, oR, okL, odL, okR, odR, bf, maxD)
    assert kL.tobytes() == okL.tobytes() and kR.tobytes() == okR.tobytes()
    assert np.array_equal(dL, odL) and np.array_equal(dR, odR)
    assert np.array_equal(ur >= 0, our >= 0) and np.all(np.abs(ur - our) <= 1e-3)
    assert np.allclose(dp, odp, rtol=1e-5, atol=0)
    assert np.array_equal(p1, oL.level(1))  # mvImagePyramid fills itself on first host access (no explicit sync)
    assert dd == O.hamming256(odL[0], odL[1])
    # mGrid (Frame.cc:1636-1655) and GetFeaturesInArea (:1808-1861), host walk and batched device form
    inv_w, inv_h = float(np.float32(64) / np.float32(W)), float(np.float32(48) / np.float32(H))
    cs, idx = O.assign_grid(okL, 0.0, 0.0, inv_w, inv_h)
    for c in range(64 * 48):
        assert np.array_equal(grid[c], idx[cs[c]:cs[c + 1]])
    qs = [(W * 0.5, H * 0.5, 40.0, -1, -1), (W * 0.25, H * 0.7, 25.0, 1, 4), (5.0, 5.0, 90.0, 0, -1)]
    for i, q in enumerate(qs):
        ref = O.features_in_area(okL, cs, idx, 0.0, 0.0, inv_w, inv_h, 64, 48, *q)
        assert np.array_equal(areas[i], ref) and np.array_equal(areas[3 + i], ref)
    assert sum(len(a) for a in areas) > 0
    # ORBmatcher::SearchByProjection (ORBmatcher.cc:68-155) through the drop-in class: F.mvpMapPoints afterwards
    sf = np.array(B.OrbExtractor(*params).tables()["scale"], np.float32)  # = F.mvScaleFactors (ORBextractor.cc:417-424)
    match, nref = O.search_by_projection(okL, odL, our, cs, idx, 0.0, 0.0, inv_w, inv_h, 64, 48, mps, md,
                                         np.zeros(N, np.uint8), sf, 3.0, 0.8)
    expect = np.full(N, -1, np.int32)
    for i, m in enumerate(match):
        if m >= 0:
            expect[m] = i
    assert nmatches == nref and nref > 50
    assert np.array_equal(holder, expect)

    # The object-mode constructor (Frame.h:86-90) built a second frame on the same extractors in between - the area and
    # projection queries above still saw the FIRST frame's grid (Frame::EnsureDeviceGrid) - and its object stage ran
    # Frame::ComputeObjStereoMatches (Frame.cc:2318-2503) on a subset of the features: same routine, same answer.
    L2, R2 = np.roll(L, 7, axis=0), np.roll(R, 7, axis=0)
    k2L, d2L = oL.extract(L2); k2R, d2R = oR.extract(R2)
    assert n_other == len(k2L) and n_obj_l == len(k2L[::3]) and n_obj_r == len(k2R[::2])
    our2, odp2 = O.stereo_match(oL, oR, k2L[::3], d2L[::3], k2R[::2], d2R[::2], bf, maxD)
    assert np.array_equal(obj_ur >= 0, our2 >= 0) and np.all(np.abs(obj_ur - our2) <= 1e-3)
    assert np.allclose(obj_dp, odp2, rtol=1e-5, atol=0) and (our2 >= 0).sum() > 10

    # Frame::ExtractObjORB -> OpencvORBDetector (Frame.cc:2623-2665): cv::ORB::create(1000, 1.2, 8, 19) on both views inside
    # the object masks, through the drop-in function
    mask = np.zeros((H, W), np.uint8)
    mask[H // 4:3 * H // 4, W // 6:W // 2] = 255
    mask[H // 8:H // 2, 2 * W // 3:W - 40] = 255
    for img, k, d in ((L, okl, odl), (R, okr, odr)):
        rk, rd = O.cvorb_detect_and_compute(img, mask)
        assert k.tobytes() == rk.tobytes() and np.array_equal(d, rd) and len(k) > 100

    # Frame::AssignFeatures (Frame.cc:759-977) through the drop-in Frame: background set, per-detection ORB and dense sets in
    # the reference's push_back order, effective-depth counts, per-object grids - against the oracle (pinned to the
    # reference's own function in tests/test_oracle_reference.py)
    def same_set(got, k, d, u, z, idx, what):
        gk, gkun, gd, gu, gz = got
        assert gk.tobytes() == k[idx].tobytes() and gkun.tobytes() == k[idx].tobytes(), what
        assert np.array_equal(gd, d[idx]) and np.array_equal(gu, u[idx]) and np.array_equal(gz, z[idx]), what
    assert n_ol2 == len(okl)
    oour, oodp = O.stereo_match(oL, oR, okl, odl, okr, odr, bf, maxD)
    assert np.array_equal(f_our >= 0, oour >= 0) and np.all(np.abs(f_our - oour) <= 1e-3)
    _, groups, _ = O.assign_features(kL, seg, det_ids)
    _, ogroups, oeff = O.assign_features(okl, seg, det_ids, depth=f_odp, depth_limit=2 * 35.0)
    same_set(bg_set, kL, dL, ur, dp, groups[0], "background")
    kept = [j for j in range(len(det_ids)) if det_ids[j] >= 0]
    assert n_kept == len(kept) and n_bg > 100
    for (head, ori, dense), j in zip(det_sets, kept):
        assert head[0] == det_ids[j] and head[1] == len(groups[1 + j]) and head[2] == len(ogroups[1 + j])
        same_set(ori, kL, dL, ur, dp, groups[1 + j], f"ORB features of detection {j}")
        same_set(dense, okl, odl, f_our, f_odp, ogroups[1 + j], f"dense features of detection {j}")
        assert head[3] == oeff[j] and head[4] == len(ogroups[1 + j])
        cs, idx = O.assign_grid(okl[ogroups[1 + j]], 0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H))
        cell_of = np.repeat(np.arange(64 * 48), np.diff(cs))
        assert head[5] == len(idx) and head[6] == int((idx.astype(np.int64) * (cell_of + 1)).sum())
    assert sum(h[2] for h, _, _ in det_sets) > 300 and sum(h[1] for h, _, _ in det_sets) > 100

    # Frame::ComputeBoW (Frame.cc:2040-2047) through the drop-in ORBVocabulary (loadFromTextFile + transform)
    bid, bval, fn, fs, ff = O.bow_vectors(voc, *O.bow_transform(voc, odL, 4))
    assert np.array_equal(bow["id"], bid) and np.array_equal(bow["val"].view(np.uint64), bval.view(np.uint64)) and n_words > 100
    assert [f[0] for f in featvec] == fn.tolist()
    for j, (_, feats) in enumerate(featvec):
        assert np.array_equal(feats, ff[fs[j]:fs[j + 1]])
