"""cv::ORB::create(1000, 1.2, 8, 19)->detectAndCompute(image, mask, ...) — the second extractor PointSLOT runs per frame
for dense object features (reference src/Frame.cc:2623-2627 `OpencvORBDetector`, called from `Frame::ExtractObjORB`
:2628-2665) — restated with OpenCV 4.13 in the loop, the way oracle/cv2_path.py restates the
ORB-SLAM extractor: the image PRIMITIVES (resize INTER_LINEAR_EXACT, threshold, copyMakeBorder, FAST with mask,
sepFilter2D, fastAtan2) are served by cv2, everything ORB-specific (level sizes and feature quotas, border filter,
retainBest before and after the Harris score, the Harris response in its exact float order, intensity-centroid
angle, float-steered BRIEF) is written out.  Test infrastructure only.

Unlike every other oracle here, this one can be pinned against the real thing: the algorithm lives in OpenCV's
modules/features2d/src/orb.cpp (ORB_Impl::detectAndCompute, computeKeyPoints, HarrisResponses, ICAngles,
computeOrbDescriptors), and cv2.ORB_create is that code.  tests/test_oracle_cvorb.py compares the two bit for bit
(keypoint position / size / angle / response / octave and descriptors; as SETS ordered by (octave, y, x): the order
cv::ORB emits is whatever std::nth_element + std::partition leave behind inside KeyPointsFilter::retainBest).
Two facts the comparison established (they differ from the ORB-SLAM extractor and matter for a device version):
  * the pyramid is built with INTER_LINEAR_EXACT, each level from the previous one, in a buffer with a 23-px border;
  * GaussianBlur runs on a SUBMATRIX of that buffer, which makes OpenCV skip its bit-exact fixed-point path and use the
    float separable filter (== cv2.sepFilter2D with getGaussianKernel(7, 2, CV_32F)): results differ by +-1 grey level
    from the blur ORBextractor.cc gets on its cloned level.
The CUDA side of this row is not built yet."""
from __future__ import annotations

import math
import re
from pathlib import Path

import numpy as np

from .cv2_path import _cosf, _sinf

F = np.float32
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])
HARRIS_BLOCK_SIZE, HARRIS_K = 9, F(0.04)


def _cv_round(v) -> int:
    return int(np.rint(np.float64(v)))  # cvRound: round half to even


def _pattern() -> np.ndarray:
    """bit_pattern_31_ (orb.cpp; identical to ORBextractor.cc:149-406): [256][x0, y0, x1, y1]."""
    txt = (Path(__file__).resolve().parents[1] / "include" / "pslot_orb_pattern.inc").read_text()
    nums = [int(x) for x in re.findall(r"-?\d+", re.sub(r"/\*.*?\*/", "", txt, flags=re.S))]
    return np.array(nums[:1024], np.float32).reshape(256, 4)


def _umax(half_patch: int):
    umax = [0] * (half_patch + 2)
    vmax = int(math.floor(half_patch * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(half_patch * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = _cv_round(math.sqrt(half_patch * half_patch - v * v))
    v0 = 0
    for v in range(half_patch, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


def _retain_best(k, n):
    """KeyPointsFilter::retainBest: everything whose response is >= the n-th largest one (ties are kept)."""
    if n < 0 or len(k) <= n:
        return k
    if n == 0:
        return []
    thr = sorted((x[2] for x in k), reverse=True)[n - 1]
    return [x for x in k if x[2] >= thr]


def _harris(ext, border, pts, block=7):
    """HarrisResponses (orb.cpp): integer gradient sums over a 7x7 block, float score in OpenCV's evaluation order."""
    r = block // 2
    scale = F(F(1.0) / F(F(4 * block) * F(255.0)))
    s4 = F(F(F(scale * scale) * scale) * scale)
    img = ext.astype(np.int32)
    out = []
    for (x, y, _) in pts:
        cy, cx = _cv_round(y) + border, _cv_round(x) + border
        P = img[cy - r - 1:cy + r + 2, cx - r - 1:cx + r + 2]
        Ix = (P[1:-1, 2:] - P[1:-1, :-2]) * 2 + (P[:-2, 2:] - P[:-2, :-2]) + (P[2:, 2:] - P[2:, :-2])
        Iy = (P[2:, 1:-1] - P[:-2, 1:-1]) * 2 + (P[2:, :-2] - P[:-2, :-2]) + (P[2:, 2:] - P[:-2, 2:])
        a, b, c = F(int((Ix * Ix).sum())), F(int((Iy * Iy).sum())), F(int((Ix * Iy).sum()))
        sm = F(a + b)
        resp = F(F(F(F(a * b) - F(c * c)) - F(F(HARRIS_K * sm) * sm)) * s4)
        out.append((x, y, float(resp)))
    return out


def _ic_angle(cv2, ext, border, x, y, umax, hp):
    cy, cx = _cv_round(y) + border, _cv_round(x) + border
    m01 = m10 = 0
    for u in range(-hp, hp + 1):
        m10 += u * int(ext[cy, cx + u])
    for v in range(1, hp + 1):
        d, vs = umax[v], 0
        for u in range(-d, d + 1):
            p, m = int(ext[cy + v, cx + u]), int(ext[cy - v, cx + u])
            vs += p - m
            m10 += u * (p + m)
        m01 += v * vs
    return float(cv2.fastAtan2(float(m01), float(m10)))


def detect_and_compute(image: np.ndarray, mask: np.ndarray | None, nfeatures=1000, scale_factor=1.2, nlevels=8,
                       edge_threshold=19, fast_threshold=20, patch_size=31):
    """-> (keypoints [n] KEYPOINT_DTYPE, descriptors [n, 32]) ordered by (octave, y, x)."""
    import cv2
    sf = float(F(scale_factor))  # ORB::create takes a float; the member is a double holding that float
    half_patch = patch_size // 2
    desc_patch = int(math.ceil(half_patch * math.sqrt(2.0)))
    border = max(edge_threshold, max(desc_patch, HARRIS_BLOCK_SIZE // 2)) + 1
    # pyramid (ORB_Impl::detectAndCompute): level sizes from the layer scale, each level resized from the previous one
    scales, levels, masks = [], [], []
    if mask is not None:  # cv::ORB looks at the mask as zero / non-zero: any non-zero value acts as 255 on every level
        mask = np.where(mask != 0, 255, 0).astype(np.uint8)
    prev, prev_m = image, mask
    for level in range(nlevels):
        s = F(math.pow(sf, float(level)))
        scales.append(s)
        # level size = cvRound(dim * (1.f / scale)): one float reciprocal, then a float product - what OpenCV 4.13 (cv2) is
        # observed to compute; a float division differs where dim / scale lands on x.5 (465 / 1.2f -> 387,
        # 465 * (1 / 1.2f) -> 388): found by tools/cvorb_fuzz.py
        inv_s = F(F(1.0) / s)
        size = (_cv_round(F(F(image.shape[1]) * inv_s)), _cv_round(F(F(image.shape[0]) * inv_s)))
        cur, cur_m = image, mask
        if level:
            cur = cv2.resize(prev, size, interpolation=cv2.INTER_LINEAR_EXACT)
            if mask is not None:
                cur_m = cv2.resize(prev_m, size, interpolation=cv2.INTER_LINEAR_EXACT)
                _, cur_m = cv2.threshold(cur_m, 254, 0, cv2.THRESH_TOZERO)
        levels.append((cur, cv2.copyMakeBorder(cur, border, border, border, border, cv2.BORDER_REFLECT_101)))
        masks.append(cur_m)
        prev, prev_m = cur, cur_m
    # features per level (computeKeyPoints)
    factor = F(1.0 / sf)
    nd = F(F(nfeatures) * F(F(1) - factor) / F(F(1) - F(math.pow(float(factor), float(nlevels)))))
    quota, total = [], 0
    for _ in range(nlevels - 1):
        quota.append(_cv_round(nd)); total += quota[-1]; nd = F(nd * factor)
    quota.append(max(nfeatures - total, 0))
    umax = _umax(half_patch)
    fd = cv2.FastFeatureDetector_create(fast_threshold, True)
    kk = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    pat = _pattern()
    deg2rad = F(np.float64(math.pi) / np.float64(F(180.0)))
    kps, descs = [], []
    for level, (cur, ext) in enumerate(levels):
        h, w = cur.shape
        k = [(p.pt[0], p.pt[1], p.response) for p in fd.detect(cur, masks[level])]
        k = [p for p in k if edge_threshold <= p[0] < w - edge_threshold and edge_threshold <= p[1] < h - edge_threshold]
        k = _retain_best(k, 2 * quota[level])            # by FAST score
        k = _retain_best(_harris(ext, border, k), quota[level])  # by Harris score
        blurred = cv2.sepFilter2D(cur, cv2.CV_8U, kk, kk, borderType=cv2.BORDER_REFLECT_101)
        s = scales[level]
        inv = F(F(1.0) / s)
        for (x, y, resp) in sorted(k, key=lambda p: (p[1], p[0])):
            ang = _ic_angle(cv2, ext, border, x, y, umax, half_patch)
            px, py = F(F(x) * s), F(F(y) * s)
            kps.append((px, py, F(F(patch_size) * s), ang, resp, level, -1))
            cy, cx = _cv_round(F(py * inv)), _cv_round(F(px * inv))
            a_ = F(F(ang) * deg2rad)
            a, b = _cosf(float(a_)), _sinf(float(a_))  # libm cosf / sinf, as OpenCV's orb.cpp calls them

            def sample(qx, qy):
                fx = (qx * a).astype(np.float32) - (qy * b).astype(np.float32)
                fy = (qx * b).astype(np.float32) + (qy * a).astype(np.float32)
                ix, iy = np.rint(fx.astype(np.float64)).astype(int), np.rint(fy.astype(np.float64)).astype(int)
                return blurred[cy + iy, cx + ix].astype(int)
            bits = (sample(pat[:, 0], pat[:, 1]) < sample(pat[:, 2], pat[:, 3])).astype(np.uint8)
            descs.append(np.packbits(bits, bitorder="little"))
    out = np.array(kps, KEYPOINT_DTYPE) if kps else np.zeros(0, KEYPOINT_DTYPE)
    return out, (np.array(descs, np.uint8) if descs else np.zeros((0, 32), np.uint8))


# ---------------------------------------------------------------------------------------------------------------------
# The two OpenCV primitives of this path that the ORB-SLAM extractor does not use, restated (and pinned against cv2 by
# tests/test_oracle_cvorb.py) so that a device version has an arithmetic to follow.
# ---------------------------------------------------------------------------------------------------------------------
def resize_linear_exact(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv::resize(..., INTER_LINEAR_EXACT) for CV_8UC1 (resize.cpp, the bit-exact fixed-point path): per axis
    f = (d + 0.5) * (n_src / n_dst) - 0.5, s = floor(f) clamped to the image, 8-bit coefficient a = round(frac * 256);
    H = s0 * (256 - ax) + s1 * ax (16 bit), dst = (H0 * (256 - ay) + H1 * ay + 2^15) >> 16."""
    def taps(n_dst, n_src):
        scale = n_src / n_dst
        s0 = np.zeros(n_dst, np.int64); a = np.zeros(n_dst, np.int64)
        for d in range(n_dst):
            f = (d + 0.5) * scale - 0.5
            s = math.floor(f); fr = f - s
            if s < 0:
                s, fr = 0, 0.0
            if s >= n_src - 1:
                s, fr = n_src - 1, 0.0
            s0[d] = s; a[d] = int(np.rint(fr * 256))
        return s0, a
    sh, sw = src.shape
    x0, ax = taps(dw, sw); y0, ay = taps(dh, sh)
    x1, y1 = np.minimum(x0 + 1, sw - 1), np.minimum(y0 + 1, sh - 1)
    S = src.astype(np.int64)
    H = S[:, x0] * (256 - ax)[None, :] + S[:, x1] * ax[None, :]
    V = H[y0, :] * (256 - ay)[:, None] + H[y1, :] * ay[:, None]
    return ((V + (1 << 15)) >> 16).astype(np.uint8)


def blur_float_separable(img: np.ndarray, fma: bool = True) -> np.ndarray:
    """What cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) computes on a SUBMATRIX (== cv2.sepFilter2D with the
    float kernel getGaussianKernel(7, 2, CV_32F)): float32 row filter accumulated left to right, float32 column filter
    accumulated from the centre tap outwards over the symmetric pairs, round-to-nearest-even to 8 bit.  With AVX2/FMA
    dispatch (this image's CPUs) every multiply-add is FUSED; a CPU without FMA rounds the products separately, i.e.
    cv::ORB's descriptors are not bit-reproducible across CPU generations."""
    import cv2
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel().astype(np.float32)
    f32 = lambda a: a.astype(np.float32)

    def mad(c, x, acc):  # acc + c * x in float32, fused or not
        if fma:
            return f32(np.float64(c) * x.astype(np.float64) + acc.astype(np.float64))
        return f32(acc + f32(np.float32(c) * x))
    p = cv2.copyMakeBorder(img, 3, 3, 3, 3, cv2.BORDER_REFLECT_101).astype(np.float32)
    H, W = img.shape
    row = f32(np.float32(k[0]) * p[:, 0:W])
    for i in range(1, 7):
        row = mad(k[i], p[:, i:i + W], row)
    col = f32(np.float32(k[3]) * row[3:3 + H, :])
    for i in (1, 2, 3):
        col = mad(k[3 - i], f32(row[3 - i:3 - i + H, :] + row[3 + i:3 + i + H, :]), col)
    return np.clip(np.rint(col.astype(np.float64)), 0, 255).astype(np.uint8)
