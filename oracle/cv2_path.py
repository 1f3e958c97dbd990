"""Second, independent restatement of the path with OpenCV ITSELF in the loop (test infrastructure).

Every OpenCV primitive the reference calls is executed by cv2 (4.13.0 here): cv2.resize,
cv2.FastFeatureDetector per cell, cv2.GaussianBlur, cv2.fastAtan2.  Only the reference's own glue
(ORBextractor.cc / Frame.cc) is restated, in deliberately simple Python.  Used to (1) pin the C
oracle (tests/test_oracle_cv2.py) and (2) generate tests/golden/*.npz (tests/golden/make_golden.py).
Slow (seconds per image); never on a timed or product path.
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np

EDGE = 19
HALF = 15

import ctypes as _C  # noqa: E402

_libm = _C.CDLL("libm.so.6")  # the reference binary's own cosf / sinf (glibc)
_libm.cosf.restype = _C.c_float; _libm.cosf.argtypes = [_C.c_float]
_libm.sinf.restype = _C.c_float; _libm.sinf.argtypes = [_C.c_float]


def _cosf(x: float):
    return np.float32(_libm.cosf(x))


def _sinf(x: float):
    return np.float32(_libm.sinf(x))


def f32(v):
    return np.float32(v)


def cv_round(v) -> int:
    """cvRound: half-to-even on a double."""
    return int(np.rint(np.float64(v)))


def c_round(v) -> float:
    """C round(): half away from zero."""
    v = float(v)
    return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)


def load_pattern() -> np.ndarray:
    txt = (Path(__file__).resolve().parent.parent / "include" / "pslot_orb_pattern.inc").read_text()
    txt = txt[txt.index("*/") + 2:]
    return np.array([int(t) for t in txt.replace("\n", " ").split(",") if t.strip()], np.int32).reshape(256, 4)


class Cv2Extractor:
    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        import cv2
        self.cv2 = cv2
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sfd = float(f32(scale_factor))  # double member holding the float argument
        self.sf = [f32(1.0)]
        for _ in range(1, nlevels):
            self.sf.append(f32(float(self.sf[-1]) * sfd))
        self.inv_sf = [f32(1.0) / s for s in self.sf]
        self.sigma2 = [s * s for s in self.sf]
        self.inv_sigma2 = [f32(1.0) / s for s in self.sigma2]
        factor = f32(1.0 / sfd)
        nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
        self.per_level, tot = [], 0
        for _ in range(nlevels - 1):
            self.per_level.append(cv_round(nd)); tot += self.per_level[-1]
            nd = f32(nd * factor)
        self.per_level.append(max(nfeatures - tot, 0))
        # umax
        vmax = int(math.floor(float(f32(HALF) * np.sqrt(f32(2.0)) / f32(2) + f32(1))))
        vmin = int(math.ceil(float(f32(HALF) * np.sqrt(f32(2.0)) / f32(2))))
        um = [0] * (HALF + 1)
        for v in range(vmax + 1):
            um[v] = cv_round(math.sqrt(HALF * HALF - v * v))
        v0 = 0
        for v in range(HALF, vmin - 1, -1):
            while um[v0] == um[v0 + 1]:
                v0 += 1
            um[v] = v0
            v0 += 1
        self.umax = um
        self.pattern = load_pattern()
        self.pyr, self.cands, self.lkps, self.blur = [], [], [], []

    # -- ComputePyramid (borderless: the 19 px frame is never read) --------------------------------
    def compute_pyramid(self, img):
        cv2 = self.cv2
        h, w = img.shape
        self.pyr = []
        for l in range(self.nlevels):
            lw = cv_round(f32(w) * self.inv_sf[l]); lh = cv_round(f32(h) * self.inv_sf[l])
            if l == 0:
                self.pyr.append(np.ascontiguousarray(img))
            else:
                self.pyr.append(cv2.resize(self.pyr[l - 1], (lw, lh), interpolation=cv2.INTER_LINEAR))

    # -- cell loop of ComputeKeyPointsOctTree -------------------------------------------------------
    def level_candidates(self, l):
        cv2 = self.cv2
        im = self.pyr[l]
        minBX = minBY = EDGE - 3
        maxBX = im.shape[1] - EDGE + 3; maxBY = im.shape[0] - EDGE + 3
        width = f32(maxBX - minBX); height = f32(maxBY - minBY)
        nCols = int(width / f32(30)); nRows = int(height / f32(30))
        wCell = int(math.ceil(float(width / f32(nCols)))); hCell = int(math.ceil(float(height / f32(nRows))))
        det_i = cv2.FastFeatureDetector_create(threshold=self.ini_th, nonmaxSuppression=True)
        det_m = cv2.FastFeatureDetector_create(threshold=self.min_th, nonmaxSuppression=True)
        out = []
        for i in range(nRows):
            iniY = minBY + i * hCell; maxY = iniY + hCell + 6
            if iniY >= maxBY - 3: continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minBX + j * wCell; maxX = iniX + wCell + 6
                if iniX >= maxBX - 3: continue
                maxX = min(maxX, maxBX)
                sub = im[iniY:maxY, iniX:maxX]
                k = det_i.detect(sub)
                if len(k) == 0:
                    k = det_m.detect(sub)
                for kp in k:
                    out.append((int(kp.pt[0]) + j * wCell, int(kp.pt[1]) + i * hCell, int(kp.response)))
        return out, (minBX, maxBX, minBY, maxBY)

    # -- DistributeOctTree, literal list-based ------------------------------------------------------
    @staticmethod
    def distribute(keys, minX, maxX, minY, maxY, N):
        nIni = int(c_round(float(f32(maxX - minX) / f32(maxY - minY))))
        hX = f32(maxX - minX) / f32(nIni)
        seq = [0]

        def mk(ulx, uly, urx, bry):
            seq[0] += 1
            return {"ulx": ulx, "uly": uly, "urx": urx, "bry": bry, "keys": [], "nomore": False, "seq": seq[0]}

        nodes = [mk(int(hX * f32(i)), 0, int(hX * f32(i + 1)), maxY - minY) for i in range(nIni)]
        for k in keys:
            nodes[int(f32(k[0]) / hX)]["keys"].append(k)
        nodes = [n for n in nodes if n["keys"]]
        for n in nodes:
            n["nomore"] = len(n["keys"]) == 1

        def divide(n):
            halfX = int(math.ceil(float(f32(n["urx"] - n["ulx"]) / f32(2))))
            halfY = int(math.ceil(float(f32(n["bry"] - n["uly"]) / f32(2))))
            mx, my = n["ulx"] + halfX, n["uly"] + halfY
            c = [mk(n["ulx"], n["uly"], mx, my), mk(mx, n["uly"], n["urx"], my),
                 mk(n["ulx"], my, mx, n["bry"]), mk(mx, my, n["urx"], n["bry"])]
            for k in n["keys"]:
                c[(0 if k[0] < mx else 1) + (0 if k[1] < my else 2)]["keys"].append(k)
            for ch in c:
                ch["nomore"] = len(ch["keys"]) == 1
            return c

        finish = False
        while not finish:
            prev = len(nodes)
            V, nexp, new_front, keep = [], 0, [], []
            for n in nodes:
                if n["nomore"]:
                    keep.append(n); continue
                for ch in divide(n):
                    if ch["keys"]:
                        new_front.insert(0, ch)
                        if len(ch["keys"]) > 1:
                            nexp += 1; V.append(ch)
            nodes = new_front + keep
            if len(nodes) >= N or len(nodes) == prev:
                finish = True
            elif len(nodes) + 3 * nexp > N:
                while not finish:
                    prev = len(nodes)
                    P = sorted(V, key=lambda n: (len(n["keys"]), n["seq"]))  # rule: later-created sorts last
                    V = []
                    for n in reversed(P):
                        for ch in divide(n):
                            if ch["keys"]:
                                nodes.insert(0, ch)
                                if len(ch["keys"]) > 1:
                                    V.append(ch)
                        nodes.remove(n)
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prev:
                        finish = True
        res = []
        for n in nodes:
            best = n["keys"][0]
            for k in n["keys"][1:]:
                if k[2] > best[2]:
                    best = k
            res.append(best)
        return res

    def ic_angle(self, im, x, y):
        cv2 = self.cv2
        m01 = m10 = 0
        row = im[y].astype(np.int64)
        for u in range(-HALF, HALF + 1):
            m10 += u * int(row[x + u])
        for v in range(1, HALF + 1):
            d = self.umax[v]
            p = im[y + v, x - d:x + d + 1].astype(np.int64); m = im[y - v, x - d:x + d + 1].astype(np.int64)
            us = np.arange(-d, d + 1)
            m01 += v * int((p - m).sum())
            m10 += int((us * (p + m)).sum())
        return f32(cv2.fastAtan2(float(f32(m01)), float(f32(m10))))

    def descriptor(self, blurred, x, y, angle):
        factorPI = f32(math.pi / 180.0)
        ang = f32(f32(angle) * factorPI)
        a, b = _cosf(float(ang)), _sinf(float(ang))  # `(float)cos(float)` is std::cos(float) = libm cosf (:112-113)
        d = np.zeros(32, np.uint8)
        for i in range(256):
            x0, y0, x1, y1 = (f32(v) for v in self.pattern[i])
            t0 = blurred[y + cv_round(f32(f32(x0 * b) + f32(y0 * a))), x + cv_round(f32(f32(x0 * a) - f32(y0 * b)))]
            t1 = blurred[y + cv_round(f32(f32(x1 * b) + f32(y1 * a))), x + cv_round(f32(f32(x1 * a) - f32(y1 * b)))]
            if t0 < t1:
                d[i // 8] |= 1 << (i % 8)
        return d

    def extract(self, img):
        cv2 = self.cv2
        self.compute_pyramid(img)
        self.cands, self.lkps, self.blur = [], [], []
        kps, descs = [], []
        for l in range(self.nlevels):
            c, (minBX, maxBX, minBY, maxBY) = self.level_candidates(l)
            self.cands.append(c)
            sel = self.distribute(c, minBX, maxBX, minBY, maxBY, self.per_level[l])
            sel = [(k[0] + minBX, k[1] + minBY, k[2]) for k in sel]
            sel = [(x, y, r, self.ic_angle(self.pyr[l], x, y)) for (x, y, r) in sel]
            self.lkps.append(sel)
        for l in range(self.nlevels):
            if not self.lkps[l]:
                self.blur.append(None); continue
            bl = cv2.GaussianBlur(self.pyr[l].copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            self.blur.append(bl)
            size = f32(int(f32(31) * self.sf[l]))
            for (x, y, r, ang) in self.lkps[l]:
                descs.append(self.descriptor(bl, x, y, ang))
                px, py = f32(x), f32(y)
                if l != 0:
                    px = f32(px * self.sf[l]); py = f32(py * self.sf[l])
                kps.append((px, py, size, ang, f32(r), l, -1))
        from . import KEYPOINT_DTYPE
        k = np.array(kps, KEYPOINT_DTYPE) if kps else np.zeros(0, KEYPOINT_DTYPE)
        d = np.array(descs, np.uint8).reshape(-1, 32)
        return k, d


def stereo_match_cv2(eL: Cv2Extractor, eR: Cv2Extractor, kL, dL, kR, dR, bf, maxD):
    """Frame::ComputeStereoMatches (Frame.cc:2142-2316), cv2.norm for the SAD as in the reference."""
    import cv2
    nL = len(kL)
    ur = np.full(nL, -1, np.float32); dp = np.full(nL, -1, np.float32)
    nRows = eL.pyr[0].shape[0]
    rows = [[] for _ in range(nRows)]
    for iR in range(len(kR)):
        y = f32(kR["y"][iR]); r = f32(2.0) * eL.sf[int(kR["octave"][iR])]
        for yi in range(int(math.floor(float(y - r))), int(math.ceil(float(y + r))) + 1):
            rows[yi].append(iR)
    bits = np.unpackbits(dR, axis=1) if len(dR) else np.zeros((0, 256), np.uint8)
    bitsL = np.unpackbits(dL, axis=1) if len(dL) else np.zeros((0, 256), np.uint8)
    maxD = f32(maxD); bf = f32(bf)
    vd = []
    for iL in range(nL):
        lvl = int(kL["octave"][iL]); vL = f32(kL["y"][iL]); uL = f32(kL["x"][iL])
        cand = rows[int(vL)]
        if not cand: continue
        minU = f32(uL - maxD); maxU = uL
        if maxU < 0: continue
        best, bestR = 100, 0
        for iR in cand:
            o = int(kR["octave"][iR])
            if o < lvl - 1 or o > lvl + 1: continue
            uR = f32(kR["x"][iR])
            if uR >= minU and uR <= maxU:
                dist = int((bitsL[iL] != bits[iR]).sum())
                if dist < best:
                    best, bestR = dist, iR
        if best >= 75: continue
        uR0 = f32(kR["x"][bestR]); s = eL.inv_sf[lvl]
        su = c_round(f32(uL * s)); sv = c_round(f32(vL * s)); sr = c_round(f32(uR0 * s))
        w = 5
        su, sv, sr = int(su), int(sv), int(sr)
        IL = eL.pyr[lvl][sv - w:sv + w + 1, su - w:su + w + 1].astype(np.float32)
        IL = IL - IL[w, w]
        if sr < 0 or sr + 11 >= eR.pyr[lvl].shape[1]: continue
        bestD, bestInc, dists = 2 ** 31 - 1, 0, [f32(0)] * 11
        for inc in range(-5, 6):
            IR = eR.pyr[lvl][sv - w:sv + w + 1, sr + inc - w:sr + inc + w + 1].astype(np.float32)
            IR = IR - IR[w, w]
            dist = f32(cv2.norm(IL, IR, cv2.NORM_L1))
            if dist < bestD:
                bestD, bestInc = int(dist), inc
            dists[5 + inc] = dist
        if bestInc in (-5, 5): continue
        d1, d2, d3 = dists[5 + bestInc - 1], dists[5 + bestInc], dists[5 + bestInc + 1]
        with np.errstate(all="ignore"):
            delta = f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2)))
        if not (delta >= -1 and delta <= 1):  # NaN falls through in the reference and dies on disparity
            if not np.isnan(delta): continue
        bestuR = f32(eL.sf[lvl] * f32(f32(f32(sr) + f32(bestInc)) + delta))
        disp = f32(uL - bestuR)
        if disp >= 0 and disp < maxD:
            if disp <= 0:
                disp = f32(0.01); bestuR = f32(float(uL) - 0.01)
            dp[iL] = f32(bf / disp); ur[iL] = bestuR
            vd.append((bestD, iL))
    if vd:
        vd.sort()
        median = f32(vd[len(vd) // 2][0])
        th = f32(f32(f32(1.5) * f32(1.4)) * median)
        for dist, iL in reversed(vd):
            if f32(dist) < th: break
            ur[iL] = -1; dp[iL] = -1
    return ur, dp
