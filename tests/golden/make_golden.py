#!/usr/bin/env python3
"""Generate tests/golden/*.npz — golden input/output vectors for the stereo ORB front end.

The outputs are produced by oracle/cv2_path.py, i.e. by OpenCV 4.13.0 itself (cv2.resize,
cv2.FastFeatureDetector per cell, cv2.GaussianBlur, cv2.fastAtan2, cv2.norm) plus a literal Python
restatement of the reference's glue (ORBextractor.cc / Frame.cc:2142-2316) under the parity rules
of DESIGN.md.  The reference ships no golden vectors of its own; these pin the C
oracle and, through it, the CUDA path.

Run in the build container:  python tests/golden/make_golden.py   (needs cv2; ~2 min)
The natural-texture input is the grayscale of the reference's only image asset
(Examples/Stereo/000212.png, a 1242x375 viewer screenshot), with the right view synthesised by a
24 px shift plus N(0,1.5) noise; it is stored inside the fixture so nothing reads /root/reference
at test time.
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle.cv2_path import Cv2Extractor, stereo_match_cv2  # noqa: E402
from pointslot_b200.synth import synth_pair  # noqa: E402

OUT = Path(__file__).resolve().parent


def max_disparity(bf, fx):
    """mb = mbf/fx; maxD = mbf/mb in float32 (the value the reference intends, Frame.cc:197,2172-2174)."""
    mb = np.float32(bf) / np.float32(fx)
    return float(np.float32(bf) / mb)


def natural_pair():
    import cv2
    im = cv2.imread("/root/reference/Examples/Stereo/000212.png")
    g = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)
    shift = 24
    r = np.empty_like(g)
    r[:, :-shift] = g[:, shift:]
    r[:, -shift:] = g[:, -1:]
    noise = np.random.default_rng(212).normal(0, 1.5, g.shape)
    r = np.clip(np.rint(r.astype(np.float64) + noise), 0, 255).astype(np.uint8)
    return g, r


CASES = {
    # name: (params, (W,H,seed)|"natural", bf, fx, store_inputs)
    "kitti_synth_s0_2000": ((2000, 1.2, 8, 20, 7), (1242, 375, 0), 384.38148, 721.5377, True),
    "kitti_synth_s7_1000": ((1000, 1.2, 8, 20, 5), (1242, 375, 7), 384.38148, 721.5377, False),
    "kitti_natural_2000": ((2000, 1.2, 8, 20, 7), "natural", 384.38148, 721.5377, True),
    "euroc_synth_s3_1200": ((1200, 1.2, 8, 20, 7), (752, 480, 3), 47.9, 435.2, False),
    "odd_synth_s5_500": ((500, 1.2, 8, 20, 7), (333, 222, 5), 40.0, 300.0, True),
    "hd_synth_s1_4000": ((4000, 1.2, 8, 20, 7), (1920, 1080, 1), 384.38148, 721.5377, False),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(only=None):
    for name, (params, src, bf, fx, store) in CASES.items():
        if only and name not in only:
            continue
        L, R = natural_pair() if src == "natural" else synth_pair(*src)
        eL, eR = Cv2Extractor(*params), Cv2Extractor(*params)
        kL, dL = eL.extract(L)
        kR, dR = eR.extract(R)
        maxD = max_disparity(bf, fx)
        ur, dp = stereo_match_cv2(eL, eR, kL, dL, kR, dR, bf, maxD)
        rec = dict(
            params=np.array(params, np.float64), bf=np.float64(bf), fx=np.float64(fx), maxD=np.float64(maxD),
            src=np.array(src if src != "natural" else (-1, -1, -1), np.int64),
            sha_left=np.array(sha(L)), sha_right=np.array(sha(R)),
            kL=kL, dL=dL, kR=kR, dR=dR, uright=ur, depth=dp,
            ncand_left=np.array([len(c) for c in eL.cands], np.int64),
            nkp_left=np.array([len(c) for c in eL.lkps], np.int64),
            sha_pyr_left=np.array([sha(p) for p in eL.pyr]),
            sha_blur_left=np.array([sha(b) if b is not None else "" for b in eL.blur]),
        )
        if store:
            rec["left"], rec["right"] = L, R
        np.savez_compressed(OUT / f"{name}.npz", **rec)
        print(f"{name}: nL={len(kL)} nR={len(kR)} matched={(ur >= 0).sum()} "
              f"size={(OUT / (name + '.npz')).stat().st_size / 1e3:.0f} kB", flush=True)


if __name__ == "__main__":
    main(set(sys.argv[1:]))
