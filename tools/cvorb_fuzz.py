"""Randomised parity sweep of the cv::ORB oracle (oracle/orb_oracle.c: orc_cvorb_detect_and_compute) against OpenCV 4.13's
own implementation behind cv2.ORB_create (the library call of reference src/Frame.cc:2623-2665).  CPU only; test
infrastructure.

Every case draws an image shape, ORB parameters (nfeatures, scaleFactor, nlevels, edgeThreshold, fastThreshold), an image
content class and (two cases in three) an object mask of random rectangles as Frame::ExtractObjORB builds it (255 inside the
instance regions, 0 elsewhere).  Keypoints are compared as sets ordered by (octave, y, x) - cv::ORB's own output order is
whatever std::nth_element leaves - with every field and all 256 descriptor bits equal.

    python tools/cvorb_fuzz.py [cases=300] [seed=0] [ref] [nomask] [anymask] > profiles/r02_cvorb_fuzz.txt
("ref": parameters fixed to the reference's own call, cv::ORB::create(1000, 1.2, 8, 19), fastThreshold 20; "nomask": no
masks at all; "anymask": one mask rectangle in five holds a value in 1..254 instead of 255 - the reference itself only ever
builds 0 / 255 masks, Frame.cc:2637-2646)
"""
import sys
import time
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import oracle as O  # noqa: E402
from pointslot_b200.synth import synth_pair  # noqa: E402

CONTENT = ["synth", "noise", "smooth_noise", "checker", "posterised", "low_contrast", "saturated", "flat"]


def make_image(kind, W, H, rng, seed):
    if kind == "synth":
        return synth_pair(W, H, seed)[0]
    if kind == "noise":
        return rng.integers(0, 256, (H, W), dtype=np.uint8)
    if kind == "smooth_noise":
        a = cv2.GaussianBlur(rng.integers(0, 256, (H, W)).astype(np.float32), (0, 0), float(rng.uniform(0.8, 3.0)))
        return np.clip(np.rint((a - a.mean()) * 4 + 128), 0, 255).astype(np.uint8)
    if kind == "checker":
        cx, cy = int(rng.integers(4, 30)), int(rng.integers(4, 30))
        yy, xx = np.mgrid[0:H, 0:W]
        return ((((yy // cy) + (xx // cx)) & 1) * int(rng.integers(40, 236)) + 20).astype(np.uint8)
    if kind == "posterised":
        q = int(rng.choice([16, 32, 64]))
        return (synth_pair(W, H, seed)[0] // q * q).astype(np.uint8)
    if kind == "low_contrast":
        return (100 + synth_pair(W, H, seed)[0].astype(np.int32) // 5).astype(np.uint8)
    if kind == "saturated":
        return np.clip(synth_pair(W, H, seed)[0].astype(np.int32) * 3 - 200, 0, 255).astype(np.uint8)
    if kind == "flat":
        a = np.full((H, W), int(rng.integers(0, 256)), np.uint8)
        a[H // 3: H // 3 + 9, W // 4: W // 4 + 9] ^= 0x80
        return a
    raise ValueError(kind)


def make_mask(W, H, rng, any_value):
    m = np.zeros((H, W), np.uint8)
    for _ in range(int(rng.integers(1, 6))):
        w, h = int(rng.integers(8, max(9, W // 2))), int(rng.integers(8, max(9, H // 2)))
        x, y = int(rng.integers(0, W - w + 1)), int(rng.integers(0, H - h + 1))
        m[y:y + h, x:x + w] = 255 if not any_value or rng.random() < 0.8 else int(rng.integers(1, 255))
    return m


def cv2_sorted(image, mask, p):
    kps, desc = cv2.ORB_create(p[0], p[1], p[2], p[3], fastThreshold=p[4]).detectAndCompute(image, mask)
    rec = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response) for k in kps], np.float32).reshape(-1, 5)
    octv = np.array([k.octave for k in kps], np.int32)
    order = sorted(range(len(kps)), key=lambda i: (octv[i], rec[i, 1], rec[i, 0]))
    return rec[order], octv[order], (desc[order] if len(kps) else np.zeros((0, 32), np.uint8))


def draw_case(rng):
    W = int(rng.integers(64, 1300))
    H = int(rng.integers(64, 700))
    if rng.random() < 0.15:
        W, H = H, W
    nlevels = int(rng.integers(1, 10))
    scale = float(rng.choice([1.1, 1.2, 1.2, 1.25, 1.3, 1.5, round(float(rng.uniform(1.05, 1.6)), 3)]))
    nfeat = int(rng.choice([50, 200, 500, 1000, 1000, 2000, int(rng.integers(10, 4000))]))
    edge = int(rng.choice([19, 19, 31, 16, int(rng.integers(19, 40))]))  # 16: refused (as by pslot_cvorb_create)
    fast = int(rng.choice([20, 20, 10, 5, int(rng.integers(3, 50))]))
    return W, H, (nfeat, scale, nlevels, edge, fast)


def run(cases, seed0, flags=frozenset(), verbose=True):
    """-> dict(cases, refused, keypoints, masked, mismatches, by_content)"""
    ref_params = "ref" in flags
    O.build()
    cv2.setNumThreads(1)
    rng = np.random.default_rng(seed0)
    done = refused = kp = masked = mism = 0
    by = {}
    n = 0
    while done < cases:
        n += 1
        W, H, p = draw_case(rng)
        if ref_params:
            p = (1000, 1.2, 8, 19, 20)
        kind = CONTENT[n % len(CONTENT)]
        img = make_image(kind, W, H, rng, seed0 * 100003 + n)
        mask = make_mask(W, H, rng, "anymask" in flags) if n % 3 and "nomask" not in flags else None
        try:
            k, d = O.cvorb_detect_and_compute(img, mask, *p)
        except ValueError:
            refused += 1  # parameter ranges the restatement (and pslot_cvorb_create) refuses
            continue
        rec, octv, rdesc = cv2_sorted(img, mask, p)
        bad = []
        if len(k) != len(rec):
            bad.append(f"{len(k)} vs {len(rec)} keypoints")
        else:
            got = np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"]], 1) if len(k) else rec
            for c, name in enumerate(("x", "y", "size", "angle", "response")):
                if not np.array_equal(got[:, c], rec[:, c]):
                    bad.append(name)
            if not np.array_equal(k["octave"], octv):
                bad.append("octave")
            if not np.array_equal(d, rdesc):
                bad.append("descriptors")
        done += 1
        kp += len(rec)
        masked += mask is not None
        c = by.setdefault(kind, [0, 0]); c[0] += 1; c[1] += len(rec)
        if bad:
            mism += 1
            if verbose:
                print(f"MISMATCH case {n}: {W}x{H} {p} {kind} mask={'yes' if mask is not None else 'no'}: {', '.join(bad)}", flush=True)
    return dict(cases=done, refused=refused, keypoints=kp, masked=masked, mismatches=mism, by_content=by)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    flags = set(sys.argv[3:])
    t0 = time.time()
    s = run(cases, seed0, flags)
    print(f"cv::ORB oracle vs cv2.ORB_create (OpenCV {cv2.__version__}), randomised ({' '.join(sorted(flags)) or 'random parameters'}): "
          f"{s['cases']} cases (seed {seed0}, {s['masked']} with an object mask), {s['refused']} draws outside the parameter range the "
          f"restatement and pslot_cvorb_create accept, {s['keypoints']} keypoints, {s['mismatches']} mismatching cases, {time.time() - t0:.0f} s")
    print("content class: cases / keypoints")
    for name, (a, b) in sorted(s["by_content"].items()):
        print(f"  {name:14s} {a:4d} {b:8d}")
    sys.exit(1 if s["mismatches"] else 0)


if __name__ == "__main__":
    main()
