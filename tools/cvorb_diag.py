#!/usr/bin/env python3
"""GPU box: compare pslot_cvorb_detect_and_compute with the C oracle (and cv2 when present) level by level."""
import sys, collections
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle as O
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1242, 375)
L, _ = synth_pair(W, H, 3)
k, d = B.CvOrb().detect_and_compute(L)
ok, od = O.cvorb_detect_and_compute(L, None)
print("device", len(k), "oracle", len(ok))
for l in range(8):
    a = k[k["octave"] == l]; b = ok[ok["octave"] == l]
    da = d[k["octave"] == l]; db = od[ok["octave"] == l]
    sa = {(float(p["x"]), float(p["y"])): i for i, p in enumerate(a)}
    sb = {(float(p["x"]), float(p["y"])): i for i, p in enumerate(b)}
    common = sorted(set(sa) & set(sb))
    nresp = sum(a[sa[c]]["response"] != b[sb[c]]["response"] for c in common)
    nang = sum(a[sa[c]]["angle"] != b[sb[c]]["angle"] for c in common)
    ndesc = sum(not np.array_equal(da[sa[c]], db[sb[c]]) for c in common)
    srt = all((a["y"][i], a["x"][i]) <= (a["y"][i + 1], a["x"][i + 1]) for i in range(len(a) - 1))
    print(f"level {l}: device {len(a)} oracle {len(b)} common {len(common)} resp!= {nresp} angle!= {nang} desc!= {ndesc} device sorted {srt}")
    if l == 0:
        for c in common[:5]:
            print("   ", c, a[sa[c]]["response"], b[sb[c]]["response"])
        only_a = sorted(set(sa) - set(sb))[:5]; only_b = sorted(set(sb) - set(sa))[:5]
        print("    only device:", [(c, float(a[sa[c]]["response"])) for c in only_a])
        print("    only oracle:", [(c, float(b[sb[c]]["response"])) for c in only_b])
        print("    min resp device", a["response"].min() if len(a) else None, "oracle", b["response"].min() if len(b) else None)

# ---- stage by stage, level 0 and 1 -------------------------------------------------------------------------------------
orb = B.CvOrb()
orb.detect_and_compute(L)
for l in (0, 1):
    img = orb.stage(0, l)
    if l == 0:
        print("level 0 image equal:", np.array_equal(img, L))
    fx, fy, fr = O.fast9_16(img, 20)
    h, w = img.shape
    keep = (fx >= 19) & (fx < w - 19) & (fy >= 19) & (fy < h - 19)
    ref = set(zip(fx[keep].tolist(), fy[keep].tolist(), fr[keep].tolist()))
    cx, cy, cs = orb.stage(2, l)
    dev = set(zip(cx.tolist(), cy.tolist(), cs.tolist()))
    print(f"level {l}: candidates device {len(dev)} (list {len(cx)}) oracle {len(ref)} common {len(dev & ref)}")
    print("   only device", sorted(dev - ref)[:6]); print("   only oracle", sorted(ref - dev)[:6])
    sm = orb.stage(1, l)
    print("   score map nonzero", int((sm > 0).sum()), "max", int(sm.max()))
    # oracle score map through its FAST without the maxima test is not exposed; check the device maxima test on its own map
    ys, xs = np.nonzero(sm[1:-1, 1:-1]); ys += 1; xs += 1
    nb = np.stack([sm[ys + dy, xs + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dy, dx) != (0, 0)])
    mx = (sm[ys, xs] > nb.max(0)) & (xs >= 19) & (xs < w - 19) & (ys >= 19) & (ys < h - 19)
    own = set(zip(xs[mx].tolist(), ys[mx].tolist(), sm[ys, xs][mx].tolist()))
    print("   maxima of the device score map:", len(own), "== device list", own == dev, "== oracle", own == ref)
    kx, ky, ks, kr = orb.stage(3, l)
    print("   kept", len(kx), "min score", ks.min() if len(ks) else None)
