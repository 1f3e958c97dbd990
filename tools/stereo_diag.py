"""Diagnose a stereo mismatch: relabelled-octave case of tests/test_gpu_parity.py::test_stereo_search_fallbacks.

Prints, for every left keypoint whose matched state differs between the CUDA path and the oracle, the Hamming
stage result of the CUDA path (pslot_stereo_debug) next to a numpy brute force of the same search.
Run on the GPU box:  python tools/stereo_diag.py
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair

params = (3000, 1.44, 7, 20, 7)
W, H = 1241, 752
L, R = synth_pair(W, H, 77)
bf, maxD = 384.38148, B.max_disparity(384.38148, 721.5377)
eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
oL, oR = oracle.OracleExtractor(*params), oracle.OracleExtractor(*params)
kL, dL = eL.extract(L); kR, dR = eR.extract(R)
okL, odL = oL.extract(L); okR, odR = oR.extract(R)
assert kL.tobytes() == okL.tobytes() and kR.tobytes() == okR.tobytes()


def inner(k):
    return (k["x"] > 100) & (k["x"] < W - 100) & (k["y"] > 64) & (k["y"] < H - 64)


mL, mR = inner(kL), inner(kR)
kL2, kR2, dL, dR = kL[mL].copy(), kR[mR].copy(), np.ascontiguousarray(dL[mL]), np.ascontiguousarray(dR[mR])
kL2["octave"] = 5 + (np.arange(len(kL2)) & 1); kR2["octave"] = 6
ur, dp, bd, bi, bs = B.stereo_match(eL, eR, kL2, dL, kR2, dR, bf, maxD, debug=True)
our, odp = oracle.stereo_match(oL, oR, kL2, dL, kR2, dR, bf, maxD)
bad = np.nonzero((ur >= 0) != (our >= 0))[0]
print("nL", len(kL2), "nR", len(kR2), "matched gpu", (ur >= 0).sum(), "oracle", (our >= 0).sum(), "differ", len(bad))
sf = np.float32(1.44) ** np.arange(7, dtype=np.float32)
pc = np.unpackbits(dL, axis=1)
pr = np.unpackbits(dR, axis=1)
for i in bad[:12]:
    k = kL2[i]
    r = 2.0 * sf[kR2["octave"]]
    rows = (np.floor(kR2["y"] - r) <= int(k["y"])) & (np.ceil(kR2["y"] + r) >= int(k["y"]))
    lv = np.abs(kR2["octave"] - k["octave"]) <= 1
    ux = (kR2["x"] >= k["x"] - maxD) & (kR2["x"] <= k["x"])
    c = np.nonzero(rows & lv & ux)[0]
    d = (pc[i][None, :] != pr[c]).sum(1) if len(c) else np.array([])
    j = c[np.argmin(d)] if len(c) else -1
    print(f"iL={i} xy=({k['x']:.2f},{k['y']:.2f}) oct={k['octave']}  gpu: ur={ur[i]:.3f} bestDist={bd[i]} bestIdx={bi[i]} sad={bs[i]}"
          f"  oracle ur={our[i]:.3f}  brute: n={len(c)} best={d.min() if len(c) else None} idx={j}")
