#!/usr/bin/env python3
"""Single-pair latency through the reference-shaped API (what Tracking sees per frame):
ORBextractor::operator() on both views (two host threads, as Frame.cc:161-164) + ComputeStereoMatches."""
import sys, time, threading
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair

def main():
    params = (2000, 1.2, 8, 20, 7)
    bf, fx = 384.38148, 721.5377
    maxD = B.max_disparity(bf, fx)
    L, R = synth_pair(1242, 375, 0)
    eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
    def frame(threads=True):
        out = {}
        if threads:
            t = threading.Thread(target=lambda: out.__setitem__("r", eR.extract(R)))
            t.start(); kL, dL = eL.extract(L); t.join(); kR, dR = out["r"]
        else:
            kL, dL = eL.extract(L); kR, dR = eR.extract(R)
        return B.stereo_match(eL, eR, kL, dL, kR, dR, bf, maxD)
    def frame_one_thread():
        eL.extract_begin(L); eR.extract_begin(R)
        kL, dL = eL.extract_end(); kR, dR = eR.extract_end()
        return B.stereo_match(eL, eR, kL, dL, kR, dR, bf, maxD)
    for _ in range(20): frame_one_thread()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); frame_one_thread(); ts.append(time.perf_counter() - t0)
    print(f"frame latency (one host thread, both views in flight): median {np.median(ts)*1e3:.3f} ms")
    def frame_resident():
        eL.extract_begin(L); eR.extract_begin(R)
        kL, dL = eL.extract_end(); kR, dR = eR.extract_end()
        return B.stereo_match_resident(eL, eR, B.extract_token(eL), B.extract_token(eR), bf, maxD)
    for _ in range(20): frame_resident()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); frame_resident(); ts.append(time.perf_counter() - t0)
    print(f"frame latency (one host thread, stereo on the resident feature sets): median {np.median(ts)*1e3:.3f} ms")
    for mode in (True, False):
        for _ in range(20): frame(mode)
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); frame(mode); ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print(f"frame latency ({'2 host threads' if mode else 'sequential'}): median {np.median(ts):.3f} ms, p95 {np.percentile(ts,95):.3f} ms")
    t0 = time.perf_counter()
    for _ in range(200): eL.extract(L)
    print(f"single extract: {(time.perf_counter()-t0)/200*1e3:.3f} ms")
    # the same pair as ONE submit of the batch API (both views + stereo on one stream, one result round trip)
    sb = B.StereoBatch(params, 1242, 375, 1)
    out = sb.alloc_out(1, pinned=True)
    Ls, Rs = np.stack([L, L])[:1].copy(), np.stack([R, R])[:1].copy()
    for _ in range(20):
        sb.submit(Ls, Rs, bf, maxD, out); sb.wait()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); sb.submit(Ls, Rs, bf, maxD, out); sb.wait(); ts.append(time.perf_counter() - t0)
    print(f"pair as a 1-pair batch submit: median {np.median(ts)*1e3:.3f} ms")
if __name__ == "__main__":
    main()
