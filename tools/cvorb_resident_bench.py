#!/usr/bin/env python3
"""GPU box: cv::ORB on the images a StereoBatch holds on the device (pslot_cvorb_detect_and_compute_resident) - images/s for 64
images per call (32 pairs), with page-locked object masks and without, one context and two contexts on two threads."""
import json, sys, threading, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair

W, H, P = 1242, 375, 32
pairs = [synth_pair(W, H, s) for s in range(4)]
Ls = np.stack([pairs[i % 4][0] for i in range(P)]); Rs = np.stack([pairs[i % 4][1] for i in range(P)])
sb = B.StereoBatch((1000, 1.2, 8, 20, 5), W, H, P)
sb.run(Ls, Rs, 386.1448, B.max_disparity(386.1448, 718.856))
pm = B.PinnedBuffer((H, W)); pm.array[:] = 0
pm.array[H // 4:3 * H // 4, W // 6:W // 2] = 255; pm.array[H // 8:H // 2, 2 * W // 3:W - 40] = 255
out = {}
for nctx in (1, 2):
    orbs = [B.CvOrb(1000, 1.2, 8, 19) for _ in range(nctx)]
    for masked in (False, True):
        masks = [pm.array] * (2 * P) if masked else None
        done = [0] * nctx
        for o in orbs:
            for _ in range(3):
                o.detect_and_compute_resident(sb, 0, 2 * P, masks, copy=False)
        stop = time.perf_counter() + 1.0

        def work(j):
            while time.perf_counter() < stop:
                orbs[j].detect_and_compute_resident(sb, 0, 2 * P, masks, copy=False)
                done[j] += 1
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(j,)) for j in range(nctx)]
        [t.start() for t in th]; [t.join() for t in th]
        out[f"{nctx} context(s), {'masks' if masked else 'no masks'}"] = round(sum(done) * 2 * P / (time.perf_counter() - t0))
print(json.dumps(out, indent=1))
