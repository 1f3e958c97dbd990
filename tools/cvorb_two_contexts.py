#!/usr/bin/env python3
"""GPU box: 64-image batched cv::ORB calls (page-locked images + object masks) from ONE context, and from TWO / THREE contexts on
as many host threads (each context owns its streams, so one context's upload overlaps the other's kernels)."""
import json, sys, threading, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair

W, H, n = 1242, 375, 64
imgs = [synth_pair(W, H, s)[0] for s in range(8)]
mask = np.zeros((H, W), np.uint8); mask[H // 4:3 * H // 4, W // 6:W // 2] = 255; mask[H // 8:H // 2, 2 * W // 3:W - 40] = 255
out = {}
for nctx in (1, 2, 3):
    ctxs = []
    for _ in range(nctx):
        pin = B.PinnedBuffer((n, H, W)); pm = B.PinnedBuffer((H, W))
        for i in range(n):
            pin.array[i] = imgs[i % 8]
        pm.array[:] = mask
        ctxs.append((B.CvOrb(1000, 1.2, 8, 19), [pin.array[i] for i in range(n)], [pm.array] * n, pin, pm))
    for masked in (False, True):
        done = [0] * nctx
        stop = time.perf_counter() + 1.0

        def work(j):
            orb, batch, masks = ctxs[j][:3]
            for _ in range(3):
                orb.detect_and_compute_batch(batch, masks if masked else None, copy=False)
            while time.perf_counter() < stop:
                orb.detect_and_compute_batch(batch, masks if masked else None, copy=False)
                done[j] += 1
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(j,)) for j in range(nctx)]
        [t.start() for t in th]; [t.join() for t in th]
        dt = time.perf_counter() - t0
        out[f"{nctx} context(s), {'masks' if masked else 'no masks'}"] = round(sum(done) * n / dt)
print(json.dumps(out, indent=1))
