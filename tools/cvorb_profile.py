#!/usr/bin/env python3
"""GPU box, under ncu: a few batched cv::ORB calls (pslot_cvorb_detect_and_compute_batch) for the per-kernel launch list.
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/cvorb_profile.py [batch]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H = 1242, 375
imgs = [synth_pair(W, H, s)[0] for s in range(8)]
mask = np.zeros((H, W), np.uint8); mask[H // 4:3 * H // 4, W // 6:W // 2] = 255; mask[H // 8:H // 2, 2 * W // 3:W - 40] = 255
orb = B.CvOrb(1000, 1.2, 8, 19)
batch = [imgs[i % 8] for i in range(nb)]
for masks in (None, [mask] * nb):
    for _ in range(2):
        out = orb.detect_and_compute_batch(batch, masks)
print(nb, "images,", sum(len(k) for k, _ in out), "keypoints")
