#!/usr/bin/env python3
"""GPU box: pslot_cvorb_detect_and_compute (cv::ORB for the dense object features, reference Frame.cc:2623-2665) through the
C ABI with host images, next to OpenCV's own cv::ORB (cv2, when importable) and the C oracle on the host cores.
usage: python tools/cvorb_bench.py [reps]"""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
out = {}
for name, (W, H) in (("kitti 1242x375", (1242, 375)), ("euroc 752x480", (752, 480))):
    imgs = [synth_pair(W, H, s)[0] for s in range(8)]
    mask = np.zeros((H, W), np.uint8); mask[H // 4:3 * H // 4, W // 6:W // 2] = 255; mask[H // 8:H // 2, 2 * W // 3:W - 40] = 255
    orb = B.CvOrb(1000, 1.2, 8, 19)
    res = {}
    for label, m in (("no mask", None), ("object mask", mask)):
        for i in range(10):
            orb.detect_and_compute(imgs[i % 8], m)
        t0 = time.perf_counter()
        n = 0
        for i in range(reps):
            k, _ = orb.detect_and_compute(imgs[i % 8], m)
            n += len(k)
        dev_ms = (time.perf_counter() - t0) / reps * 1e3
        r = {"device_ms_per_image": dev_ms, "keypoints_per_image": n / reps}
        try:
            import cv2
            cv2.setNumThreads(1)
            c = cv2.ORB_create(1000, 1.2, 8, 19)
            t0 = time.perf_counter()
            for i in range(20):
                c.detectAndCompute(imgs[i % 8], m)
            r["cv2_1_thread_ms_per_image"] = (time.perf_counter() - t0) / 20 * 1e3
        except ImportError:
            pass
        import oracle as O
        t0 = time.perf_counter()
        for i in range(10):
            O.cvorb_detect_and_compute(imgs[i % 8], m)
        r["oracle_c_ms_per_image"] = (time.perf_counter() - t0) / 10 * 1e3
        res[label] = r
    out[name] = res
print(json.dumps(out, indent=1))
