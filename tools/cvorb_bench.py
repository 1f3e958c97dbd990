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
        # batched entry point (pslot_cvorb_detect_and_compute_batch): the left + right pair of Frame::ExtractObjORB, and
        # whole-sequence batches; host images in, host results out, wall clock per image
        for nb in (2, 8, 32, 64):
            batch = [imgs[i % 8] for i in range(nb)]
            masks = None if m is None else [m] * nb
            for _ in range(3):
                orb.detect_and_compute_batch(batch, masks)
            rb = max(4, reps // nb)
            t0 = time.perf_counter()
            for _ in range(rb):
                orb.detect_and_compute_batch(batch, masks)
            r[f"device_ms_per_image_batch{nb}"] = (time.perf_counter() - t0) / (rb * nb) * 1e3
            r[f"abi_phases_ms_batch{nb}"] = [round(float(v), 4) for v in orb.last_timing()]
        # the same 64-image batch from page-locked host memory (pslot_host_alloc): the uploads run at link speed
        pin = B.PinnedBuffer((64, H, W)); pm = B.PinnedBuffer((H, W))
        for i in range(64):
            pin.array[i] = imgs[i % 8]
        pm.array[:] = mask
        batch = [pin.array[i] for i in range(64)]
        masks = None if m is None else [pm.array] * 64
        for _ in range(3):
            orb.detect_and_compute_batch(batch, masks)
        t0 = time.perf_counter()
        for _ in range(max(4, reps // 64)):
            orb.detect_and_compute_batch(batch, masks)
        r["device_ms_per_image_batch64_pinned"] = (time.perf_counter() - t0) / (max(4, reps // 64) * 64) * 1e3
        r["abi_phases_ms_batch64_pinned"] = [round(float(v), 4) for v in orb.last_timing()]
        del batch, masks
        pin.close(); pm.close()
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
