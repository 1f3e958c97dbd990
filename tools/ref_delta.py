#!/usr/bin/env python3
"""Measure how far the reference binary built with the STOCK allocator (glibc malloc; oracle/_ref/libref_orb.so) is from
the oracle / CUDA path, i.e. the size of the heap-address tie of ORBextractor.cc:684 (rule 1).
Prints one block per configuration; the committed copy is profiles/r02_ref_delta.txt.   python tools/ref_delta.py [pairs]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from oracle import ref_path  # noqa: E402
from pointslot_b200.synth import synth_pair  # noqa: E402
from test_oracle_reference import stock_delta  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for name, (W, H), params, n in (("kitti 1242x375 (1000,1.2,8,20,5)", (1242, 375), (1000, 1.2, 8, 20, 5), pairs),
                                ("kitti 1242x375 (2000,1.2,8,20,7)", (1242, 375), (2000, 1.2, 8, 20, 7), pairs),
                                ("euroc 752x480 (1200,1.2,8,20,7)", (752, 480), (1200, 1.2, 8, 20, 7), pairs // 2),
                                ("hd 1920x1080 (4000,1.2,8,20,7)", (1920, 1080), (4000, 1.2, 8, 20, 7), max(pairs // 10, 2))):
    imgs = [im for s in range(500, 500 + n) for im in synth_pair(W, H, s)]
    st = stock_delta(ref_path, oracle, params, imgs)
    o = oracle.OracleExtractor(*params)
    order = cut = 0
    for im in imgs:
        o.extract(im)
        fl = [o.tie_flags(l) for l in range(params[2])]
        order += sum(1 for f in fl if f & 2); cut += sum(1 for f in fl if f & 1)
    print(f"== {name}: {st['images']} images")
    print(f"   images byte-identical           {st['images_identical']}/{st['images']}")
    print(f"   levels byte-identical           {st['levels_identical']}/{st['levels']}")
    print(f"   images with a different count   {st['images_count_differs']}/{st['images']}")
    print(f"   reference keypoints             {st['ref_kps']}  (oracle {st['oracle_kps']})")
    print(f"   shared (same octave,x,y)        {st['common']}  = {100.0 * st['common'] / st['ref_kps']:.2f} % of the reference's")
    print(f"   only in reference / only oracle {st['ref_only']} / {st['oracle_only']}  = {100.0 * st['ref_only'] / st['ref_kps']:.2f} % / {100.0 * st['oracle_only'] / st['oracle_kps']:.2f} %")
    print(f"   shared keypoints whose angle/response/size/descriptor differ: {st['common_mismatch']}")
    print(f"   oracle tie flags: order-sensitive levels {order}/{st['levels']}, cut at a node of an equal-size run (set-sensitive) {cut}/{st['levels']}")

# ---- FMA contraction (rule 3): the reference's own CMake flags (-O3 -march=native, GCC default -ffp-contract=fast) let the
# compiler fuse x*b + y*a (ORBextractor.cc:119-120) and uL - scale*(...) (Frame.cc:2284-2286); the oracle is strict IEEE.
if ref_path.available("mono_fma"):
    params = (2000, 1.2, 8, 20, 7)
    n = pairs
    kp = kp_bad = bits = rows = matched = matched_diff = ur_bits = 0
    max_du = 0.0
    for s in range(700, 700 + n):
        L, R = synth_pair(1242, 375, s)
        a = [ref_path.ReferenceExtractor(*params, variant="mono_fma", primitives="native") for _ in range(2)]
        b = [ref_path.ReferenceExtractor(*params, variant="mono", primitives="native") for _ in range(2)]
        ka, da = zip(*[e.extract(im) for e, im in zip(a, (L, R))]); kb, db = zip(*[e.extract(im) for e, im in zip(b, (L, R))])
        for x, y, dx, dy in zip(ka, kb, da, db):
            kp += len(x); kp_bad += int(x.tobytes() != y.tobytes())
            d = np.unpackbits(dx ^ dy, axis=1).sum(axis=1); bits += int(d.sum()); rows += int((d > 0).sum())
        mb = float(np.float32(384.38148) / np.float32(721.5377))
        ua, _ = ref_path.stereo_match(a[0], a[1], ka[0], da[0], ka[1], da[1], 384.38148, mb)
        ub, _ = ref_path.stereo_match(b[0], b[1], kb[0], db[0], kb[1], db[1], 384.38148, mb)
        matched += int((ub >= 0).sum()); matched_diff += int(((ua >= 0) != (ub >= 0)).sum())
        both = (ua >= 0) & (ub >= 0)
        ur_bits += int((ua[both].view(np.uint32) != ub[both].view(np.uint32)).sum())
        if both.any():
            max_du = max(max_du, float(np.abs(ua[both] - ub[both]).max()))
    print(f"== FMA-contracted build (-ffp-contract=fast) vs strict, kitti (2000,1.2,8,20,7), {n} pairs, both with the monotonic allocator")
    print(f"   keypoint arrays differing        {kp_bad}/{2 * n} images ({kp} keypoints)")
    print(f"   descriptor rows / bits differing {rows} / {bits}")
    print(f"   stereo: matched {matched}, membership differs {matched_diff}, uRight bit patterns differing {ur_bits}, max |d uRight| {max_du:.3g} px")
