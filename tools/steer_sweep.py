#!/usr/bin/env python3
"""Exhaustive check of the descriptor steering coefficients: for EVERY float32 angle in [0, 360] degrees
(1,135,869,953 bit patterns) the device's a = (float)cos(angle*pi/180), b = (float)sin(...) (CUDA double
sincos, rounded to float) must equal the oracle's (glibc double cos/sin, rounded to float) bit for bit.
Run on the GPU box:  python tools/steer_sweep.py [stride]   -> gpurun_out/steer_sweep.txt
"""
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402
from pointslot_b200 import binding as B  # noqa: E402


def oracle_parallel(ang, threads=16):
    a = np.empty_like(ang); b = np.empty_like(ang)
    parts = np.array_split(np.arange(ang.size), threads)

    def work(idx):
        if idx.size:
            aa, bb = O.steer_coeffs(ang[idx[0]:idx[-1] + 1])
            a[idx[0]:idx[-1] + 1] = aa; b[idx[0]:idx[-1] + 1] = bb

    ths = [threading.Thread(target=work, args=(p,)) for p in parts]
    for t in ths: t.start()
    for t in ths: t.join()
    return a, b


def main():
    stride = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    O.build()
    hi = int(np.float32(360.0).view(np.uint32)) + 1
    chunk = 1 << 24
    bad = total = 0
    first = []
    t0 = time.time()
    for lo in range(0, hi, chunk * stride):
        bits = np.arange(lo, min(lo + chunk * stride, hi), stride, dtype=np.uint32)
        ang = bits.view(np.float32)
        ga, gb = B.debug_steer(ang)
        oa, ob = oracle_parallel(ang)
        m = (ga.view(np.uint32) != oa.view(np.uint32)) | (gb.view(np.uint32) != ob.view(np.uint32))
        bad += int(m.sum()); total += ang.size
        for i in np.flatnonzero(m)[:3]:
            first.append((float(ang[i]), float(ga[i]), float(oa[i]), float(gb[i]), float(ob[i])))
    msg = (f"steer sweep: {total} float32 angles in [0,360] (stride {stride}), {bad} mismatches vs libm, "
           f"{time.time() - t0:.1f} s; first mismatches: {first[:6]}")
    print(msg)
    out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
    (out / "steer_sweep.txt").write_text(msg + "\n")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
