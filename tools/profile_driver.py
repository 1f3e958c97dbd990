#!/usr/bin/env python3
"""Minimal workload for ncu: N submits of one sub-batch of the bench workload (no torch, no CPU arm).

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_driver.py --batch 125 --submits 3
  ncu --set full --clock-control none --import-source on -k regex:k_fast -s 1 -c 1 -o gpurun_out/prof_fast \
      python tools/profile_driver.py --batch 125 --submits 2
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from pointslot_b200 import binding as B  # noqa: E402
from pointslot_b200.synth import synth_batch_parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=125)
    ap.add_argument("--submits", type=int, default=3)
    ap.add_argument("--distinct", type=int, default=25)
    ap.add_argument("--width", type=int, default=1242)
    ap.add_argument("--height", type=int, default=375)
    ap.add_argument("--params", default="1000,1.2,8,20,5")
    a = ap.parse_args()
    p = a.params.split(",")
    params = (int(p[0]), float(p[1]), int(p[2]), int(p[3]), int(p[4]))
    Ls, Rs = synth_batch_parallel(a.width, a.height, range(a.distinct))
    reps = (a.batch + a.distinct - 1) // a.distinct
    Ls = np.ascontiguousarray(np.tile(Ls, (reps, 1, 1))[: a.batch])
    Rs = np.ascontiguousarray(np.tile(Rs, (reps, 1, 1))[: a.batch])
    sb = B.StereoBatch(params, a.width, a.height, a.batch)
    out = sb.alloc_out(a.batch)
    bf, fx = 384.38148, 721.5377
    for _ in range(a.submits):
        sb.submit(Ls, Rs, bf, B.max_disparity(bf, fx), out)
        sb.wait()
    print("stats", sb.stats(), "keypoints image0", int(out["counts"][0]))


if __name__ == "__main__":
    main()
