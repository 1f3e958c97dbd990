#!/usr/bin/env python3
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair
L, _ = synth_pair(1242, 375, 3)
orb = B.CvOrb()
maps = []
for rep in range(3):
    orb.detect_and_compute(L)
    maps.append(orb.stage(1, 0))
    print(rep, int((maps[-1] > 0).sum()), int(maps[-1].max()))
orb2 = B.CvOrb(1000, 1.2, 1, 19)   # one level only
orb2.detect_and_compute(L)
m1 = orb2.stage(1, 0)
print("one level:", int((m1 > 0).sum()), int(m1.max()))
np.savez_compressed("gpurun_out/r2d/score.npz", a=maps[0], b=maps[1], c=maps[2], one=m1, img=orb.stage(0, 0))
