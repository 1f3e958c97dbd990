#!/usr/bin/env python3
"""GPU box, under ncu: a few single-frame extractions + stereo matches (the per-frame path) for the per-kernel launch list."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointslot_b200 import binding as B
from pointslot_b200.synth import synth_pair

params = (2000, 1.2, 8, 20, 7)
L, R = synth_pair(1242, 375, 0)
eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
for _ in range(3):
    kL, dL = eL.extract(L); kR, dR = eR.extract(R)
    B.stereo_match_resident(eL, eR, B.extract_token(eL), B.extract_token(eR), 384.38, B.max_disparity(384.38, 721.54))
print(len(kL), len(kR))
