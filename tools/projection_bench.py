#!/usr/bin/env python3
"""Timing of the two 'next' rows built on the feature grid (GPU box):  python tools/projection_bench.py
  pslot_assign_grid + pslot_search_by_projection through the C ABI (host arrays in, host arrays out, wall clock over
  repeated calls) next to the oracle port on one host core, for a KITTI frame with 2000 features and 4000 map points."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle
from pointslot_b200 import binding as B
from test_oracle_projection import make_case, run_oracle


def main():
    for n_feat, n_mp, th in ((2000, 4000, 1.0), (2000, 4000, 3.0), (1000, 1500, 1.0)):
        c = make_case(oracle, 11, n_feat=n_feat, n_mp=n_mp)
        e = B.OrbExtractor(n_feat, 1.2, 8, 20, 7)
        spec = B.GridSpec.for_image(c["W"], c["H"])
        for _ in range(5):
            B.assign_grid(e, c["k"], spec)
            got, n, sweeps = B.search_by_projection(e, c["mps"], c["md"], c["d"], c["ur"], c["occupied"], c["sf"], th, 0.8)
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps):
            B.assign_grid(e, c["k"], spec)
        t1 = time.perf_counter()
        for _ in range(reps):
            B.search_by_projection(e, c["mps"], c["md"], c["d"], c["ur"], c["occupied"], c["sf"], th, 0.8)
        t2 = time.perf_counter()
        r0 = time.perf_counter()
        for _ in range(10):
            ref, nref = run_oracle(oracle, c, th, 0.8)
        r1 = time.perf_counter()
        assert np.array_equal(got, ref)
        print(f"features {len(c['k'])}, map points {n_mp}, th {th}: matches {n}, sweeps {sweeps}; "
              f"assign_grid {1e3 * (t1 - t0) / reps:.3f} ms/call, search_by_projection {1e3 * (t2 - t1) / reps:.3f} ms/call "
              f"(C ABI, host arrays); oracle grid+search {1e3 * (r1 - r0) / 10:.3f} ms on one core")


if __name__ == "__main__":
    main()
