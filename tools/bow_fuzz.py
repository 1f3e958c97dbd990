"""Randomised parity sweep of the Frame::ComputeBoW oracle (orc_bow_transform / orc_bow_vectors) against the reference's own
vendored DBoW2 (oracle/_ref/libref_dbow2.so): random branching factors, depths, scoring and weighting types, levelsup; word ids,
word values bit for bit, node ids and feature lists on 1500 / 37 / 1 features per vocabulary (the comparison is the body of
tests/test_oracle_bow.py).  CPU only; test infrastructure.

    python tools/bow_fuzz.py [vocabularies=80] [seed=0] > profiles/r02_bow_fuzz.txt
"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle as O  # noqa: E402
import test_oracle_bow as TB  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if not TB.REF.exists():
        sys.exit("oracle/_ref/libref_dbow2.so is not built (make -C oracle ref REF=/root/reference)")
    O.build()
    rng = np.random.default_rng(seed)
    bad = nodes = 0
    t0 = time.time()
    for _ in range(n):
        k = int(rng.integers(2, 11)); L = int(rng.integers(1, 7))
        while k ** L > 60000:
            L -= 1
        sc = int(rng.integers(0, 6)); w = int(rng.integers(0, 4)); lu = int(rng.integers(0, L + 2))
        nodes += sum(k ** i for i in range(L + 1))
        with tempfile.TemporaryDirectory() as d:
            try:
                TB.test_bow_oracle_equals_reference_dbow2(Path(d), O, k, L, sc, w, lu)
            except AssertionError as e:
                bad += 1
                print(f"MISMATCH k={k} L={L} scoring={sc} weighting={w} levelsup={lu}: {e}", flush=True)
    print(f"BoW oracle vs the reference's DBoW2, randomised (seed {seed}): {n} vocabularies ({nodes} tree nodes in all), 3 feature sets each "
          f"(1500 / 37 / 1 features), {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
