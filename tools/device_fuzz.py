"""GPU: randomised device-vs-oracle sweep of the hot path through the C ABI (single-frame form: two pslot_orb_extract calls +
pslot_stereo_match), on the cases tools/ref_fuzz.py draws (random image shapes, extractor parameters, calibrations, nine
content classes).  Keypoints and descriptors must be byte-identical to the oracle's, the stereo matches the same set with
|uRight - oracle| <= 1e-3 (the bar of tests/test_gpu_parity.py); a shape one side refuses must be refused by the other.
Stops drawing after `seconds` of wall clock.

    python tools/device_fuzz.py [seconds=8] [seed=0] > gpurun_out/device_fuzz.txt
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    t_start = time.time()
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    import oracle as O
    from pointslot_b200 import binding as B
    import ref_fuzz as RF

    assert B.device_count() > 0, "needs a CUDA device"
    rng = np.random.default_rng(seed0)
    F = np.float32
    cases = rejected = kp = matches = bad = bit_equal_stereo = stereo_calls = n = 0
    while time.time() - t_start < seconds:
        n += 1
        W, H, params, bf, fx = RF.draw_case(rng)
        W, H = min(W, 1000), min(H, 520)  # keep the CPU oracle's share of the time small
        kind = RF.CONTENT[n % len(RF.CONTENT)]
        L, R = RF.make_pair(kind, W, H, rng, seed0 * 100003 + n)
        what = f"draw {n}: {W}x{H} {params} {kind}"
        try:
            oL, oR = O.OracleExtractor(*params), O.OracleExtractor(*params)
            okL, odL = oL.extract(L)
            okR, odR = oR.extract(R)
            oracle_ok = True
        except (RuntimeError, ValueError):
            oracle_ok = False
        try:
            eL, eR = B.OrbExtractor(*params), B.OrbExtractor(*params)
            kL, dL = eL.extract(L)
            kR, dR = eR.extract(R)
            device_ok = True
        except B.PslotError as e:
            device_ok = False
            dev_err = str(e)
        if not oracle_ok or not device_ok:
            if oracle_ok != device_ok:
                if oracle_ok and any(t in dev_err for t in ("scaleFactor", "exceeds", "staging tile", "too many", "need room", "on-chip quadtree tables")):
                    rejected += 1  # a documented device limit (scale factor cap, image size, list capacity), not a parity matter
                else:
                    bad += 1
                    print(f"MISMATCH {what}: oracle {'ok' if oracle_ok else 'refused'}, device {'ok' if device_ok else 'refused: ' + dev_err}",
                          flush=True)
            else:
                rejected += 1
            continue
        cases += 1
        kp += len(okL) + len(okR)
        if kL.tobytes() != okL.tobytes() or kR.tobytes() != okR.tobytes():
            bad += 1
            print(f"MISMATCH {what}: keypoints ({len(kL)}/{len(okL)} left, {len(kR)}/{len(okR)} right)", flush=True)
            continue
        if not (np.array_equal(dL, odL) and np.array_equal(dR, odR)):
            bad += 1
            print(f"MISMATCH {what}: descriptors", flush=True)
            continue
        if len(kL) and len(kR):
            max_d = float(F(bf) / (F(bf) / F(fx)))
            try:
                ur, dp = B.stereo_match(eL, eR, kL, dL, kR, dR, bf, max_d)
            except B.PslotError as e:
                if "too many" in str(e):  # > 16000 right keypoints: documented capacity of the matcher
                    continue
                bad += 1
                print(f"MISMATCH {what}: stereo raised {e}", flush=True)
                continue
            our, odp = O.stereo_match(oL, oR, okL, odL, okR, odR, bf, max_d)
            stereo_calls += 1
            matches += int((our >= 0).sum())
            if not np.array_equal(ur >= 0, our >= 0) or not np.all(np.abs(ur - our) <= 1e-3):
                bad += 1
                print(f"MISMATCH {what}: stereo ({int((ur >= 0).sum())} vs {int((our >= 0).sum())} matches)", flush=True)
            elif ur.tobytes() == our.tobytes() and dp.tobytes() == odp.tobytes():
                bit_equal_stereo += 1
    print(f"device vs oracle through the C ABI, randomised (seed {seed0}, {time.time() - t_start:.1f} s): {cases} cases, {rejected} draws "
          f"refused by both sides or by a documented device limit, {kp} keypoints byte-identical, {stereo_calls} stereo calls with "
          f"{matches} matches ({bit_equal_stereo} calls bit-equal in uRight and depth), {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
