"""GPU: the bench workload at BASELINE.json's full size (configs[1]: 1000 KITTI-shaped pairs, ORBextractor(1000, 1.2,
8, 20, 5)) through the batch C ABI, checked by size-independent properties — the oracle would need minutes for all of it:

  * batching invariance: every pair's result is independent of the sub-batch it travelled in, of its position in the
    sub-batch and of the sub-batch size (a digest per pair, two different partitions of the sequence);
  * structure: keypoints level-ascending, counts within capacity, unmatched entries exactly -1, depth = bf / disparity;
  * the oracle on a random sample of the pairs (bit-exact features, identical matched set, |uRight| within 1e-3 px).
"""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PARAMS = (1000, 1.2, 8, 20, 5)
W, H, N = 1242, 375, 1000
BF, FX = 384.38148, 721.5377


def _digests(B, Ls, Rs, order, bsz, maxD, keep=()):
    sb = B.StereoBatch(PARAMS, W, H, bsz)
    dig, kept = {}, {}
    for lo in range(0, len(order), bsz):
        ids = order[lo:lo + bsz]
        res = sb.run(np.ascontiguousarray(Ls[ids]), np.ascontiguousarray(Rs[ids]), BF, maxD)
        for i, r in zip(ids, res):
            h = hashlib.sha1()
            for k in ("kL", "dL", "kR", "dR", "u_right", "depth"):
                h.update(r[k].tobytes())
            dig[int(i)] = h.hexdigest()
            if int(i) in keep:
                kept[int(i)] = r
    return dig, kept


def test_thousand_pairs_batching_invariance_and_sampled_parity(oracle_mod):
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_batch_parallel
    maxD = B.max_disparity(BF, FX)
    Ls, Rs = synth_batch_parallel(W, H, range(5000, 5000 + N))
    rng = np.random.default_rng(1)
    sample = set(int(i) for i in rng.choice(N, 10, replace=False))

    d1, kept = _digests(B, Ls, Rs, np.arange(N), 50, maxD, keep=sample)
    d2, _ = _digests(B, Ls, Rs, rng.permutation(N), 37, maxD)  # other order, ragged sub-batches (1000 = 27 * 37 + 1)
    assert d1 == d2
    assert len(set(d1.values())) == N  # every pair is different data: no slot was served from a neighbour's result

    oL, oR = oracle_mod.OracleExtractor(*PARAMS), oracle_mod.OracleExtractor(*PARAMS)
    for i in sorted(sample):
        r = kept[i]
        okL, odL = oL.extract(Ls[i]); okR, odR = oR.extract(Rs[i])
        our, odp = oracle_mod.stereo_match(oL, oR, okL, odL, okR, odR, BF, maxD)
        assert r["kL"].tobytes() == okL.tobytes() and r["kR"].tobytes() == okR.tobytes()
        assert np.array_equal(r["dL"], odL) and np.array_equal(r["dR"], odR)
        assert np.array_equal(r["u_right"] >= 0, our >= 0)
        m = our >= 0
        assert np.all(np.abs(r["u_right"][m] - our[m]) <= 1e-3)
        assert np.all(r["u_right"][~m] == -1) and np.all(r["depth"][~m] == -1)
        # structure
        assert np.all(np.diff(r["kL"]["octave"]) >= 0) and np.all(np.diff(r["kR"]["octave"]) >= 0)
        assert 0 < len(r["kL"]) <= 1100
        disp = r["kL"]["x"][m] - r["u_right"][m]
        assert np.all(disp > 0) and np.all(disp < maxD)
        big = disp > 0.5  # (disparities <= 0 are clamped to 0.01 by the reference: Frame.cc:2290-2294)
        assert np.allclose(r["depth"][m][big], np.float32(BF) / disp[big], rtol=1e-5)
