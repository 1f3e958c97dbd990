"""Oracle vs the reference's own ORBextractor.cc compiled here (oracle/ref_compat/README.md).  Skipped unless
oracle/_ref/libref_orb.so has been built by the opt-in recipe `make -C oracle ref` — which round 1 did not run."""
import numpy as np
import pytest


def _ref():
    from oracle import ref_path
    if not ref_path.available():
        pytest.skip("oracle/_ref/libref_orb.so not built (opt-in: make -C oracle ref REF=/root/reference)")
    return ref_path


@pytest.mark.parametrize("shape,params,seed", [((375, 1242), (2000, 1.2, 8, 20, 7), 0), ((375, 1242), (1000, 1.2, 8, 20, 5), 7),
                                               ((480, 752), (1200, 1.2, 8, 20, 7), 3), ((301, 403), (500, 1.3, 6, 15, 5), 5)])
def test_oracle_equals_reference_extractor(shape, params, seed, oracle_mod):
    ref_path = _ref()
    from pointslot_b200.synth import synth_pair
    H, W = shape
    L, _ = synth_pair(W, H, seed)
    r = ref_path.ReferenceExtractor(*params); o = oracle_mod.OracleExtractor(*params)
    rk, rd = r.extract(L); ok, od = o.extract(L)
    for l in range(params[2]):
        assert np.array_equal(r.level(l), o.level(l)), f"pyramid level {l}"
    assert rk.tobytes() == ok.tobytes()
    assert np.array_equal(rd, od)
