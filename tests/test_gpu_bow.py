"""GPU: pslot_bow_transform / pslot_bow_vectors (bow.cu) = Frame::ComputeBoW (reference src/Frame.cc:2040-2047:
ORBVocabulary::transform(descriptors, mBowVec, mFeatVec, 4)) against the oracle, which tests/test_oracle_bow.py pins bit
for bit to the reference's own DBoW2; and directly against that DBoW2 build when oracle/_ref is present."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REF = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libref_dbow2.so"


@pytest.fixture(scope="module")
def B():
    from pointslot_b200 import binding
    return binding


def _features(voc, n, seed, noise=14):
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(voc.is_leaf)[0]
    d = voc.desc[rng.choice(leaves, n)].copy()
    for r in d:
        for b in rng.choice(256, int(rng.integers(0, noise + 1)), replace=False):
            r[b >> 3] ^= np.uint8(1 << (b & 7))
    d[::7] = rng.integers(0, 256, (len(d[::7]), 32), dtype=np.uint8)
    return d


@pytest.mark.parametrize("k,L,scoring,weighting,levelsup,early", [(10, 4, 0, 0, 4, 0.0), (10, 4, 0, 0, 2, 0.0), (7, 5, 1, 0, 3, 0.1),
                                                                  (9, 3, 5, 1, 1, 0.0), (10, 3, 2, 2, 0, 0.2), (5, 6, 4, 3, 4, 0.0)])
def test_bow_equals_oracle(B, oracle_mod, k, L, scoring, weighting, levelsup, early):
    from pointslot_b200.vocab import random_vocabulary
    voc = random_vocabulary(k, L, seed=3 * k + L, scoring=scoring, weighting=weighting, early_leaf=early)
    dv = B.OrbVocabulary(voc)
    for n, seed in ((2000, 1), (33, 2), (1, 3), (0, 4)):
        desc = _features(voc, n, seed) if n else np.zeros((0, 32), np.uint8)
        leaf, nid = dv.transform(desc, levelsup)
        oleaf, onid = oracle_mod.bow_transform(voc, desc, levelsup)
        assert np.array_equal(leaf, oleaf) and np.array_equal(nid, onid)
        got, want = dv.vectors(leaf, nid), oracle_mod.bow_vectors(voc, oleaf, onid)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint64), want[1].view(np.uint64))
        assert all(np.array_equal(a, b) for a, b in zip(got[2:], want[2:]))


def test_bow_of_extracted_frame_equals_reference_dbow2(B, oracle_mod, tmp_path):
    """The whole row as Tracking uses it: extract a frame, ComputeBoW on its descriptors - read on the device where the
    extractor left them - and compare mBowVec / mFeatVec with the reference's own DBoW2 loading the same vocabulary file."""
    from pointslot_b200.synth import synth_pair
    from pointslot_b200.vocab import random_vocabulary
    L, _ = synth_pair(752, 480, 5)
    ext = B.OrbExtractor(1200, 1.2, 8, 20, 7)
    kps, desc = ext.extract(L)
    rng = np.random.default_rng(0)
    voc = random_vocabulary(10, 4, seed=11)
    # centre the leaves on real descriptors so that the walk is not degenerate
    leaves = np.nonzero(voc.is_leaf)[0]
    voc.desc[leaves[:len(desc)]] = desc[rng.permutation(len(desc))][:len(leaves)]
    dv = B.OrbVocabulary(voc)
    from_dev = dv.compute_bow(None, 4, extractor=ext, n=len(desc))
    from_host = dv.compute_bow(desc, 4)
    assert all(np.array_equal(a, b) for a, b in zip(from_dev, from_host))
    want = oracle_mod.bow_vectors(voc, *oracle_mod.bow_transform(voc, desc, 4))
    assert np.array_equal(from_dev[1].view(np.uint64), want[1].view(np.uint64)) and all(np.array_equal(a, b) for a, b in zip(from_dev, want))
    if REF.exists():
        path = tmp_path / "voc.txt"
        voc.to_text(path)
        lib = C.CDLL(str(REF))
        lib.refbow_load.restype = C.c_void_p; lib.refbow_load.argtypes = [C.c_char_p]
        h = lib.refbow_load(str(path).encode())
        n = len(desc)
        bid = np.zeros(n, np.int32); bval = np.zeros(n, np.float64); fn = np.zeros(n, np.int32); fs = np.zeros(n + 1, np.int32); ff = np.zeros(n, np.int32)
        nf = C.c_int(0)
        vp = C.c_void_p
        lib.refbow_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.POINTER(C.c_int)]
        d = np.ascontiguousarray(desc)
        nw = lib.refbow_transform(h, d.ctypes.data, n, 4, bid.ctypes.data, bval.ctypes.data, fn.ctypes.data, fs.ctypes.data, ff.ctypes.data, C.byref(nf))
        assert np.array_equal(from_dev[0], bid[:nw]) and np.array_equal(from_dev[1].view(np.uint64), bval[:nw].view(np.uint64))
        assert np.array_equal(from_dev[2], fn[:nf.value]) and np.array_equal(from_dev[4], ff[:fs[nf.value]])
    assert len(from_dev[0]) > 300


def test_bow_argument_errors(B):
    from pointslot_b200.vocab import random_vocabulary
    voc = random_vocabulary(4, 2, seed=1)
    bad = random_vocabulary(4, 2, seed=1)
    bad.child = bad.child.copy(); bad.child[0] = 0   # a child that is not a later node
    with pytest.raises(B.PslotError):
        B.OrbVocabulary(bad)
    dv = B.OrbVocabulary(voc)
    with pytest.raises(B.PslotError):
        dv.vectors(np.array([0], np.int32), np.array([0], np.int32))   # the root is not a word
    with pytest.raises(B.PslotError):
        dv.transform(None, 4, extractor=None, n=5)
