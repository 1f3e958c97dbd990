"""Frame::ComputeBoW (reference src/Frame.cc:2040-2047) = ORBVocabulary::transform(descriptors, mBowVec, mFeatVec, 4): the
oracle's restatement of the DBoW2 tree walk and of the BowVector / FeatureVector arithmetic, pinned against the REFERENCE'S
OWN vendored DBoW2 (Thirdparty/DBoW2 compiled into oracle/_ref/libref_dbow2.so, loading the same vocabulary text file with
loadFromTextFile as System.cc does): word ids, values bit for bit (doubles), node ids, feature lists."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from pointslot_b200.vocab import Vocabulary, random_vocabulary

REF = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libref_dbow2.so"


def _features(voc, n, seed, noise=14):
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(voc.is_leaf)[0]
    d = voc.desc[rng.choice(leaves, n)].copy()
    for r in d:
        for b in rng.choice(256, int(rng.integers(0, noise + 1)), replace=False):
            r[b >> 3] ^= np.uint8(1 << (b & 7))
    d[::7] = rng.integers(0, 256, (len(d[::7]), 32), dtype=np.uint8)  # some features far from every centre
    return d


def _ref_transform(lib, h, desc, levelsup):
    n = len(desc)
    vp = C.c_void_p
    bid = np.zeros(n, np.int32); bval = np.zeros(n, np.float64)
    fn = np.zeros(n, np.int32); fs = np.zeros(n + 1, np.int32); ff = np.zeros(n, np.int32)
    nf = C.c_int(0)
    lib.refbow_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.POINTER(C.c_int)]
    nw = lib.refbow_transform(h, desc.ctypes.data, n, levelsup, bid.ctypes.data, bval.ctypes.data, fn.ctypes.data, fs.ctypes.data,
                              ff.ctypes.data, C.byref(nf))
    return bid[:nw], bval[:nw], fn[:nf.value], fs[:nf.value + 1], ff[:fs[nf.value]]


@pytest.mark.parametrize("k,L,scoring,weighting,levelsup", [(10, 4, 0, 0, 2), (10, 3, 0, 0, 4), (6, 5, 1, 0, 3), (8, 3, 5, 1, 1),
                                                            (9, 3, 2, 2, 2), (10, 3, 4, 3, 0), (4, 6, 3, 0, 4)])
def test_bow_oracle_equals_reference_dbow2(tmp_path, oracle_mod, k, L, scoring, weighting, levelsup):
    if not REF.exists():
        pytest.skip("oracle/_ref/libref_dbow2.so not built (make -C oracle ref; needs /root/reference)")
    voc = random_vocabulary(k, L, seed=k * 10 + L, scoring=scoring, weighting=weighting)
    path = tmp_path / "voc.txt"
    voc.to_text(path)
    again = Vocabulary.from_text(path)
    assert np.array_equal(again.desc, voc.desc) and np.array_equal(again.weight, voc.weight) and np.array_equal(again.child, voc.child)
    lib = C.CDLL(str(REF))
    lib.refbow_load.restype = C.c_void_p
    lib.refbow_load.argtypes = [C.c_char_p]
    lib.refbow_size.argtypes = [C.c_void_p]
    lib.refbow_free.argtypes = [C.c_void_p]
    h = lib.refbow_load(str(path).encode())
    assert h and lib.refbow_size(h) == int(voc.is_leaf.sum())
    for n, seed in ((1500, 1), (37, 2), (1, 3)):
        desc = _features(voc, n, seed)
        leaf, nid = oracle_mod.bow_transform(voc, desc, levelsup)
        got = oracle_mod.bow_vectors(voc, leaf, nid)
        want = _ref_transform(lib, h, desc, levelsup)
        assert np.array_equal(got[0], want[0]), "word ids"
        assert np.array_equal(got[1].view(np.uint64), want[1].view(np.uint64)), "word values (bitwise)"
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3]) and np.array_equal(got[4], want[4]), "FeatureVector"
        assert len(got[0]) > 0
    lib.refbow_free(h)


def test_bow_oracle_properties(oracle_mod):
    """Sizes the reference cannot load quickly and trees with early leaves (where the reference leaves *nid unset, so they
    are not part of the pin): every feature lands on a leaf below the root, nid is an ancestor-or-self of the leaf at
    the requested level, values are L1-normalised, feature lists partition the non-stopped features."""
    voc = random_vocabulary(10, 4, seed=5, early_leaf=0.15)
    desc = _features(voc, 2000, 9)
    for levelsup in (0, 2, 4, 9):
        leaf, nid = oracle_mod.bow_transform(voc, desc, levelsup)
        assert np.all(voc.is_leaf[leaf])
        depth = np.zeros(voc.n_nodes, np.int32)
        for i in range(1, voc.n_nodes):
            depth[i] = depth[voc.parent[i]] + 1
        for f in range(0, 2000, 97):
            a = leaf[f]
            while depth[a] > max(voc.L - levelsup, 0):
                a = voc.parent[a]
            assert nid[f] == a
        bid, bval, fn, fs, ff = oracle_mod.bow_vectors(voc, leaf, nid)
        assert np.all(np.diff(bid) > 0) and np.all(np.diff(fn) > 0)
        assert abs(bval.sum() - 1.0) < 1e-12
        kept = np.nonzero(voc.weight[leaf] > 0)[0]
        assert np.array_equal(np.sort(ff), kept)
        for j in range(len(fn)):
            seg = ff[fs[j]:fs[j + 1]]
            assert np.all(np.diff(seg) > 0) and np.all(nid[seg] == fn[j])
    k, v, *_ = oracle_mod.bow_vectors(voc, np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert len(k) == 0
