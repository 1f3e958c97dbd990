#!/usr/bin/env python3
"""GPU box: Frame::ComputeBoW (reference Frame.cc:2040-2047) through the C ABI (pslot_bow_transform + pslot_bow_vectors,
host descriptors in, containers out) next to the reference's own DBoW2 (oracle/_ref/libref_dbow2.so, one host thread, when
present) and the C oracle, on a synthetic k = 10, L = 5 vocabulary (111 111 nodes).  usage: python tools/bow_bench.py"""
import ctypes as C, json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import oracle as O
from pointslot_b200 import binding as B
from pointslot_b200.vocab import random_vocabulary

voc = random_vocabulary(10, 5, seed=1)
rng = np.random.default_rng(0)
leaves = np.nonzero(voc.is_leaf)[0]
out = {"vocabulary": {"k": 10, "L": 5, "nodes": voc.n_nodes, "words": int(voc.is_leaf.sum())}}
dv = B.OrbVocabulary(voc)
ref = ROOT / "oracle" / "_ref" / "libref_dbow2.so"
h = lib = None
if ref.exists():
    voc.to_text("/tmp/bow_bench_voc.txt")
    lib = C.CDLL(str(ref)); lib.refbow_load.restype = C.c_void_p; lib.refbow_load.argtypes = [C.c_char_p]
    vp = C.c_void_p
    lib.refbow_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.POINTER(C.c_int)]
    h = lib.refbow_load(b"/tmp/bow_bench_voc.txt")
for n in (1000, 2000, 8000):
    d = voc.desc[rng.choice(leaves, n)].copy()
    for r in d:
        for b in rng.choice(256, 10, replace=False):
            r[b >> 3] ^= np.uint8(1 << (b & 7))
    for _ in range(5):
        dv.compute_bow(d, 4)
    t0 = time.perf_counter()
    for _ in range(100):
        leaf, nid = dv.transform(d, 4)
    t_walk = (time.perf_counter() - t0) / 100 * 1e3
    t0 = time.perf_counter()
    for _ in range(100):
        dv.vectors(leaf, nid)
    t_vec = (time.perf_counter() - t0) / 100 * 1e3
    r = {"device_walk_ms": t_walk, "host_containers_ms": t_vec}
    t0 = time.perf_counter()
    for _ in range(10):
        O.bow_vectors(voc, *O.bow_transform(voc, d, 4))
    r["oracle_c_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    if h:
        bid = np.zeros(n, np.int32); bval = np.zeros(n, np.float64); fn = np.zeros(n, np.int32); fs = np.zeros(n + 1, np.int32); ff = np.zeros(n, np.int32)
        nf = C.c_int(0)
        t0 = time.perf_counter()
        for _ in range(10):
            lib.refbow_transform(h, d.ctypes.data, n, 4, bid.ctypes.data, bval.ctypes.data, fn.ctypes.data, fs.ctypes.data, ff.ctypes.data, C.byref(nf))
        r["reference_dbow2_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    out[f"{n} features"] = r
print(json.dumps(out, indent=1))
