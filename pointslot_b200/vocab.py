"""The ORB vocabulary as flat arrays - what `ORBVocabulary::loadFromTextFile` (reference
Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1389-1474, called from src/System.cc) builds from `Vocabulary/ORBvoc.txt`:
first line `k L scoring weighting`, then one line per node `parent isLeaf d0 .. d31 weight`; node ids are line numbers
(root = 0), children keep file order, leaves get word ids in file order."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Vocabulary:
    k: int
    L: int
    scoring: int       # 0 L1_NORM, 1 L2_NORM, 2 CHI_SQUARE, 3 KL, 4 BHATTACHARYYA, 5 DOT_PRODUCT
    weighting: int     # 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY
    parent: np.ndarray   # [n] int32, parent[0] = -1
    desc: np.ndarray     # [n][32] uint8
    weight: np.ndarray   # [n] float64
    is_leaf: np.ndarray  # [n] bool

    def __post_init__(self):
        n = len(self.parent)
        order = np.argsort(self.parent[1:], kind="stable") + 1   # children grouped by parent, file order kept
        counts = np.bincount(self.parent[1:], minlength=n)
        self.child_start = np.zeros(n + 1, np.int32)
        np.cumsum(counts, out=self.child_start[1:])
        self.child = order.astype(np.int32)
        self.word_id = np.full(n, -1, np.int32)
        self.word_id[self.is_leaf] = np.arange(int(self.is_leaf.sum()), dtype=np.int32)
        assert np.array_equal(self.is_leaf[1:], counts[1:] == 0), "isLeaf flags must agree with the tree"

    @property
    def n_nodes(self) -> int:
        return len(self.parent)

    @staticmethod
    def from_text(path) -> "Vocabulary":
        with open(path) as f:
            k, L, scoring, weighting = (int(t) for t in f.readline().split())
            rows = [ln.split() for ln in f if ln.strip()]
        n = len(rows) + 1
        parent = np.full(n, -1, np.int32); desc = np.zeros((n, 32), np.uint8)
        weight = np.zeros(n, np.float64); leaf = np.zeros(n, bool)
        for i, r in enumerate(rows, 1):
            parent[i] = int(r[0]); leaf[i] = int(r[1]) > 0
            desc[i] = [int(t) for t in r[2:34]]
            weight[i] = float(r[34])
        return Vocabulary(k, L, scoring, weighting, parent, desc, weight, leaf)

    def to_text(self, path) -> None:
        """No newline after the last node: the reference's `while(!f.eof())` loop would turn a trailing empty line into an
        extra node with an uninitialised parent."""
        lines = [f"{self.k} {self.L} {self.scoring} {self.weighting}"]
        for i in range(1, self.n_nodes):
            lines.append(f"{self.parent[i]} {int(self.is_leaf[i])} " + " ".join(str(int(b)) for b in self.desc[i]) + f" {float(self.weight[i])!r}")
        with open(path, "w") as f:
            f.write("\n".join(lines))


def random_vocabulary(k=10, L=4, seed=0, scoring=0, weighting=0, early_leaf=0.0, stop_fraction=0.02, flip_bits=40) -> Vocabulary:
    """A synthetic vocabulary tree for the tests and benches (the real ORBvoc.txt, 145 MB, is not in the reference tree):
    depth-first node order as DBoW2 writes it, children = the parent's descriptor with `flip_bits` random bits flipped
    (so that the walk is a real nearest-centre descent), `early_leaf` = probability that an inner node above level L ends
    as a leaf, `stop_fraction` of the words stopped (weight 0)."""
    rng = np.random.default_rng(seed)
    parent, desc, weight, leaf = [-1], [np.zeros(32, np.uint8)], [0.0], [False]

    def grow(pid, level):
        for _ in range(int(rng.integers(max(2, k - 3), k + 1)) if early_leaf else k):
            d = desc[pid].copy() if pid else rng.integers(0, 256, 32).astype(np.uint8)
            if pid:
                bits = rng.choice(256, max(4, flip_bits // level), replace=False)
                for b in bits:
                    d[b >> 3] ^= np.uint8(1 << (b & 7))
            nid = len(parent)
            is_leaf = level == L or (level > 1 and rng.random() < early_leaf)
            parent.append(pid); desc.append(d); leaf.append(is_leaf)
            weight.append(0.0 if not is_leaf else (0.0 if rng.random() < stop_fraction else float(rng.uniform(0.5, 9.0))))
            if not is_leaf:
                grow(nid, level + 1)

    grow(0, 1)
    return Vocabulary(k, L, scoring, weighting, np.array(parent, np.int32), np.stack(desc), np.array(weight), np.array(leaf))
