"""Synthetic stereo pairs: layered smooth background + textured rectangles at
integer disparities drawn in painter's order in both views, mild blur + sensor noise.

Pure input generation (numpy + cv2 on the host); shared by the oracle, the CPU baseline and the
CUDA path so that all three see identical bytes.  Deterministic in (W, H, seed).
"""
from __future__ import annotations

import numpy as np


def synth_pair(W: int, H: int, seed: int):
    """-> (left, right) uint8 [H, W] images."""
    import cv2

    rng = np.random.default_rng(seed)
    Wc = W + 104
    base = np.full((H, Wc), 60.0, np.float32)
    for cell, amp in ((64, 50.0), (16, 30.0), (4, 18.0)):
        g = rng.random((H // cell + 2, Wc // cell + 2), dtype=np.float32)
        base += amp * cv2.resize(g, (Wc, H), interpolation=cv2.INTER_CUBIC)
    nrect = int(W * H / 1200)
    ds = np.sort(np.rint(2 + 94 * rng.random(nrect) ** 2.5)).astype(np.int64)  # far -> near
    L = base.copy()
    R = np.empty_like(base)
    R[:, :-2] = base[:, 2:]
    R[:, -2:] = base[:, -2:]
    for d in ds:
        w = int(rng.integers(8, 140)); h = int(rng.integers(8, 90))
        w = min(w, Wc - 97); h = min(h, H - 1)
        x = int(rng.integers(96, Wc - w)); y = int(rng.integers(0, H - h))
        val = float(rng.uniform(10, 245))
        tex = 3.0 * cv2.GaussianBlur(rng.normal(0, 12, (h, w)).astype(np.float32), (0, 0), 1.2)
        patch = val + tex
        L[y:y + h, x:x + w] = patch
        R[y:y + h, x - d:x - d + w] = patch
    out = []
    for k, im in enumerate((L, R)):
        im = cv2.GaussianBlur(im, (0, 0), 0.7)
        im = im + np.random.default_rng(2 * seed + 1 + k).normal(0, 1.0, im.shape).astype(np.float32)
        out.append(np.clip(np.rint(im), 0, 255).astype(np.uint8)[:, :W].copy())
    return out[0], out[1]


def synth_batch(W: int, H: int, seeds):
    """-> (left [B,H,W], right [B,H,W]) uint8"""
    seeds = list(seeds)
    L = np.empty((len(seeds), H, W), np.uint8)
    R = np.empty((len(seeds), H, W), np.uint8)
    for i, s in enumerate(seeds):
        L[i], R[i] = synth_pair(W, H, int(s))
    return L, R


def _worker(args):
    W, H, s = args
    return synth_pair(W, H, s)


def synth_batch_parallel(W: int, H: int, seeds, procs: int = 0):
    """synth_batch over a process pool (generation is ~20 ms/pair on one core)."""
    import multiprocessing as mp
    import os

    seeds = [int(s) for s in seeds]
    procs = procs or min(len(seeds), os.cpu_count() or 1, 32)
    if procs <= 1 or len(seeds) < 4:
        return synth_batch(W, H, seeds)
    L = np.empty((len(seeds), H, W), np.uint8)
    R = np.empty((len(seeds), H, W), np.uint8)
    with mp.get_context("fork").Pool(procs) as pool:
        for i, (l, r) in enumerate(pool.imap(_worker, [(W, H, s) for s in seeds], chunksize=4)):
            L[i], R[i] = l, r
    return L, R
