"""Randomised parity sweep of the oracle against the reference's own code (oracle/_ref/libref_orb_mono.so: src/ORBextractor.cc
unmodified + Frame::ComputeStereoMatches lifted verbatim, monotonic allocator - DESIGN.md §2).  CPU only; test infrastructure.

Every case draws an image shape, extractor parameters (nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST), stereo
calibration and an image content class, runs both implementations on the left and the right view and compares pyramid levels,
keypoints, descriptors, mvuRight and mvDepth byte for byte.  Shapes the reference cannot process (a level with less than one
30-px cell per axis, or zero quadtree roots: it divides by zero, ORBextractor.cc:784-787 / :543-545) are the ones the oracle
rejects; they are counted, and the reference is not called on them.

    python tools/ref_fuzz.py [cases=200] [seed=0] [--cv2] > profiles/r02_ref_fuzz.txt
(--cv2: OpenCV 4.13 itself, through cv2, serves the five primitives to the reference's glue instead of the oracle's restatements)
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import oracle as O  # noqa: E402
from oracle import ref_path as RP  # noqa: E402
from pointslot_b200.synth import synth_pair  # noqa: E402

F = np.float32
VERBOSE = "-v" in sys.argv
if VERBOSE:
    sys.argv.remove("-v")
# who serves resize / copyMakeBorder / FAST / GaussianBlur / fastAtan2 to the reference's glue: the oracle's restatements
# ("native", fast) or OpenCV 4.13 itself through cv2 ("cv2": the complete check - glue AND primitives against the library)
PRIMITIVES = "cv2" if "--cv2" in sys.argv else "native"
if "--cv2" in sys.argv:
    sys.argv.remove("--cv2")
CONTENT = ["synth", "noise", "smooth_noise", "checker", "posterised", "low_contrast", "saturated", "flat", "shifted_noise"]


def make_pair(kind, W, H, rng, seed):
    """-> (left, right) u8 images of one content class; the right view is a horizontally displaced left view where the class
    has no stereo geometry of its own, so the stereo matcher has something to find."""
    if kind == "synth":
        return synth_pair(W, H, seed)
    if kind == "noise":
        L = rng.integers(0, 256, (H, W + 40), dtype=np.uint8)
    elif kind == "smooth_noise":
        import cv2
        L = cv2.GaussianBlur(rng.integers(0, 256, (H, W + 40)).astype(np.float32), (0, 0), float(rng.uniform(0.8, 3.0)))
        L = np.clip(np.rint((L - L.mean()) * 4 + 128), 0, 255).astype(np.uint8)
    elif kind == "checker":
        c = int(rng.integers(5, 40))
        yy, xx = np.mgrid[0:H, 0:W + 40]
        L = ((((yy // c) + (xx // c)) & 1) * int(rng.integers(40, 255))).astype(np.uint8)
        L = (L.astype(np.int16) + rng.integers(-3, 4, L.shape)).clip(0, 255).astype(np.uint8)
    elif kind == "posterised":
        a, _ = synth_pair(W + 40, H, seed)
        q = int(rng.choice([16, 32, 64]))
        L = (a // q * q).astype(np.uint8)
    elif kind == "low_contrast":
        a, _ = synth_pair(W + 40, H, seed)
        L = (100 + a.astype(np.int32) // 12).astype(np.uint8)
    elif kind == "saturated":
        a, _ = synth_pair(W + 40, H, seed)
        L = np.clip(a.astype(np.int32) * 3 - 200, 0, 255).astype(np.uint8)
    elif kind == "flat":
        L = np.full((H, W + 40), int(rng.integers(0, 256)), np.uint8)
        L[H // 3: H // 3 + 9, W // 4: W // 4 + 9] ^= 0x80  # one feature-bearing blob, the rest falls to minThFAST and finds nothing
    elif kind == "shifted_noise":
        L = rng.integers(0, 256, (H, W + 40), dtype=np.uint8)
        L[::2] = L[::2] // 2
    else:
        raise ValueError(kind)
    d = int(rng.integers(1, 40))
    return np.ascontiguousarray(L[:, 40:40 + W]), np.ascontiguousarray(L[:, 40 - d:40 - d + W])


def draw_case(rng):
    W = int(rng.integers(96, 1400))
    H = int(rng.integers(96, 800))
    if rng.random() < 0.15:
        W, H = H, W
    nlevels = int(rng.integers(1, 10))
    scale = float(F(rng.choice([1.1, 1.15, 1.2, 1.25, 1.3, 1.4, round(float(rng.uniform(1.05, 1.45)), 3)])))
    nfeat = int(rng.choice([50, 200, 500, 1000, 1200, 2000, 3000, int(rng.integers(20, 5000))]))
    ini = int(rng.integers(6, 60))
    mn = int(rng.integers(1, ini + 1))
    fx = float(rng.uniform(300, 1000))
    bf = float(rng.uniform(20, 500))
    return W, H, (nfeat, scale, nlevels, ini, mn), bf, fx


def reference_undefined(W, H, params):
    """True where the reference's own arithmetic is undefined: some level has less than one 30-px cell per axis
    (nCols / nRows == 0 -> ceil(width / 0), ORBextractor.cc:784-787) or zero quadtree roots (nIni == 0 -> hX = width / 0 and
    vpIniNodes[...] on an empty vector, :543-563).  Restated here independently of the oracle, from the constructor's tables
    (:418-430) and ComputePyramid's level sizes (:1112)."""
    _, scale, nlevels, _, _ = params
    sf = F(1.0)
    for l in range(nlevels):
        if l:
            sf = F(float(sf) * float(F(scale)))  # float * double -> float (:421; scaleFactor is a double member)
        inv = F(1.0) / sf
        w = int(np.rint(F(W) * inv)); h = int(np.rint(F(H) * inv))  # cvRound((float)cols * mvInvScaleFactor[l])
        ww, hh = w - 32, h - 32  # the detection window [16, dim - 16)
        if ww // 30 < 1 or hh // 30 < 1:
            return True
        if int(np.floor(F(ww) / F(hh) + F(0.5))) < 1:  # round(): half away from zero, positive here
            return True
    return False


def run(cases, seed0, primitives=None):
    primitives = primitives or PRIMITIVES
    rng = np.random.default_rng(seed0)
    stats = dict(cases=0, rejected=0, images=0, keypoints=0, stereo_pairs=0, matches=0, mismatches=0, by_content={})
    t0 = time.time()
    n = 0
    while stats["cases"] < cases:
        n += 1
        W, H, params, bf, fx = draw_case(rng)
        kind = CONTENT[n % len(CONTENT)]
        undefined = reference_undefined(W, H, params)
        if VERBOSE:
            print(f"draw {n}: {W}x{H} {params} {kind} undefined={undefined}", file=sys.stderr, flush=True)
        oL, oR = O.OracleExtractor(*params), O.OracleExtractor(*params)
        L, R = make_pair(kind, W, H, rng, seed0 * 100003 + n)
        try:
            okL, odL = oL.extract(L)
            okR, odR = oR.extract(R)
        except RuntimeError:
            if not undefined:  # the oracle may only refuse what the reference cannot compute (parity rule 6)
                stats["mismatches"] += 1
                print(f"OVER-REJECTION draw {n}: {W}x{H} {params}", flush=True)
            stats["rejected"] += 1
            continue
        if undefined:
            stats["mismatches"] += 1
            print(f"UNDER-REJECTION draw {n}: {W}x{H} {params}: the oracle accepted a shape the reference is undefined on", flush=True)
            continue
        rL = RP.ReferenceExtractor(*params, variant="mono", primitives=primitives)
        rR = RP.ReferenceExtractor(*params, variant="mono", primitives=primitives)
        kL, dL = rL.extract(L)
        kR, dR = rR.extract(R)
        bad = []
        for l in range(params[2]):
            if not (np.array_equal(rL.level(l), oL.level(l)) and np.array_equal(rR.level(l), oR.level(l))):
                bad.append(f"pyramid level {l}")
        if kL.tobytes() != okL.tobytes() or kR.tobytes() != okR.tobytes():
            bad.append(f"keypoints ({len(kL)}/{len(okL)} left, {len(kR)}/{len(okR)} right)")
        elif not (np.array_equal(dL, odL) and np.array_equal(dR, odR)):
            bad.append("descriptors")
        nm = 0
        if not bad and len(kL) and len(kR):
            mb = float(F(bf) / F(fx))
            ur, dp = RP.stereo_match(rL, rR, kL, dL, kR, dR, bf, mb)
            our, odp = O.stereo_match(oL, oR, okL, odL, okR, odR, bf, float(F(bf) / F(mb)))
            if ur.tobytes() != our.tobytes():
                bad.append("mvuRight")
            if dp.tobytes() != odp.tobytes():
                bad.append("mvDepth")
            nm = int((ur >= 0).sum())
            stats["stereo_pairs"] += 1
        stats["cases"] += 1
        stats["images"] += 2
        stats["keypoints"] += len(kL) + len(kR)
        stats["matches"] += nm
        c = stats["by_content"].setdefault(kind, [0, 0, 0])
        c[0] += 1; c[1] += len(kL) + len(kR); c[2] += nm
        if bad:
            stats["mismatches"] += 1
            print(f"MISMATCH case {n}: {W}x{H} {params} bf={bf:.3f} fx={fx:.3f} {kind}: {', '.join(bad)}", flush=True)
    stats["seconds"] = time.time() - t0
    return stats


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if not RP.available("mono"):
        sys.exit("oracle/_ref/libref_orb_mono.so is not built (make -C oracle ref REF=/root/reference)")
    O.build()
    # The mono build never reuses heap memory (that is its point), so a long sweep runs as chunks of CHUNK cases in fresh
    # processes (chunk i uses seed seed0 + i), a few of them side by side.
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    chunks = [(min(CHUNK, cases - c0), seed0 + i, PRIMITIVES) for i, c0 in enumerate(range(0, cases, CHUNK))]
    t0 = time.time()
    if len(chunks) == 1:
        parts = [run(*chunks[0])]
    else:
        with ProcessPoolExecutor(max_workers=min(len(chunks), max(1, (mp.cpu_count() or 2) // 2)), mp_context=mp.get_context("spawn"),
                                 max_tasks_per_child=1) as ex:
            parts = list(ex.map(_run_chunk, chunks))
    s = dict(cases=0, rejected=0, images=0, keypoints=0, stereo_pairs=0, matches=0, mismatches=0, by_content={})
    for part in parts:
        for k in ("cases", "rejected", "images", "keypoints", "stereo_pairs", "matches", "mismatches"):
            s[k] += part[k]
        for k, v in part["by_content"].items():
            c = s["by_content"].setdefault(k, [0, 0, 0])
            for i in range(3):
                c[i] += v[i]
    print(f"oracle vs reference binary (mono build, primitives served by {'OpenCV 4.13 / cv2' if PRIMITIVES == 'cv2' else 'the restatements'}), randomised: {s['cases']} cases ({s['images']} images, seeds {seed0}..{seed0 + len(chunks) - 1}), "
          f"{s['rejected']} draws rejected by both (the reference's arithmetic is undefined there), {s['keypoints']} keypoints, "
          f"{s['stereo_pairs']} stereo calls with {s['matches']} matches, {s['mismatches']} mismatching cases, {time.time() - t0:.0f} s")
    print("content class: cases / keypoints / stereo matches")
    for k, (a, b, c) in sorted(s["by_content"].items()):
        print(f"  {k:14s} {a:4d} {b:8d} {c:7d}")
    sys.exit(1 if s["mismatches"] else 0)


CHUNK = 100


def _run_chunk(args):
    return run(*args)


if __name__ == "__main__":
    main()
