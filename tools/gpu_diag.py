#!/usr/bin/env python3
"""Stage-by-stage GPU-vs-oracle comparison that keeps going after a mismatch (pytest stops at the
first one).  Run on the GPU box:  python tools/gpu_diag.py [case ...]  -> gpurun_out/diag.txt
"""
from __future__ import annotations

import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle as O  # noqa: E402
from pointslot_b200 import binding as B  # noqa: E402
from pointslot_b200.synth import synth_pair  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
LOG = open(OUT / "diag.txt", "w")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def cmp_seq(name, got, exp):
    got = [np.asarray(g) for g in got]; exp = [np.asarray(e) for e in exp]
    n_g, n_e = len(got[0]), len(exp[0])
    if n_g != n_e:
        say(f"  {name}: COUNT MISMATCH gpu={n_g} oracle={n_e}")
    n = min(n_g, n_e)
    bad = np.zeros(n, bool)
    for g, e in zip(got, exp):
        bad |= g[:n] != e[:n]
    nb = int(bad.sum())
    if nb or n_g != n_e:
        idx = np.flatnonzero(bad)[:5]
        say(f"  {name}: {nb}/{n} differ; first at {idx.tolist()}")
        for i in idx:
            say("     i=%d gpu=%s oracle=%s" % (i, [g[i].item() for g in got], [e[i].item() for e in exp]))
        return False
    say(f"  {name}: ok ({n})")
    return True


def run_case(name, params, L, R, bf, maxD):
    say(f"== {name} params={params} shape={L.shape}")
    ok = True
    oL, oR = O.OracleExtractor(*params), O.OracleExtractor(*params)
    t = time.time(); okL, odL = oL.extract(L); okR, odR = oR.extract(R); t_or = time.time() - t
    gL, gR = B.OrbExtractor(*params), B.OrbExtractor(*params)
    t = time.time(); gkL, gdL = gL.extract(L); t_first = time.time() - t
    t = time.time(); gkL, gdL = gL.extract(L); gkR, gdR = gR.extract(R); t_g = time.time() - t
    say(f"  oracle 2 images {t_or*1e3:.1f} ms; gpu first call {t_first*1e3:.1f} ms, 2 images warm {t_g*1e3:.2f} ms")
    tb = gL.tables()
    ok &= np.array_equal(tb["scale"], oL.scale_factors) and np.array_equal(tb["features_per_level"], oL.features_per_level)
    for l in range(params[2]):
        a, b = gL.level(l), oL.level(l)
        if a.shape != b.shape or not np.array_equal(a, b):
            ok = False
            say(f"  L{l} pyramid: MISMATCH shape {a.shape} vs {b.shape}, diff px = {(a != b).sum() if a.shape == b.shape else -1}")
            if a.shape == b.shape:
                ys, xs = np.nonzero(a != b)
                say("     first diffs", list(zip(ys[:5].tolist(), xs[:5].tolist())), a[ys[:5], xs[:5]], b[ys[:5], xs[:5]])
        else:
            say(f"  L{l} pyramid ok {a.shape}")
        ok &= cmp_seq(f"L{l} candidates", gL.debug_candidates(l), oL.level_candidates(l))
        ok &= cmp_seq(f"L{l} keypoints", gL.debug_level_keypoints(l), oL.level_keypoints(l))
        say(f"     tie-sensitive(oracle)={oL.tie_sensitive(l)}")
    for nm, gk, gd, ok_, od in (("left", gkL, gdL, okL, odL), ("right", gkR, gdR, okR, odR)):
        same_k = len(gk) == len(ok_) and gk.tobytes() == ok_.tobytes()
        same_d = gd.shape == od.shape and np.array_equal(gd, od)
        say(f"  final {nm}: n={len(gk)}/{len(ok_)} keypoints {'ok' if same_k else 'MISMATCH'} descriptors {'ok' if same_d else 'MISMATCH'}")
        if not same_k and len(gk) == len(ok_):
            for f in gk.dtype.names:
                nb = int((gk[f] != ok_[f]).sum())
                if nb:
                    i = int(np.flatnonzero(gk[f] != ok_[f])[0])
                    say(f"     field {f}: {nb} differ, first i={i} gpu={gk[f][i]!r} oracle={ok_[f][i]!r}")
        if not same_d and gd.shape == od.shape:
            rows = np.flatnonzero((gd != od).any(axis=1))
            bits = np.unpackbits(gd ^ od, axis=1).sum()
            say(f"     {len(rows)} descriptor rows differ ({bits} bits); first rows {rows[:5].tolist()}")
        ok &= same_k and same_d
    # stereo on the ORACLE's features (isolates the matcher), then on the GPU's own
    our, odp, obd, obi, obs = O.stereo_match(oL, oR, okL, odL, okR, odR, bf, maxD, debug=True)
    if gL.level(0).shape == oL.level(0).shape:
        gur, gdp, gbd, gbi, gbs = B.stereo_match(gL, gR, okL, odL, okR, odR, bf, maxD, debug=True)
        ok &= cmp_seq("stereo best_dist", [gbd], [obd])
        ok &= cmp_seq("stereo best_idx", [gbi], [obi])
        ok &= cmp_seq("stereo sad(pre-reject)", [gbs], [obs])
        ok &= cmp_seq("stereo uRight bits", [gur.view(np.uint32)], [our.view(np.uint32)])
        ok &= cmp_seq("stereo depth bits", [gdp.view(np.uint32)], [odp.view(np.uint32)])
        say(f"     matched gpu={(gur >= 0).sum()} oracle={(our >= 0).sum()} max|duR|={np.abs(gur - our).max() if len(our) else 0}")
    # batch API
    try:
        sb = B.StereoBatch(params, L.shape[1], L.shape[0], 3)
        Ls = np.stack([L, R, L]); Rs = np.stack([R, L, L])
        res = sb.run(Ls, Rs, bf, maxD)
        r0 = res[0]
        b_ok = (r0["kL"].tobytes() == okL.tobytes() and np.array_equal(r0["dL"], odL) and r0["kR"].tobytes() == okR.tobytes()
                and np.array_equal(r0["dR"], odR) and r0["u_right"].tobytes() == our.tobytes() and r0["depth"].tobytes() == odp.tobytes())
        say(f"  batch pair0 vs oracle: {'ok' if b_ok else 'MISMATCH'} (launches/batch={sb.kernel_launches}, K={sb.K})")
        r1 = res[1]
        b1 = r1["kL"].tobytes() == okR.tobytes() and r1["kR"].tobytes() == okL.tobytes()
        say(f"  batch pair1 (views swapped) features: {'ok' if b1 else 'MISMATCH'}")
        ok &= b_ok and b1
        sb.close()
    except Exception:
        ok = False
        say("  batch API raised:\n" + traceback.format_exc())
    say(f"== {name}: {'ALL OK' if ok else 'FAILED'}")
    return ok


def main():
    from conftest import load_golden
    O.build()
    say("devices:", B.device_count())
    cases = sys.argv[1:] or ["odd_synth_s5_500", "kitti_synth_s0_2000", "kitti_natural_2000", "kitti_synth_s7_1000",
                             "euroc_synth_s3_1200", "hd_synth_s1_4000"]
    allok = True
    for c in cases:
        try:
            g = load_golden(c)
            allok &= run_case(c, g["params_t"], g["left"], g["right"], float(g["bf"]), float(g["maxD"]))
        except BaseException as e:  # pytest.skip raises too
            allok = False
            say(f"== {c}: EXCEPTION {type(e).__name__}: {e}\n" + traceback.format_exc())
    say("DIAG", "PASS" if allok else "FAIL")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
