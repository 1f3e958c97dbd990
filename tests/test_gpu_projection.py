"""GPU parity of pslot_search_by_projection (ORBmatcher::SearchByProjection, reference src/ORBmatcher.cc:68-155) against
the oracle: identical match per map point (index work: bit-exact), including the reference's order dependence."""
import numpy as np
import pytest

from test_oracle_projection import make_case, run_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from pointslot_b200 import binding
    assert binding.device_count() > 0
    return binding


@pytest.mark.parametrize("seed,th,nn,n_feat,n_mp", [(1, 1.0, 0.8, 1200, 1800), (2, 3.0, 0.8, 1200, 1800),
                                                    (3, 5.0, 0.6, 2000, 4000), (4, 1.0, 0.9, 500, 200)])
def test_search_by_projection_vs_oracle(seed, th, nn, n_feat, n_mp, B, oracle_mod):
    c = make_case(oracle_mod, seed, n_feat=n_feat, n_mp=n_mp)
    ref, nref = run_oracle(oracle_mod, c, th, nn)
    e = B.OrbExtractor(n_feat, 1.2, 8, 20, 7)
    spec = B.GridSpec.for_image(c["W"], c["H"])
    B.assign_grid(e, c["k"], spec)
    got, n, sweeps = B.search_by_projection(e, c["mps"], c["md"], c["d"], c["ur"], c["occupied"], c["sf"], th, nn)
    assert np.array_equal(got, ref) and n == nref
    assert 2 <= sweeps <= 64  # dependency chains are short; the sequential loop would be n_mp steps
    # same call again on the same context (scratch reuse), and without pre-occupied features
    got2, n2, _ = B.search_by_projection(e, c["mps"], c["md"], c["d"], c["ur"], c["occupied"], c["sf"], th, nn)
    assert np.array_equal(got2, ref)
    c0 = dict(c); c0["occupied"] = np.zeros_like(c["occupied"])
    ref0, nref0 = run_oracle(oracle_mod, c0, th, nn)
    got0, n0, _ = B.search_by_projection(e, c["mps"], c["md"], c["d"], c["ur"], None, c["sf"], th, nn)
    assert np.array_equal(got0, ref0) and n0 == nref0


def test_search_by_projection_edge_cases(B, oracle_mod):
    c = make_case(oracle_mod, 7, n_feat=600, n_mp=300)
    e = B.OrbExtractor(600, 1.2, 8, 20, 7)
    with pytest.raises(B.PslotError):  # no grid yet
        B.search_by_projection(e, c["mps"], c["md"], c["d"], c["ur"], None, c["sf"], 1.0, 0.8)
    B.assign_grid(e, c["k"], B.GridSpec.for_image(c["W"], c["H"]))
    got, n, _ = B.search_by_projection(e, c["mps"][:0], c["md"][:0], c["d"], c["ur"], None, c["sf"], 1.0, 0.8)
    assert len(got) == 0 and n == 0
    off = c["mps"].copy(); off["flags"] = 0  # nothing in view
    got, n, _ = B.search_by_projection(e, off, c["md"], c["d"], c["ur"], None, c["sf"], 1.0, 0.8)
    assert np.all(got == -1) and n == 0
    bad = c["mps"].copy(); bad["level"][5] = 9; bad["flags"][5] = 3
    with pytest.raises(B.PslotError):
        B.search_by_projection(e, bad, c["md"], c["d"], c["ur"], None, c["sf"], 1.0, 0.8)
    with pytest.raises(B.PslotError):  # frame arrays must describe the gridded keypoints
        B.search_by_projection(e, c["mps"], c["md"], c["d"][:-1], c["ur"][:-1], None, c["sf"], 1.0, 0.8)


@pytest.mark.parametrize("seed,mode,ori,n_feat", [(1, "neutral", True, 1200), (2, "forward", True, 1200),
                                                  (3, "backward", True, 2000), (4, "neutral", False, 600)])
def test_search_by_projection_last_vs_oracle(seed, mode, ori, n_feat, B, oracle_mod):
    """Last-frame variant (ORBmatcher.cc:1613-1756): matches, rotation-rejected set and return value."""
    from test_oracle_projection import make_last_frame_case, run_oracle_last
    c = make_last_frame_case(oracle_mod, seed, n_feat=n_feat, mode=mode)
    rm, rrej, rn = run_oracle_last(oracle_mod, c, ori)
    e = B.OrbExtractor(n_feat, 1.2, 8, 20, 7)
    B.assign_grid(e, c["k"], B.GridSpec.for_image(c["W"], c["H"]))
    m, rej, n, sweeps = B.search_by_projection_last(e, c["pts"], c["pd"], c["d"], c["ur"], c["occupied"], ori)
    assert np.array_equal(m, rm) and np.array_equal(rej, rrej) and n == rn
    assert 2 <= sweeps <= 64
    # the local-map search still works on the same context afterwards (shared scratch, different mode)
    c2 = make_case(oracle_mod, seed, n_feat=n_feat, n_mp=500)
    B.assign_grid(e, c2["k"], B.GridSpec.for_image(c2["W"], c2["H"]))
    ref, nref = run_oracle(oracle_mod, c2, 3.0, 0.8)
    got, n2, _ = B.search_by_projection(e, c2["mps"], c2["md"], c2["d"], c2["ur"], c2["occupied"], c2["sf"], 3.0, 0.8)
    assert np.array_equal(got, ref) and n2 == nref


@pytest.mark.parametrize("seed,ori,orb_dist,n_feat", [(1, True, 100, 1200), (2, True, 64, 2000), (3, False, 100, 600), (4, True, 30, 1200)])
def test_search_by_projection_kf_vs_oracle(seed, ori, orb_dist, n_feat, B, oracle_mod):
    """Relocalisation overload (ORBmatcher.cc:2529-2654): no stereo test, every association blocks, ORBdist acceptance,
    30-degree rotation bins; the oracle is pinned to the reference's own function (tests/test_oracle_reference.py)."""
    from test_oracle_projection import make_last_frame_case, COLS, ROWS
    c = make_last_frame_case(oracle_mod, seed, n_feat=n_feat, mode="neutral")
    c["pts"]["angle"] = (c["pts"]["angle"] + np.float32(60.0)) % np.float32(360.0)  # off the first 30-degree bin
    cs, idx = oracle_mod.assign_grid(c["k"], 0.0, 0.0, c["inv_w"], c["inv_h"])
    rm, rrej, rn = oracle_mod.search_by_projection_kf(c["k"], c["d"], cs, idx, 0.0, 0.0, c["inv_w"], c["inv_h"], COLS, ROWS,
                                                      c["pts"], c["pd"], c["occupied"], orb_dist, ori)
    e = B.OrbExtractor(n_feat, 1.2, 8, 20, 7)
    B.assign_grid(e, c["k"], B.GridSpec.for_image(c["W"], c["H"]))
    m, rej, n = B.search_by_projection_kf(e, c["pts"], c["pd"], c["d"], c["occupied"], orb_dist, ori)
    assert np.array_equal(m, rm) and np.array_equal(rej, rrej) and n == rn
    assert (m >= 0).sum() > 50
    taken = m[m >= 0]
    assert len(set(taken.tolist())) == len(taken) and not c["occupied"][taken].any()  # every feature matched at most once


def test_search_by_projection_kf_serves_the_sim3_overload(B, oracle_mod):
    """Loop closing (ORBmatcher.cc:413-527) = the key-frame search with the level window [L - 1, L], TH_LOW = 50, no rotation
    check, vpMatched as the occupancy (tests/test_oracle_reference.py pins that reading to the reference's own function)."""
    from test_oracle_projection import make_last_frame_case, COLS, ROWS
    c = make_last_frame_case(oracle_mod, 7, n_feat=1500, mode="neutral")
    c["pts"]["max_level"] = c["pts"]["min_level"] + 1
    cs, idx = oracle_mod.assign_grid(c["k"], 0.0, 0.0, c["inv_w"], c["inv_h"])
    rm, rrej, rn = oracle_mod.search_by_projection_kf(c["k"], c["d"], cs, idx, 0.0, 0.0, c["inv_w"], c["inv_h"], COLS, ROWS,
                                                      c["pts"], c["pd"], c["occupied"], 50, False)
    e = B.OrbExtractor(1500, 1.2, 8, 20, 7)
    B.assign_grid(e, c["k"], B.GridSpec.for_image(c["W"], c["H"]))
    m, rej, n = B.search_by_projection_kf(e, c["pts"], c["pd"], c["d"], c["occupied"], 50, False)
    assert np.array_equal(m, rm) and not rej.any() and n == rn and n > 100


@pytest.mark.parametrize("seed,th,nn", [(1, 1.0, 0.8), (2, 3.0, 0.8), (3, 5.0, 0.6)])
def test_search_by_projection_obj_vs_oracle(seed, th, nn, B, oracle_mod):
    """Object-point overload (ORBmatcher.cc:157-246): fixed 5 px window, levels L - 1 .. L + 1, TH_HIGH_FORDYNAMIC, features
    outside the detection box passed as occupied; oracle pinned to the reference's own functions."""
    c = make_case(oracle_mod, seed, n_feat=1500, n_mp=1800)
    k = c["k"]
    inside = (k["x"] >= c["W"] / 5) & (k["x"] < 4 * c["W"] / 5) & (k["y"] >= c["H"] / 6) & (k["y"] < 5 * c["H"] / 6)
    occ = (c["occupied"].astype(bool) | ~inside).astype(np.uint8)
    from test_oracle_projection import COLS, ROWS
    cs, idx = oracle_mod.assign_grid(k, 0.0, 0.0, c["inv_w"], c["inv_h"])
    ref, nref = oracle_mod.search_by_projection_obj(k, c["d"], c["ur"], cs, idx, 0.0, 0.0, c["inv_w"], c["inv_h"], COLS, ROWS,
                                                   c["mps"], c["md"], occ, c["sf"], th, nn)
    e = B.OrbExtractor(1500, 1.2, 8, 20, 7)
    B.assign_grid(e, k, B.GridSpec.for_image(c["W"], c["H"]))
    got, n, _ = B.search_by_projection_obj(e, c["mps"], c["md"], c["d"], c["ur"], occ, c["sf"], th, nn)
    assert np.array_equal(got, ref) and n == nref and n > 100
