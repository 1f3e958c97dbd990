"""GPU: the C++ drop-in classes (ORB_SLAM2::ORBextractor, Frame stereo constructor, ORBmatcher::
DescriptorDistance) driven exactly as Tracking::GrabImageStereo drives the reference's, vs the oracle."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
SHIM = ROOT / "pointslot_b200" / "cpp"


def test_cpp_shim_matches_oracle(tmp_path, oracle_mod):
    from pointslot_b200 import binding as B
    from pointslot_b200.synth import synth_pair

    subprocess.check_call(["make", "-C", str(SHIM), "-s"])
    W, H = 640, 400
    params = (800, 1.2, 8, 20, 7)
    bf, fx = 300.0, 600.0
    L, R = synth_pair(W, H, 21)
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<6i3f", W, H, params[0], params[2], params[3], params[4], params[1], bf, fx))
        f.write(L.tobytes()); f.write(R.tobytes())
    subprocess.check_call([str(SHIM / "test_shim"), str(inp), str(out)])
    raw = out.read_bytes()
    N, Nr = struct.unpack_from("<2i", raw, 0)
    o = 8
    kL = np.frombuffer(raw, B.KEYPOINT_DTYPE, N, o); o += 28 * N
    dL = np.frombuffer(raw, np.uint8, 32 * N, o).reshape(N, 32); o += 32 * N
    kR = np.frombuffer(raw, B.KEYPOINT_DTYPE, Nr, o); o += 28 * Nr
    dR = np.frombuffer(raw, np.uint8, 32 * Nr, o).reshape(Nr, 32); o += 32 * Nr
    ur = np.frombuffer(raw, np.float32, N, o); o += 4 * N
    dp = np.frombuffer(raw, np.float32, N, o); o += 4 * N
    pw, ph = struct.unpack_from("<2i", raw, o); o += 8
    p1 = np.frombuffer(raw, np.uint8, pw * ph, o).reshape(ph, pw); o += pw * ph
    (dd,) = struct.unpack_from("<i", raw, o)

    O = oracle_mod
    oL, oR = O.OracleExtractor(*params), O.OracleExtractor(*params)
    okL, odL = oL.extract(L); okR, odR = oR.extract(R)
    maxD = B.max_disparity(bf, fx)  # the shim sets mb = mbf/fx before matching (parity rule 2)
    our, odp = O.stereo_match(oL, oR, okL, odL, okR, odR, bf, maxD)
    assert kL.tobytes() == okL.tobytes() and kR.tobytes() == okR.tobytes()
    assert np.array_equal(dL, odL) and np.array_equal(dR, odR)
    assert np.array_equal(ur >= 0, our >= 0) and np.all(np.abs(ur - our) <= 1e-3)
    assert np.allclose(dp, odp, rtol=1e-5, atol=0)
    assert np.array_equal(p1, oL.level(1))  # mvImagePyramid mirror
    assert dd == O.hamming256(odL[0], odL[1])
