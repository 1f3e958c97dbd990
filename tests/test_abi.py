"""CPU: the C-ABI library loads, exports every symbol include/pslot_orb.h declares, and fails loudly
(never silently falls back) when no CUDA device is present."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def B():
    from pointslot_b200 import binding, build
    build.build()
    return binding


def test_header_symbols_exported(B):
    hdr = (ROOT / "include" / "pslot_orb.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pslot_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(B.ABI_SYMBOLS)
    L = B.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.pslot_abi_version() == 1


def test_keypoint_layout(B):
    assert B.KEYPOINT_DTYPE.itemsize == 28  # cv::KeyPoint
    hdr = (ROOT / "include" / "pslot_orb.h").read_text()
    assert "float x, y, size, angle, response;" in hdr and "int32_t octave, class_id;" in hdr


def test_hamming_host(B):
    rng = np.random.default_rng(1)
    for _ in range(100):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert B.hamming256(a, b) == int(np.unpackbits(a ^ b).sum())


def test_no_silent_cpu_fallback(B):
    """Without a GPU every compute entry point must raise; with one this test has nothing to check."""
    try:
        n = B.device_count()
    except B.PslotError as e:
        n = 0
        assert e.code == -3
    if n > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(B.PslotError):
        B.OrbExtractor(1000, 1.2, 8, 20, 7)
    with pytest.raises(B.PslotError):
        B.StereoBatch((1000, 1.2, 8, 20, 7), 1242, 375, 4)


def test_product_never_imports_oracle():
    """The shipped package must not route through the CPU oracle."""
    for p in (ROOT / "pointslot_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".h", ".cc", ".cpp", ".hpp") and p.is_file():
            txt = p.read_text(errors="ignore")
            assert "import oracle" not in txt and "from oracle" not in txt and "orb_oracle" not in txt, p


def test_record_layouts_match_the_header():
    """The numpy / ctypes records the binding passes must have the sizes the C header asserts."""
    import ctypes as C
    from pointslot_b200 import binding as B
    assert B.KEYPOINT_DTYPE.itemsize == 28
    assert B.MAP_POINT_DTYPE.itemsize == 24
    assert B.PROJ_POINT_DTYPE.itemsize == 32
    assert B.AREA_QUERY_DTYPE.itemsize == 20
    assert C.sizeof(B.GridSpec) == 24
    import oracle
    assert oracle.MAP_POINT_DTYPE == B.MAP_POINT_DTYPE and oracle.PROJ_POINT_DTYPE == B.PROJ_POINT_DTYPE
