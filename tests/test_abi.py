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


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: include/pslot_orb.h compiles as strict C11 (no C++ in the signatures), a C program links against
    libpslot_orb.so and calls the entry points that need no GPU; a compute call without a device returns an error code and a
    message, never a result."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    from pointslot_b200 import build
    lib = build.build()
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "pslot_orb.h"
int main(void)
{
    unsigned char a[32], b[32];
    memset(a, 0, 32); memset(b, 0xff, 32);
    if (pslot_hamming256(a, b) != 256 || pslot_hamming256(a, a) != 0) return 2;
    if (sizeof(pslot_keypoint) != 28) return 3;
    printf("abi %d\n", pslot_abi_version());
    pslot_orb *ctx = NULL;
    int rc = pslot_orb_create(&ctx, 1000, 1.2f, 8, 20, 5, 0);
    if (rc == PSLOT_OK) { pslot_orb_destroy(ctx); printf("device present\n"); return 0; }
    if (ctx != NULL) return 4;
    const char *msg = pslot_last_error();
    if (!msg || !msg[0]) return 5;
    printf("no device: rc=%d (%s)\n", rc, msg);
    /* bad arguments are refused before anything else */
    if (pslot_orb_create(NULL, 1000, 1.2f, 8, 20, 5, 0) != PSLOT_ERR_ARG) return 6;
    return 0;
}
''')
    exe = tmp_path / "abi"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
           str(lib), f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi " in r.stdout
