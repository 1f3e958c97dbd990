"""The REFERENCE's own ORBextractor (oracle/_ref/libref_orb.so) driven from Python, with its OpenCV primitive calls
served by cv2 4.13 through ctypes callbacks.  Test infrastructure only.

The library exists only after the opt-in recipe `make -C oracle ref REF=/root/reference` has been run; that compiles and
then executes third-party source, so it needs a maintainer's go-ahead (oracle/ref_compat/README.md).  Round 1 never ran
it: everything here is skipped by the tests when the library is absent."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB = Path(__file__).resolve().parent / "_ref" / "libref_orb.so"
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])
_state = {}


def available() -> bool:
    try:
        import cv2  # noqa: F401
    except Exception:
        return False
    return LIB.exists()


def _view(ptr, w, h, step):
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(h, step))
    return a[:, :w]


def _load():
    if _state:
        return _state["lib"]
    import cv2
    lib = C.CDLL(str(LIB))
    u8p, sz = C.c_void_p, C.c_size_t
    RESIZE = C.CFUNCTYPE(None, u8p, C.c_int, C.c_int, sz, u8p, C.c_int, C.c_int, sz, C.c_int)
    BORDER = C.CFUNCTYPE(None, u8p, C.c_int, C.c_int, sz, u8p, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
    FAST = C.CFUNCTYPE(C.c_int, u8p, C.c_int, C.c_int, sz, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int)
    BLUR = C.CFUNCTYPE(None, u8p, C.c_int, C.c_int, sz, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int)
    ATAN2 = C.CFUNCTYPE(C.c_float, C.c_float, C.c_float)

    def resize(src, sw, sh, sstep, dst, dw, dh, dstep, interp):
        s = np.ascontiguousarray(_view(src, sw, sh, sstep))
        _view(dst, dw, dh, dstep)[:] = cv2.resize(s, (dw, dh), interpolation=interp)

    def border(src, w, h, sstep, dst, dstep, top, bottom, left, right, btype):
        s = np.array(_view(src, w, h, sstep))  # copy first: the reference passes a view of dst as src
        _view(dst, w + left + right, h + top + bottom, dstep)[:] = cv2.copyMakeBorder(s, top, bottom, left, right, btype & ~16)

    def fast(img, w, h, step, threshold, nonmax, out, cap):
        det = cv2.FastFeatureDetector_create(threshold=threshold, nonmaxSuppression=bool(nonmax))
        kps = det.detect(np.ascontiguousarray(_view(img, w, h, step)), None)
        if len(kps) <= cap:
            for i, k in enumerate(kps):
                out[3 * i], out[3 * i + 1], out[3 * i + 2] = k.pt[0], k.pt[1], k.response
        return len(kps)

    def blur(img, w, h, step, kw, kh, sx, sy, btype):
        v = _view(img, w, h, step)
        v[:] = cv2.GaussianBlur(np.ascontiguousarray(v), (kw, kh), sx, sigmaY=sy, borderType=btype)

    def atan2(y, x):
        return float(cv2.fastAtan2(float(y), float(x)))

    cbs = (RESIZE(resize), BORDER(border), FAST(fast), BLUR(blur), ATAN2(atan2))
    lib.ref_set_callbacks(*cbs)
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    lib.ref_destroy.argtypes = [C.c_void_p]
    lib.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    lib.ref_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _state.update(lib=lib, cbs=cbs)  # keep the callback objects alive
    return lib


class ReferenceExtractor:
    """ORB_SLAM2::ORBextractor of the reference tree (include/ORBextractor.h:45-111), as compiled from its own source."""

    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self._lib = _load()
        self._h = self._lib.ref_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        self.cap = nfeatures + 16 * nlevels + 64

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_destroy(self._h)
            self._h = None

    def extract(self, img: np.ndarray):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.cap
        while True:
            kps = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
            n = self._lib.ref_extract(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                      kps.ctypes.data, desc.ctypes.data, cap)
            if n >= 0:
                return kps[:n].copy(), desc[:n].copy()
            cap = -n

    def level(self, l: int) -> np.ndarray:
        w, h = C.c_int(0), C.c_int(0)
        self._lib.ref_level(self._h, l, None, 0, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        self._lib.ref_level(self._h, l, out.ctypes.data, out.strides[0], C.byref(w), C.byref(h))
        return out
