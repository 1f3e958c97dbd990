"""CPU ORACLE — test infrastructure, NOT product code.

ctypes wrapper over oracle/liborb_oracle.so (built from oracle/orb_oracle.c by `make -C oracle`).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; pointslot_b200 never does.

Parity status: "parity unpinned" by the reference (no tests, un-vendored OpenCV); the primitives are
pinned against OpenCV 4.13.0 (tests/test_oracle_cv2.py, tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liborb_oracle.so"

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
assert KEYPOINT_DTYPE.itemsize == 28


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc, seconds)."""
    src = _HERE / "orb_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        vp, ip, fp, u8p = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_void_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [vp]
        for name in ("orc_scale_factors", "orc_inv_scale_factors", "orc_level_sigma2", "orc_inv_level_sigma2"):
            getattr(L, name).restype = fp
            getattr(L, name).argtypes = [vp]
        for name in ("orc_features_per_level", "orc_umax"):
            getattr(L, name).restype = ip
            getattr(L, name).argtypes = [vp]
        L.orc_extract.restype = C.c_int
        L.orc_extract.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_ssize_t, vp, u8p, C.c_int, ip]
        L.orc_level_size.argtypes = [vp, C.c_int, ip, ip]
        L.orc_level_ptr.restype = vp
        L.orc_level_ptr.argtypes = [vp, C.c_int]
        L.orc_blurred_ptr.restype = vp
        L.orc_blurred_ptr.argtypes = [vp, C.c_int]
        L.orc_level_candidates.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
        L.orc_level_keypoints.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp]
        L.orc_level_tie_sensitive.argtypes = [vp, C.c_int]
        L.orc_stereo_match.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_float, C.c_float,
                                       vp, vp, vp, vp, vp]
        L.orc_hamming256.argtypes = [vp, vp]
        L.orc_resize_linear_u8.argtypes = [vp, C.c_int, C.c_int, C.c_ssize_t, vp, C.c_int, C.c_int, C.c_ssize_t]
        L.orc_resize_linear_u8.restype = None
        L.orc_fast9_16.argtypes = [vp, C.c_int, C.c_int, C.c_ssize_t, C.c_int, C.c_int, vp, vp, vp]
        L.orc_gaussian7x7_s2.argtypes = [vp, C.c_int, C.c_int, C.c_ssize_t, vp, C.c_ssize_t]
        L.orc_gaussian7x7_s2.restype = None
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_cv_round.argtypes = [C.c_double]
        L.orc_cvt_gray.argtypes = [vp, C.c_int, C.c_int, C.c_ssize_t, C.c_int, C.c_int, vp, C.c_ssize_t]
        L.orc_cvt_gray.restype = None
        L.orc_steer_coeffs.argtypes = [vp, C.c_int, vp, vp]
        L.orc_steer_coeffs.restype = None
        L.orc_glibc_sincosf_n.argtypes = [vp, C.c_int, vp, vp]
        L.orc_glibc_sincosf_n.restype = None
        L.orc_assign_features.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_float, vp, vp, vp, vp]
        L.orc_assign_grid.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, vp, vp]
        L.orc_features_in_area.argtypes = [vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                           C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, vp]
        L.orc_search_by_projection.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                               C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_float, C.c_float, vp]
        L.orc_search_by_projection_obj.argtypes = L.orc_search_by_projection.argtypes
        L.orc_search_by_projection_last.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                                    C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp]
        L.orc_search_by_projection_kf.argtypes = [vp, vp, C.c_int, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                                  C.c_int, vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.orc_cvorb_detect_and_compute.argtypes = [vp, C.c_int, C.c_int, C.c_ssize_t, vp, C.c_ssize_t, C.c_int, C.c_float,
                                                   C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, ip]
        L.orc_bow_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
        L.orc_bow_transform.restype = None
        L.orc_bow_vectors.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, ip]
        L.orc_distribute_octtree.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, vp, vp, vp, ip]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
# primitives (for the known-answer tests)
# ---------------------------------------------------------------------------------------------
def resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dw, dh, dw)
    return dst


def fast9_16(img: np.ndarray, threshold: int):
    """-> (x, y, response) int arrays, raster order (cv::FAST with nonmaxSuppression=true)."""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    h, w = img.shape
    cap = ((w + 1) // 2) * ((h + 1) // 2) + 1
    x = np.empty(cap, np.int32); y = np.empty(cap, np.int32); r = np.empty(cap, np.int32)
    n = lib().orc_fast9_16(_ptr(img), w, h, img.strides[0], threshold, cap, _ptr(x), _ptr(y), _ptr(r))
    return x[:n].copy(), y[:n].copy(), r[:n].copy()


def gaussian7x7(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_gaussian7x7_s2(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(out), out.strides[0])
    return out


def fast_atan2(y: float, x: float) -> float:
    return float(lib().orc_fast_atan2(float(y), float(x)))


def cv_round(v: float) -> int:
    return int(lib().orc_cv_round(float(v)))


def cvt_gray(img: np.ndarray, rgb: bool) -> np.ndarray:
    """cv::cvtColor(..., RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY) of an interleaved uint8 [H,W,3|4] image."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, ch = img.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_cvt_gray(_ptr(img), w, h, img.strides[0], ch, 1 if rgb else 0, _ptr(out), out.strides[0])
    return out


def steer_coeffs(angle_deg: np.ndarray):
    """(a, b) of computeOrbDescriptor for float32 angles in degrees (libm cos/sin in double, rounded to float)."""
    ang = np.ascontiguousarray(angle_deg, np.float32)
    a = np.empty_like(ang); b = np.empty_like(ang)
    lib().orc_steer_coeffs(_ptr(ang), ang.size, _ptr(a), _ptr(b))
    return a, b


def glibc_sincosf(x: np.ndarray):
    """-> (cosf(x), sinf(x)) by the oracle's restatement of glibc's algorithm (orb_oracle.c: orc_glibc_cosf)."""
    x = np.ascontiguousarray(x, np.float32)
    c = np.empty_like(x); s = np.empty_like(x)
    lib().orc_glibc_sincosf_n(_ptr(x), x.size, _ptr(c), _ptr(s))
    return c, s


def libm_sincosf(x: np.ndarray):
    """-> (cosf(x), sinf(x)) by the host's libm, the function the reference binary calls (ORBextractor.cc:112-113)."""
    m = C.CDLL("libm.so.6")
    m.cosf.restype = C.c_float; m.cosf.argtypes = [C.c_float]
    m.sinf.restype = C.c_float; m.sinf.argtypes = [C.c_float]
    x = np.ascontiguousarray(x, np.float32)
    return (np.array([m.cosf(float(v)) for v in x.ravel()], np.float32).reshape(x.shape),
            np.array([m.sinf(float(v)) for v in x.ravel()], np.float32).reshape(x.shape))


def hamming256(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    assert a.size == 32 and b.size == 32
    return int(lib().orc_hamming256(_ptr(a), _ptr(b)))


def distribute_octtree(x, y, resp, min_x, max_x, min_y, max_y, n_features):
    """-> (x, y, resp, tie_sensitive) of the retained keys in list order."""
    x = np.ascontiguousarray(x, np.int32); y = np.ascontiguousarray(y, np.int32)
    resp = np.ascontiguousarray(resp, np.int32)
    cap = max(len(x), 1) + 8
    ox = np.empty(cap, np.int32); oy = np.empty(cap, np.int32); orr = np.empty(cap, np.int32)
    tie = C.c_int(0)
    n = lib().orc_distribute_octtree(_ptr(x), _ptr(y), _ptr(resp), len(x), min_x, max_x, min_y, max_y,
                                     n_features, cap, _ptr(ox), _ptr(oy), _ptr(orr), C.byref(tie))
    return ox[:n].copy(), oy[:n].copy(), orr[:n].copy(), bool(tie.value)


def assign_grid(kps, min_x, min_y, inv_w, inv_h, cols=64, rows=48):
    """Frame::AssignFeaturesToGrid (Frame.cc:1636-1655) -> (cell_start [cols*rows+1], indices [n_in_grid]);
    cell = i * rows + j for mGrid[i][j]."""
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE)
    cs = np.zeros(cols * rows + 1, np.int32); idx = np.zeros(max(len(kps), 1), np.int32)
    n = lib().orc_assign_grid(_ptr(kps), len(kps), min_x, min_y, inv_w, inv_h, cols, rows, _ptr(cs), _ptr(idx))
    if n < 0:
        raise ValueError("bad grid arguments")
    return cs, idx[:n].copy()


def assign_features(kps, mask, object_ids, depth=None, depth_limit=0.0):
    """Frame::AssignFeatures (Frame.cc:759-977) -> (cls [n], groups, effective): groups[0] = indices of the background
    features, groups[1 + j] = those of detection j, each in push_back order."""
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE)
    mask = np.ascontiguousarray(mask, np.uint8)
    ids = np.ascontiguousarray(object_ids, np.int32)
    n, J = len(kps), len(ids)
    d = None if depth is None else np.ascontiguousarray(depth, np.float32)
    cls = np.zeros(max(n, 1), np.int32); order = np.zeros(max(n, 1), np.int32)
    cs = np.zeros(J + 2, np.int32); eff = np.zeros(max(J, 1), np.int32)
    r = lib().orc_assign_features(_ptr(kps), n, _ptr(mask), mask.shape[1], mask.shape[0], mask.strides[0], _ptr(ids), J,
                                  None if d is None else _ptr(d), depth_limit, _ptr(cls), _ptr(order), _ptr(cs), _ptr(eff))
    if r < 0:
        raise ValueError("bad arguments")
    return cls[:n], [order[cs[c]:cs[c + 1]].copy() for c in range(J + 1)], (eff[:J] if d is not None else None)


def features_in_area(kps, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows, x, y, r, min_level=-1,
                     max_level=-1):
    """Frame::GetFeaturesInArea (Frame.cc:1808-1861) -> indices in the reference's visiting order."""
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE)
    cell_start = np.ascontiguousarray(cell_start, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    out = np.zeros(max(len(kps), 1), np.int32)
    n = lib().orc_features_in_area(_ptr(kps), _ptr(cell_start), _ptr(indices), min_x, min_y, inv_w, inv_h, cols, rows,
                                   x, y, r, min_level, max_level, len(out), _ptr(out))
    return out[:n].copy()


MAP_POINT_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"),
                            ("level", "<i4"), ("flags", "<i4")])


def search_by_projection(kps_un, frame_desc, u_right, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows,
                         mps, mp_desc, occupied, scale_factors, th, nn_ratio):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:68-155) -> (match_of_mp [n_mp], nmatches)."""
    kps_un = np.ascontiguousarray(kps_un, KEYPOINT_DTYPE); frame_desc = np.ascontiguousarray(frame_desc, np.uint8)
    u_right = np.ascontiguousarray(u_right, np.float32); mps = np.ascontiguousarray(mps, MAP_POINT_DTYPE)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8); occupied = np.ascontiguousarray(occupied, np.uint8)
    cell_start = np.ascontiguousarray(cell_start, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    out = np.full(max(len(mps), 1), -1, np.int32)
    n = lib().orc_search_by_projection(_ptr(kps_un), _ptr(frame_desc), _ptr(u_right), len(kps_un), _ptr(cell_start),
                                       _ptr(indices), min_x, min_y, inv_w, inv_h, cols, rows, _ptr(mps), _ptr(mp_desc),
                                       len(mps), _ptr(occupied), _ptr(sf), th, nn_ratio, _ptr(out))
    return out[:len(mps)].copy(), int(n)


PROJ_POINT_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"),
                             ("max_level", "<i4"), ("angle", "<f4"), ("flags", "<i4")])



def search_by_projection_obj(kps_un, frame_desc, u_right, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows,
                         mps, mp_desc, occupied, scale_factors, th, nn_ratio):
    """ORBmatcher::SearchByProjection(F, nOrder, vpMapObjectPoints, th) (ORBmatcher.cc:157-246) on one object's feature set
    (`occupied` also carries !isInBBox) -> (match_of_mp [n_mp], nmatches)."""
    kps_un = np.ascontiguousarray(kps_un, KEYPOINT_DTYPE); frame_desc = np.ascontiguousarray(frame_desc, np.uint8)
    u_right = np.ascontiguousarray(u_right, np.float32); mps = np.ascontiguousarray(mps, MAP_POINT_DTYPE)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8); occupied = np.ascontiguousarray(occupied, np.uint8)
    cell_start = np.ascontiguousarray(cell_start, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    out = np.full(max(len(mps), 1), -1, np.int32)
    n = lib().orc_search_by_projection_obj(_ptr(kps_un), _ptr(frame_desc), _ptr(u_right), len(kps_un), _ptr(cell_start),
                                       _ptr(indices), min_x, min_y, inv_w, inv_h, cols, rows, _ptr(mps), _ptr(mp_desc),
                                       len(mps), _ptr(occupied), _ptr(sf), th, nn_ratio, _ptr(out))
    return out[:len(mps)].copy(), int(n)


PROJ_POINT_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"),
                             ("max_level", "<i4"), ("angle", "<f4"), ("flags", "<i4")])

def search_by_projection_last(kps_un, frame_desc, u_right, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows,
                              pts, pt_desc, occupied, check_orientation=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) after the projection (ORBmatcher.cc:1671-1756)
    -> (match_of_pt, rejected, nmatches)."""
    kps_un = np.ascontiguousarray(kps_un, KEYPOINT_DTYPE); frame_desc = np.ascontiguousarray(frame_desc, np.uint8)
    u_right = np.ascontiguousarray(u_right, np.float32); pts = np.ascontiguousarray(pts, PROJ_POINT_DTYPE)
    pt_desc = np.ascontiguousarray(pt_desc, np.uint8); occupied = np.ascontiguousarray(occupied, np.uint8)
    cell_start = np.ascontiguousarray(cell_start, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    out = np.full(max(len(pts), 1), -1, np.int32); rej = np.zeros(max(len(pts), 1), np.uint8)
    n = lib().orc_search_by_projection_last(_ptr(kps_un), _ptr(frame_desc), _ptr(u_right), len(kps_un), _ptr(cell_start),
                                            _ptr(indices), min_x, min_y, inv_w, inv_h, cols, rows, _ptr(pts), _ptr(pt_desc),
                                            len(pts), _ptr(occupied), int(check_orientation), _ptr(out), _ptr(rej))
    return out[:len(pts)].copy(), rej[:len(pts)].copy(), int(n)


def search_by_projection_kf(kps_un, frame_desc, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows, pts, pt_desc,
                            occupied, orb_dist, check_orientation=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) after the projection
    (ORBmatcher.cc:2583-2654) -> (match_of_pt, rejected, nmatches)."""
    kps_un = np.ascontiguousarray(kps_un, KEYPOINT_DTYPE); frame_desc = np.ascontiguousarray(frame_desc, np.uint8)
    pts = np.ascontiguousarray(pts, PROJ_POINT_DTYPE)
    pt_desc = np.ascontiguousarray(pt_desc, np.uint8); occupied = np.ascontiguousarray(occupied, np.uint8)
    cell_start = np.ascontiguousarray(cell_start, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    out = np.full(max(len(pts), 1), -1, np.int32); rej = np.zeros(max(len(pts), 1), np.uint8)
    n = lib().orc_search_by_projection_kf(_ptr(kps_un), _ptr(frame_desc), len(kps_un), _ptr(cell_start), _ptr(indices), min_x,
                                          min_y, inv_w, inv_h, cols, rows, _ptr(pts), _ptr(pt_desc), len(pts), _ptr(occupied),
                                          int(orb_dist), int(check_orientation), _ptr(out), _ptr(rej))
    return out[:len(pts)].copy(), rej[:len(pts)].copy(), int(n)


def cvorb_detect_and_compute(img, mask=None, nfeatures=1000, scale_factor=1.2, nlevels=8, edge_threshold=19,
                             fast_threshold=20):
    """cv::ORB::create(nfeatures, scale_factor, nlevels, edge_threshold)->detectAndCompute(img, mask) (reference
    src/Frame.cc:2623-2627) -> (keypoints ordered by (octave, y, x), descriptors)."""
    img = np.ascontiguousarray(img, np.uint8)
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    cap = nfeatures + 64 * nlevels + 256
    while True:
        kps = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rc = lib().orc_cvorb_detect_and_compute(_ptr(img), img.shape[1], img.shape[0], img.strides[0],
                                                None if m is None else _ptr(m), 0 if m is None else m.strides[0],
                                                nfeatures, scale_factor, nlevels, edge_threshold, fast_threshold,
                                                _ptr(kps), _ptr(desc), cap, C.byref(n))
        if rc == -2:
            cap *= 4
            continue
        if rc:
            raise ValueError("bad cv::ORB arguments")
        return kps[:n.value].copy(), desc[:n.value].copy()


# ---------------------------------------------------------------------------------------------
# the extractor / stereo matcher
# ---------------------------------------------------------------------------------------------
class OracleExtractor:
    """Restatement of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111)."""

    def __init__(self, nfeatures: int, scale_factor: float, nlevels: int, ini_th: int, min_th: int):
        self._h = lib().orc_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        if not self._h:
            raise ValueError("bad extractor parameters")
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.cap = nfeatures + 16 * nlevels + 64

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def _f(self, fn, n=None):
        p = getattr(lib(), fn)(self._h)
        return np.ctypeslib.as_array(p, shape=(n or self.nlevels,)).copy()

    @property
    def scale_factors(self): return self._f("orc_scale_factors")
    @property
    def inv_scale_factors(self): return self._f("orc_inv_scale_factors")
    @property
    def level_sigma2(self): return self._f("orc_level_sigma2")
    @property
    def inv_level_sigma2(self): return self._f("orc_inv_level_sigma2")
    @property
    def features_per_level(self): return self._f("orc_features_per_level")
    @property
    def umax(self): return self._f("orc_umax", 16)

    def extract(self, img: np.ndarray):
        """-> (keypoints structured array [n], descriptors uint8 [n,32])"""
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        cap = self.cap
        while True:
            kps = np.zeros(cap, KEYPOINT_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = C.c_int(0)
            rc = lib().orc_extract(self._h, _ptr(img), img.shape[1], img.shape[0], img.strides[0],
                                   _ptr(kps), _ptr(desc), cap, C.byref(n))
            if rc == -1:
                cap = n.value + 8
                continue
            if rc != 0:
                raise RuntimeError(f"orc_extract failed rc={rc}")
            return kps[: n.value].copy(), desc[: n.value].copy()

    def level_size(self, level: int):
        w, h = C.c_int(0), C.c_int(0)
        lib().orc_level_size(self._h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def _img(self, fn, level):
        w, h = self.level_size(level)
        p = getattr(lib(), fn)(self._h, level)
        if not p:
            return None
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(h, w)).copy()

    def level(self, level: int): return self._img("orc_level_ptr", level)
    def blurred(self, level: int): return self._img("orc_blurred_ptr", level)

    def level_candidates(self, level: int):
        cap = 1 << 16
        while True:
            x = np.empty(cap, np.int32); y = np.empty(cap, np.int32); r = np.empty(cap, np.int32)
            n = lib().orc_level_candidates(self._h, level, cap, _ptr(x), _ptr(y), _ptr(r))
            if n <= cap:
                return x[:n].copy(), y[:n].copy(), r[:n].copy()
            cap = n

    def level_keypoints(self, level: int):
        cap = 1 << 16
        x = np.empty(cap, np.int32); y = np.empty(cap, np.int32); r = np.empty(cap, np.int32)
        a = np.empty(cap, np.float32)
        n = lib().orc_level_keypoints(self._h, level, cap, _ptr(x), _ptr(y), _ptr(r), _ptr(a))
        return x[:n].copy(), y[:n].copy(), r[:n].copy(), a[:n].copy()

    def tie_sensitive(self, level: int) -> bool:
        return bool(lib().orc_level_tie_sensitive(self._h, level))

    def tie_flags(self, level: int) -> int:
        """bit 0: the retained set depends on the heap-address tie of ORBextractor.cc:684; bit 1: the order does (orb_oracle.h)."""
        return int(lib().orc_level_tie_sensitive(self._h, level))


def stereo_match(eL: OracleExtractor, eR: OracleExtractor, kL, dL, kR, dR, bf: float, maxD: float, debug=False):
    """Frame::ComputeStereoMatches on the pyramids currently held by eL / eR."""
    kL = np.ascontiguousarray(kL, KEYPOINT_DTYPE); kR = np.ascontiguousarray(kR, KEYPOINT_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    nL, nR = len(kL), len(kR)
    ur = np.empty(max(nL, 1), np.float32); dp = np.empty(max(nL, 1), np.float32)
    bd = np.empty(max(nL, 1), np.int32); bi = np.empty(max(nL, 1), np.int32); bs = np.empty(max(nL, 1), np.int32)
    rc = lib().orc_stereo_match(eL._h, eR._h, _ptr(kL), nL, _ptr(dL), _ptr(kR), nR, _ptr(dR),
                                C.c_float(bf), C.c_float(maxD), _ptr(ur), _ptr(dp), _ptr(bd), _ptr(bi), _ptr(bs))
    if rc != 0:
        raise RuntimeError(f"orc_stereo_match rc={rc}")
    if debug:
        return ur[:nL], dp[:nL], bd[:nL], bi[:nL], bs[:nL]
    return ur[:nL], dp[:nL]


def stereo_pair(params, imL, imR, bf, maxD):
    """Whole unit of work: extract both views + stereo match (reference Frame.cc:144-173)."""
    eL = OracleExtractor(*params); eR = OracleExtractor(*params)
    kL, dL = eL.extract(imL); kR, dR = eR.extract(imR)
    ur, dp = stereo_match(eL, eR, kL, dL, kR, dR, bf, maxD)
    return kL, dL, kR, dR, ur, dp


# ---------------------------------------------------------------------------------------------
# Frame::ComputeBoW = DBoW2 transform (reference src/Frame.cc:2040-2047, Thirdparty/DBoW2)
# ---------------------------------------------------------------------------------------------
class _OrcVocab(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("k", C.c_int32), ("L", C.c_int32), ("scoring", C.c_int32), ("weighting", C.c_int32),
                ("desc", C.c_void_p), ("child_start", C.c_void_p), ("child", C.c_void_p), ("word_id", C.c_void_p),
                ("weight", C.c_void_p)]


def _orc_vocab(voc):
    """`voc`: pointslot_b200.vocab.Vocabulary (plain arrays; the container is shared with the product's host side)."""
    keep = [np.ascontiguousarray(voc.desc, np.uint8), np.ascontiguousarray(voc.child_start, np.int32),
            np.ascontiguousarray(voc.child, np.int32), np.ascontiguousarray(voc.word_id, np.int32),
            np.ascontiguousarray(voc.weight, np.float64)]
    s = _OrcVocab(voc.n_nodes, voc.k, voc.L, voc.scoring, voc.weighting, *(a.ctypes.data for a in keep))
    return s, keep


def bow_transform(voc, desc: np.ndarray, levelsup: int = 4):
    """-> (leaf node id, node id at level L - levelsup) per descriptor."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(desc)
    s, keep = _orc_vocab(voc)
    leaf = np.zeros(n, np.int32); nid = np.zeros(n, np.int32)
    lib().orc_bow_transform(C.byref(s), _ptr(desc), n, levelsup, _ptr(leaf), _ptr(nid))
    return leaf, nid


def bow_vectors(voc, leaf: np.ndarray, nid: np.ndarray):
    """-> (word ids ascending, values) = mBowVec and (node ids ascending, start offsets, feature indices) = mFeatVec."""
    n = len(leaf)
    s, keep = _orc_vocab(voc)
    leaf = np.ascontiguousarray(leaf, np.int32); nid = np.ascontiguousarray(nid, np.int32)
    m = max(n, 1)
    bid = np.zeros(m, np.int32); bval = np.zeros(m, np.float64)
    fn = np.zeros(m, np.int32); fs = np.zeros(m + 1, np.int32); ff = np.zeros(m, np.int32)
    nf = C.c_int(0)
    nw = lib().orc_bow_vectors(C.byref(s), _ptr(leaf), _ptr(nid), n, _ptr(bid), _ptr(bval), _ptr(fn), _ptr(fs), _ptr(ff), C.byref(nf))
    return bid[:nw].copy(), bval[:nw].copy(), fn[:nf.value].copy(), fs[:nf.value + 1].copy(), ff[:fs[nf.value]].copy()
