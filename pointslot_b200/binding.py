"""ctypes binding of include/pslot_orb.h — the host-side mirror used by tests and bench.py.

Mirrors the reference's interface for this path (ORBextractor / Frame stereo constructor):
  OrbExtractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)   ~ ORBextractor.h:45-111
  OrbExtractor.extract(image) -> (keypoints, descriptors)               ~ operator() (ORBextractor.cc:1043)
  stereo_match(extL, extR, kL, dL, kR, dR, bf, maxD) -> (uRight, depth) ~ Frame::ComputeStereoMatches
  StereoBatch: many independent pairs per call on one GPU stream.
There is no fallback: if libpslot_orb.so is missing or no CUDA device exists, calls raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libpslot_orb.so"

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)

# every symbol include/pslot_orb.h declares (tests check the .so exports exactly these)
ABI_SYMBOLS = [
    "pslot_last_error", "pslot_abi_version", "pslot_device_count", "pslot_host_alloc", "pslot_host_free",
    "pslot_orb_create", "pslot_orb_destroy", "pslot_orb_levels", "pslot_orb_tables", "pslot_orb_max_keypoints",
    "pslot_orb_extract", "pslot_orb_extract_color", "pslot_orb_level", "pslot_stereo_match", "pslot_hamming256",
    "pslot_batch_create", "pslot_batch_destroy", "pslot_batch_set_copy_stream", "pslot_batch_max_keypoints", "pslot_batch_stream",
    "pslot_batch_submit", "pslot_batch_submit_color", "pslot_batch_wait", "pslot_batch_kernel_launches", "pslot_batch_device_out",
    "pslot_batch_set_profiling", "pslot_batch_stage_times", "pslot_batch_stats", "pslot_batch_stereo_counters",
    "pslot_assign_grid", "pslot_grid_token", "pslot_features_in_area", "pslot_batch_assign_grid",
    "pslot_search_by_projection", "pslot_search_by_projection_obj", "pslot_search_by_projection_last", "pslot_search_by_projection_kf", "pslot_debug_projection_sweeps",
    "pslot_cvorb_create", "pslot_cvorb_destroy", "pslot_cvorb_detect_and_compute", "pslot_cvorb_detect_and_compute_batch", "pslot_cvorb_detect_and_compute_resident", "pslot_cvorb_stereo_match_resident", "pslot_cvorb_last_timing", "pslot_cvorb_debug",
    "pslot_assign_features", "pslot_batch_assign_features", "pslot_orb_extract_begin", "pslot_orb_extract_end", "pslot_orb_extract_token", "pslot_stereo_match_resident",
    "pslot_vocab_create", "pslot_vocab_destroy", "pslot_bow_transform", "pslot_bow_vectors",
    "pslot_orb_debug_candidates", "pslot_orb_debug_level_keypoints", "pslot_stereo_debug", "pslot_debug_steer",
]


class PslotError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pslot error {code}: {msg}")
        self.code = code


class GridSpec(C.Structure):
    """pslot_grid_spec: mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv, FRAME_GRID_COLS / ROWS."""
    _fields_ = [("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float), ("inv_h", C.c_float),
                ("cols", C.c_int32), ("rows", C.c_int32)]

    @classmethod
    def for_image(cls, width: int, height: int, cols: int = 64, rows: int = 48):
        """Rectified input (k1 == 0): bounds = the image (Frame.cc:2134-2139), inverse cell sizes as :184-185."""
        w, h = np.float32(width), np.float32(height)
        return cls(0.0, 0.0, float(np.float32(cols) / (w - np.float32(0))), float(np.float32(rows) / (h - np.float32(0))),
                   cols, rows)


MAP_POINT_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"),
                            ("level", "<i4"), ("flags", "<i4")])  # pslot_map_point
PROJ_POINT_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"),
                             ("max_level", "<i4"), ("angle", "<f4"), ("flags", "<i4")])  # pslot_proj_point
AREA_QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4")])


class BatchOut(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("kps", C.c_void_p), ("desc", C.c_void_p),
                ("u_right", C.c_void_p), ("depth", C.c_void_p)]


_lib = None


def lib():
    """Load libpslot_orb.so (raises if it has not been built — there is no CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m pointslot_b200.build` (needs nvcc). "
                           "pointslot_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    L.pslot_last_error.restype = C.c_char_p
    L.pslot_orb_create.argtypes = [C.POINTER(vp), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pslot_orb_destroy.argtypes = [vp]
    L.pslot_orb_destroy.restype = None
    L.pslot_orb_levels.argtypes = [vp]
    L.pslot_orb_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pslot_orb_max_keypoints.argtypes = [vp]
    L.pslot_orb_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_ssize_t, vp, vp, C.c_int, ip]
    L.pslot_orb_extract_color.argtypes = [vp, vp, C.c_int, C.c_int, C.c_ssize_t, C.c_int, C.c_int, vp, vp, C.c_int, ip]
    L.pslot_orb_level.argtypes = [vp, C.c_int, vp, C.c_ssize_t, ip, ip]
    L.pslot_stereo_match.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_float, C.c_float, vp, vp]
    L.pslot_hamming256.argtypes = [vp, vp]
    L.pslot_assign_grid.argtypes = [vp, vp, C.c_int, C.POINTER(GridSpec), vp, vp, C.POINTER(C.c_int)]
    L.pslot_grid_token.argtypes = [vp]
    L.pslot_grid_token.restype = C.c_uint64
    L.pslot_features_in_area.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    L.pslot_batch_assign_grid.argtypes = [vp, C.POINTER(GridSpec), vp, vp]
    L.pslot_search_by_projection.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, C.c_float,
                                             C.c_float, vp, C.POINTER(C.c_int)]
    L.pslot_search_by_projection_obj.argtypes = L.pslot_search_by_projection.argtypes
    L.pslot_search_by_projection_last.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, vp,
                                                  C.POINTER(C.c_int)]
    L.pslot_search_by_projection_kf.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(C.c_int)]
    L.pslot_debug_projection_sweeps.argtypes = [vp]
    L.pslot_batch_create.argtypes = [C.POINTER(vp), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, vp]
    L.pslot_batch_destroy.argtypes = [vp]
    L.pslot_batch_destroy.restype = None
    L.pslot_batch_set_copy_stream.argtypes = [vp, vp]
    L.pslot_batch_max_keypoints.argtypes = [vp]
    L.pslot_batch_stream.argtypes = [vp]
    L.pslot_batch_stream.restype = vp
    L.pslot_batch_submit.argtypes = [vp, vp, vp, C.c_ssize_t, C.c_ssize_t, C.c_int, C.c_int, C.c_float, C.c_float,
                                     C.POINTER(BatchOut)]
    L.pslot_batch_submit_color.argtypes = [vp, vp, vp, C.c_ssize_t, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_float,
                                           C.c_float, C.POINTER(BatchOut)]
    L.pslot_batch_wait.argtypes = [vp]
    L.pslot_batch_kernel_launches.argtypes = [vp]
    L.pslot_batch_device_out.argtypes = [vp, C.POINTER(BatchOut)]
    L.pslot_batch_set_profiling.argtypes = [vp, C.c_int]
    L.pslot_batch_stage_times.argtypes = [vp, vp]
    L.pslot_batch_stats.argtypes = [vp, vp]
    L.pslot_cvorb_create.argtypes = [C.POINTER(vp), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pslot_cvorb_destroy.argtypes = [vp]
    L.pslot_cvorb_destroy.restype = None
    L.pslot_cvorb_detect_and_compute.argtypes = [vp, vp, C.c_int, C.c_int, C.c_ssize_t, vp, C.c_ssize_t, vp, vp, C.c_int,
                                                 C.POINTER(C.c_int)]
    L.pslot_cvorb_detect_and_compute_batch.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_ssize_t, vp, C.c_ssize_t, vp, vp,
                                                       C.c_int, vp]
    L.pslot_cvorb_detect_and_compute_resident.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_ssize_t, vp, vp, C.c_int, vp]
    L.pslot_cvorb_stereo_match_resident.argtypes = [vp, vp, C.c_float, C.c_float, vp, vp, C.c_int]
    L.pslot_cvorb_last_timing.argtypes = [vp, vp]
    L.pslot_orb_extract_begin.argtypes = [vp, vp, C.c_int, C.c_int, C.c_ssize_t]
    L.pslot_orb_extract_end.argtypes = [vp, vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.pslot_orb_extract_token.argtypes = [vp]
    L.pslot_orb_extract_token.restype = C.c_uint64
    L.pslot_stereo_match_resident.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_float, C.c_float, vp, vp, C.c_int, C.POINTER(C.c_int)]
    L.pslot_assign_features.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_ssize_t, vp, C.c_int, vp, C.c_float, vp, vp, vp, vp]
    L.pslot_batch_assign_features.argtypes = [vp, vp, C.c_ssize_t, C.c_ssize_t, vp, C.c_int, vp, vp, vp]
    L.pslot_cvorb_debug.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.pslot_vocab_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int]
    L.pslot_vocab_destroy.argtypes = [vp]
    L.pslot_vocab_destroy.restype = None
    L.pslot_bow_transform.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp]
    L.pslot_bow_vectors.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.POINTER(C.c_int), vp, vp, vp, C.POINTER(C.c_int)]
    L.pslot_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]
    L.pslot_host_free.argtypes = [vp]
    L.pslot_batch_stereo_counters.argtypes = [vp, vp]
    L.pslot_orb_debug_candidates.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, ip]
    L.pslot_orb_debug_level_keypoints.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, ip]
    L.pslot_stereo_debug.argtypes = [vp, C.c_int, vp, vp, vp]
    L.pslot_debug_steer.argtypes = [C.c_int, vp, C.c_int, vp, vp]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise PslotError(rc, lib().pslot_last_error().decode(errors="replace"))
    return rc


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return _check(lib().pslot_device_count())


def hamming256(a, b) -> int:
    """ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:2704-2720)."""
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    assert a.size == 32 and b.size == 32
    return lib().pslot_hamming256(_p(a), _p(b))


def debug_steer(angle_deg: np.ndarray, device: int = 0):
    """(a, b) = ((float)cos, (float)sin) of angle*pi/180 as the descriptor kernel computes them."""
    ang = np.ascontiguousarray(angle_deg, np.float32)
    a = np.empty_like(ang); b = np.empty_like(ang)
    _check(lib().pslot_debug_steer(device, _p(ang), ang.size, _p(a), _p(b)))
    return a, b


def max_disparity(bf: float, fx: float) -> float:
    """The value the reference intends for maxD: mb = mbf/fx, maxD = mbf/mb in float32."""
    mb = np.float32(bf) / np.float32(fx)
    return float(np.float32(bf) / mb)


class OrbExtractor:
    """Drop-in mirror of ORB_SLAM2::ORBextractor over the C ABI."""

    def __init__(self, nfeatures: int, scale_factor: float, nlevels: int, ini_th: int, min_th: int, device: int = 0):
        h = C.c_void_p()
        _check(lib().pslot_orb_create(C.byref(h), nfeatures, scale_factor, nlevels, ini_th, min_th, device))
        self._h = h
        self.nlevels = nlevels

    def close(self):
        if getattr(self, "_h", None):
            lib().pslot_orb_destroy(self._h)
            self._h = None

    __del__ = close

    def tables(self):
        n = self.nlevels
        out = [np.zeros(n, np.float32) for _ in range(4)] + [np.zeros(n, np.int32)]
        _check(lib().pslot_orb_tables(self._h, *[_p(a) for a in out]))
        return dict(zip(("scale", "inv_scale", "sigma2", "inv_sigma2", "features_per_level"), out))

    def extract(self, img: np.ndarray, rgb: bool = False):
        """-> (keypoints [n] KEYPOINT_DTYPE, descriptors [n,32] uint8).  img: gray [H,W], or interleaved colour
        [H,W,3|4] (converted on the device like Tracking::GrabImageStereo's cvtColor; rgb=True for RGB(A) order)."""
        assert img.dtype == np.uint8 and img.ndim in (2, 3)
        channels = 1 if img.ndim == 2 else img.shape[2]
        assert img.size == 0 or img.strides[-1] == 1 and (img.ndim == 2 or img.strides[1] == channels)
        cap = _check(lib().pslot_orb_max_keypoints(self._h))
        h, w = img.shape[:2]
        for _ in range(2):
            kps = np.zeros(cap, KEYPOINT_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = C.c_int(0)
            if channels == 1:
                rc = lib().pslot_orb_extract(self._h, _p(img), w, h, img.strides[0] if img.size else 0, _p(kps),
                                             _p(desc), cap, C.byref(n))
            else:
                rc = lib().pslot_orb_extract_color(self._h, _p(img), w, h, img.strides[0], channels, 1 if rgb else 0,
                                                   _p(kps), _p(desc), cap, C.byref(n))
            if rc == -2:  # PSLOT_ERR_CAPACITY
                cap = n.value
                continue
            _check(rc)
            return kps[: n.value].copy(), desc[: n.value].copy()
        raise PslotError(-2, "capacity retry failed")

    def extract_begin(self, img: np.ndarray):
        """First half of extract() for a gray image: upload + all kernels + the result download are enqueued; pair with
        extract_end().  Lets one host thread keep several extractors (the two views) in flight."""
        assert img.dtype == np.uint8 and img.ndim == 2 and (img.size == 0 or img.strides[1] == 1)
        h, w = img.shape
        _check(lib().pslot_orb_extract_begin(self._h, _p(img), w, h, img.strides[0] if img.size else 0))

    def extract_end(self):
        cap = _check(lib().pslot_orb_max_keypoints(self._h))
        for _ in range(2):
            kps = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
            n = C.c_int(0)
            rc = lib().pslot_orb_extract_end(self._h, _p(kps), _p(desc), cap, C.byref(n))
            if rc == -2:  # PSLOT_ERR_CAPACITY: the results wait in the staging buffer
                cap = n.value
                continue
            _check(rc)
            return kps[: n.value].copy(), desc[: n.value].copy()
        raise PslotError(-2, "capacity retry failed")

    def level(self, level: int) -> np.ndarray:
        """mvImagePyramid[level] of the last extraction (borderless)."""
        w, h = C.c_int(0), C.c_int(0)
        _check(lib().pslot_orb_level(self._h, level, None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        _check(lib().pslot_orb_level(self._h, level, _p(out), out.strides[0], C.byref(w), C.byref(h)))
        return out

    def debug_candidates(self, level: int):
        cap = 1 << 18
        x = np.empty(cap, np.int32); y = np.empty(cap, np.int32); r = np.empty(cap, np.int32)
        n = C.c_int(0)
        _check(lib().pslot_orb_debug_candidates(self._h, level, cap, _p(x), _p(y), _p(r), C.byref(n)))
        return x[: n.value].copy(), y[: n.value].copy(), r[: n.value].copy()

    def debug_level_keypoints(self, level: int):
        cap = 1 << 16
        x = np.empty(cap, np.int32); y = np.empty(cap, np.int32); r = np.empty(cap, np.int32)
        a = np.empty(cap, np.float32)
        n = C.c_int(0)
        _check(lib().pslot_orb_debug_level_keypoints(self._h, level, cap, _p(x), _p(y), _p(r), _p(a), C.byref(n)))
        return x[: n.value].copy(), y[: n.value].copy(), r[: n.value].copy(), a[: n.value].copy()


def assign_grid(ext: OrbExtractor, kps_un, spec: GridSpec):
    """Frame::AssignFeaturesToGrid (reference src/Frame.cc:1636-1655) -> (cell_start [cols*rows+1], indices
    [n_in_grid]); mGrid[i][j] = indices[cell_start[i*rows+j] : cell_start[i*rows+j+1]]."""
    kps_un = np.ascontiguousarray(kps_un, KEYPOINT_DTYPE)
    n = len(kps_un)
    cs = np.zeros(spec.cols * spec.rows + 1, np.int32); idx = np.zeros(max(n, 1), np.int32)
    tot = C.c_int(0)
    _check(lib().pslot_assign_grid(ext._h, _p(kps_un), n, C.byref(spec), _p(cs), _p(idx), C.byref(tot)))
    return cs, idx[:tot.value].copy()


def assign_features(ext: OrbExtractor, kps, mask, object_ids, depth=None, depth_limit=0.0):
    """Frame::AssignFeatures (reference src/Frame.cc:759-977): classify the features by the instance mask and partition them
    in the reference's push_back order -> (cls [n], groups, effective) with groups[0] = indices of the background features,
    groups[1 + j] = indices of detection j's features; effective[j] (when `depth` is given) = features of detection j with
    0 < depth < depth_limit.  `mask` = None reuses the mask of the previous call on `ext`."""
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE)
    ids = np.ascontiguousarray(object_ids, np.int32)
    n, J = len(kps), len(ids)
    m = mask
    if m is not None and not (m.dtype == np.uint8 and m.ndim == 2 and m.strides[1] == 1):
        m = np.ascontiguousarray(m, np.uint8)
    H, W = (getattr(ext, "_mask_shape", (1, 1)) if m is None else m.shape)
    pitch = W if m is None else m.strides[0]
    d = None if depth is None else np.ascontiguousarray(depth, np.float32)
    cls = np.zeros(max(n, 1), np.int32); order = np.zeros(max(n, 1), np.int32)
    cs = np.zeros(J + 2, np.int32); eff = np.zeros(max(J, 1), np.int32)
    _check(lib().pslot_assign_features(ext._h, _p(kps), n, None if m is None else _p(m), W, H, pitch, _p(ids), J,
                                       None if d is None else _p(d), C.c_float(depth_limit), _p(cls), _p(order), _p(cs),
                                       _p(eff) if d is not None else None))
    if m is not None:
        ext._mask_shape = (H, W)
    groups = [order[cs[c]:cs[c + 1]].copy() for c in range(J + 1)]
    return cls[:n], groups, (eff[:J] if d is not None else None)


class CvOrb:
    """cv::ORB::create(nfeatures, scaleFactor, nlevels, edgeThreshold) over the C ABI (reference src/Frame.cc:2623-2627,
    the dense object features): HARRIS_SCORE, patch 31, WTA_K 2; detect_and_compute -> keypoints ordered by
    (octave, y, x) and their descriptors."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, edge_threshold=19, fast_threshold=20, device=0):
        h = C.c_void_p()
        _check(lib().pslot_cvorb_create(C.byref(h), nfeatures, scale_factor, nlevels, edge_threshold, fast_threshold, device))
        self._h = h
        self.cap = nfeatures + 64 * nlevels + 256

    def close(self):
        if getattr(self, "_h", None):
            lib().pslot_cvorb_destroy(self._h)
            self._h = None

    __del__ = close

    def detect_and_compute(self, img: np.ndarray, mask=None):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        cap = self.cap
        while True:
            kps = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
            n = C.c_int(0)
            rc = lib().pslot_cvorb_detect_and_compute(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                                      None if m is None else _p(m), 0 if m is None else m.strides[0],
                                                      _p(kps), _p(desc), cap, C.byref(n))
            if rc == -2 and n.value > cap:
                cap = n.value
                continue
            _check(rc)
            return kps[:n.value].copy(), desc[:n.value].copy()


def _cvorb_batch(self, imgs, masks=None, copy=True):
    """pslot_cvorb_detect_and_compute_batch: `imgs` = equally sized u8 images (e.g. the left and right view), `masks` = None
    or one entry (array or None) per image -> [(keypoints, descriptors)] per image, each exactly the single-image result.
    copy=False returns views into the context's result arrays, valid until the next call (no per-image copies)."""
    n = len(imgs)
    if n == 0:
        return []
    def rows(a):  # a view with contiguous rows travels as it is (its pitch is passed on), anything else is copied
        a = np.asarray(a)
        return a if a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1 and a.strides[0] >= a.shape[1] else np.ascontiguousarray(a, np.uint8)
    ims = [rows(i) for i in imgs]
    H, W = ims[0].shape
    assert all(i.shape == (H, W) for i in ims)
    if len({i.strides[0] for i in ims}) > 1:
        ims = [np.ascontiguousarray(i) for i in ims]
    pitch = ims[0].strides[0]
    ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in ims])
    ms, mptrs, mpitch = None, None, W
    if masks is not None and any(m is not None for m in masks):
        ms = [None if m is None else rows(m) for m in masks]
        assert len(ms) == n and all(m is None or m.shape == (H, W) for m in ms)
        if len({m.strides[0] for m in ms if m is not None}) > 1:
            ms = [None if m is None else np.ascontiguousarray(m) for m in ms]
        mpitch = next(m.strides[0] for m in ms if m is not None)
        mptrs = (C.c_void_p * n)(*[None if m is None else m.ctypes.data for m in ms])
    cap = self.cap
    while True:
        # the result arrays are kept between calls (fresh pages for every call cost more than the device work of a batch)
        buf = getattr(self, "_batch_out", None)
        if buf is None or buf[0].shape[0] < n or buf[0].shape[1] != cap:
            buf = (np.zeros((n, cap), KEYPOINT_DTYPE), np.zeros((n, cap, 32), np.uint8))
            self._batch_out = buf
        kps, desc = buf
        counts = np.zeros(n, np.int32)
        rc = lib().pslot_cvorb_detect_and_compute_batch(self._h, n, ptrs, W, H, pitch, mptrs, mpitch, _p(kps), _p(desc), cap, _p(counts))
        if rc == -2 and counts.max() > cap:
            cap = int(counts.max())
            continue
        _check(rc)
        if not copy:
            return [(kps[s, :counts[s]], desc[s, :counts[s]]) for s in range(n)]
        return [(kps[s, :counts[s]].copy(), desc[s, :counts[s]].copy()) for s in range(n)]


def _cvorb_resident(self, batch, first_image, n_images, masks=None, copy=True):
    """pslot_cvorb_detect_and_compute_resident: cv::ORB on images first_image .. of the last submit of `batch` (a StereoBatch;
    left views first, then right views), read where they lie on the device; `masks` as in detect_and_compute_batch."""
    n = n_images
    H, W = batch.H, batch.W
    ms, mptrs, mpitch = None, None, W
    if masks is not None and any(m is not None for m in masks):
        ms = [None if m is None else np.ascontiguousarray(m, np.uint8) for m in masks]
        assert len(ms) == n and all(m is None or m.shape == (H, W) for m in ms)
        mptrs = (C.c_void_p * n)(*[None if m is None else m.ctypes.data for m in ms])
    cap = self.cap
    while True:
        buf = getattr(self, "_batch_out", None)
        if buf is None or buf[0].shape[0] < n or buf[0].shape[1] != cap:
            buf = (np.zeros((n, cap), KEYPOINT_DTYPE), np.zeros((n, cap, 32), np.uint8))
            self._batch_out = buf
        kps, desc = buf
        counts = np.zeros(max(n, 1), np.int32)
        rc = lib().pslot_cvorb_detect_and_compute_resident(self._h, batch._h, first_image, n, mptrs, mpitch, _p(kps), _p(desc), cap,
                                                           _p(counts))
        if rc == -2 and counts.max() > cap:
            cap = int(counts.max())
            continue
        _check(rc)
        if not copy:
            return [(kps[s, :counts[s]], desc[s, :counts[s]]) for s in range(n)]
        return [(kps[s, :counts[s]].copy(), desc[s, :counts[s]].copy()) for s in range(n)]


CvOrb.detect_and_compute_resident = _cvorb_resident


def _cvorb_stereo_resident(self, batch, npairs, bf, max_d):
    """pslot_cvorb_stereo_match_resident: ComputeObjStereoMatches for every pair of `batch` on the cv::ORB features the last
    detect_and_compute_resident(batch, 0, 2 * npairs) left on the device -> (u_right [npairs, cap], depth [npairs, cap])."""
    ur = np.full((npairs, self.cap), -1, np.float32); dp = np.full((npairs, self.cap), -1, np.float32)
    _check(lib().pslot_cvorb_stereo_match_resident(self._h, batch._h, C.c_float(bf), C.c_float(max_d), _p(ur), _p(dp), self.cap))
    return ur, dp


CvOrb.stereo_match_resident = _cvorb_stereo_resident


def _cvorb_timing(self):
    """Host-clock phases of the last call in ms: uploads enqueued, kernels enqueued, device wait, copy-out."""
    t = np.zeros(4, np.float64)
    _check(lib().pslot_cvorb_last_timing(self._h, _p(t)))
    return t


CvOrb.last_timing = _cvorb_timing
CvOrb.detect_and_compute_batch = _cvorb_batch


def _cvorb_stage(self, what: int, level: int):
    """Stage accessor (pslot_cvorb_debug): 0 pyramid level, 1 FAST score map -> (h, w) u8; 2 candidates, 3 kept after
    retainBest(2 * quota) -> (x, y, score[, harris]) arrays."""
    cap = 1 << 24
    buf = np.zeros(cap, np.uint8); aux = np.zeros(cap // 4, np.float32)
    w, h, n = C.c_int(0), C.c_int(0), C.c_int(0)
    _check(lib().pslot_cvorb_debug(self._h, what, level, _p(buf), _p(aux), cap, C.byref(w), C.byref(h), C.byref(n)))
    if what < 2:
        return buf[:w.value * h.value].reshape(h.value, w.value).copy()
    c = buf[:4 * n.value].view(np.uint32)
    x, y, s = ((c >> 8) & 0xFFF).astype(np.int32), (c >> 20).astype(np.int32), (c & 0xFF).astype(np.int32)
    return (x, y, s) if what == 2 else (x, y, s, aux[:n.value].copy())


CvOrb.stage = _cvorb_stage


class OrbVocabulary:
    """The ORB vocabulary on the device (pslot_vocab_*): `voc` is a pointslot_b200.vocab.Vocabulary (what
    ORBVocabulary::loadFromTextFile builds).  compute_bow = Frame::ComputeBoW (reference src/Frame.cc:2040-2047)."""

    def __init__(self, voc, device=0):
        h = C.c_void_p()
        self._keep = [np.ascontiguousarray(voc.desc, np.uint8), np.ascontiguousarray(voc.child_start, np.int32),
                      np.ascontiguousarray(voc.child, np.int32), np.ascontiguousarray(voc.word_id, np.int32),
                      np.ascontiguousarray(voc.weight, np.float64)]
        _check(lib().pslot_vocab_create(C.byref(h), voc.n_nodes, voc.k, voc.L, voc.scoring, voc.weighting,
                                        *(_p(a) for a in self._keep), device))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().pslot_vocab_destroy(self._h)
            self._h = None

    __del__ = close

    def transform(self, desc=None, levelsup=4, extractor: "OrbExtractor" = None, n=None):
        """-> (leaf node, node at level L - levelsup) per descriptor; `desc` None = the descriptors the extractor's last
        extract() left on the device (n of them)."""
        if desc is not None:
            desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
            n = len(desc)
        leaf = np.zeros(max(n, 1), np.int32); nid = np.zeros(max(n, 1), np.int32)
        _check(lib().pslot_bow_transform(self._h, extractor._h if extractor is not None else None,
                                         None if desc is None else _p(desc), n, levelsup, _p(leaf), _p(nid)))
        return leaf[:n], nid[:n]

    def vectors(self, leaf, nid):
        n = len(leaf)
        leaf = np.ascontiguousarray(leaf, np.int32); nid = np.ascontiguousarray(nid, np.int32)
        m = max(n, 1)
        bid = np.zeros(m, np.int32); bval = np.zeros(m, np.float64)
        fn = np.zeros(m, np.int32); fs = np.zeros(m + 1, np.int32); ff = np.zeros(m, np.int32)
        nw, nf = C.c_int(0), C.c_int(0)
        _check(lib().pslot_bow_vectors(self._h, _p(leaf), _p(nid), n, _p(bid), _p(bval), C.byref(nw), _p(fn), _p(fs), _p(ff), C.byref(nf)))
        return bid[:nw.value].copy(), bval[:nw.value].copy(), fn[:nf.value].copy(), fs[:nf.value + 1].copy(), ff[:fs[nf.value]].copy()

    def compute_bow(self, desc=None, levelsup=4, extractor=None, n=None):
        return self.vectors(*self.transform(desc, levelsup, extractor, n))


class PinnedBuffer:
    """Page-locked host memory from pslot_host_alloc, viewed as a numpy array (freed with the object)."""

    MODES = {"default": 0, "wc": 1, "thp": 2}

    def __init__(self, shape, dtype=np.uint8, write_combined: bool = False, mode: str = ""):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        m = self.MODES[mode] if mode else (1 if write_combined else 0)
        _check(lib().pslot_host_alloc(C.byref(p), self.nbytes, m))
        self._p = p
        self.ptr = p.value or 0
        buf = (C.c_ubyte * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if getattr(self, "_p", None) is not None and self._p.value:
            self.array = None
            lib().pslot_host_free(self._p)
            self._p = None

    __del__ = close


def grid_token(ext: OrbExtractor) -> int:
    """Identity of the keypoint set currently gridded on `ext` (0 = none); changes with every assign_grid."""
    return int(lib().pslot_grid_token(ext._h))


def features_in_area(ext: OrbExtractor, queries, cap: int):
    """Frame::GetFeaturesInArea (reference src/Frame.cc:1808-1861) for an array of AREA_QUERY_DTYPE queries against
    the grid of the last assign_grid on `ext` -> list of index arrays (reference visiting order)."""
    queries = np.ascontiguousarray(queries, AREA_QUERY_DTYPE)
    nq = len(queries)
    counts = np.zeros(max(nq, 1), np.int32); out = np.zeros((max(nq, 1), max(cap, 1)), np.int32)
    _check(lib().pslot_features_in_area(ext._h, _p(queries), nq, cap, _p(counts), _p(out)))
    if nq and counts[:nq].max() > cap:
        raise PslotError(-2, f"area query found {int(counts[:nq].max())} features, cap is {cap}")
    return [out[q, :counts[q]].copy() for q in range(nq)]


def search_by_projection(ext: OrbExtractor, mps, mp_desc, frame_desc, u_right, occupied, scale_factors, th: float,
                         nn_ratio: float):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (reference src/ORBmatcher.cc:68-155) against the frame whose
    undistorted keypoints were last passed to assign_grid on `ext` -> (match_of_mp [n_mp] feature index or -1,
    nmatches, device sweeps)."""
    mps = np.ascontiguousarray(mps, MAP_POINT_DTYPE); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    frame_desc = np.ascontiguousarray(frame_desc, np.uint8); u_right = np.ascontiguousarray(u_right, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    out = np.full(max(len(mps), 1), -1, np.int32)
    nm = C.c_int(0)
    _check(lib().pslot_search_by_projection(ext._h, _p(mps), _p(mp_desc), len(mps), _p(frame_desc), _p(u_right),
                                            None if occ is None else _p(occ), len(u_right), _p(sf), len(sf), th, nn_ratio,
                                            _p(out), C.byref(nm)))
    return out[:len(mps)].copy(), nm.value, lib().pslot_debug_projection_sweeps(ext._h)



def search_by_projection_obj(ext: OrbExtractor, mps, mp_desc, frame_desc, u_right, occupied, scale_factors, th: float,
                         nn_ratio: float):
    """ORBmatcher::SearchByProjection(F, nOrder, vpMapObjectPoints, th) (reference src/ORBmatcher.cc:157-246) against the
    object feature set last passed to assign_grid on `ext` -> (match_of_mp [n_mp] feature index or -1,
    nmatches, device sweeps)."""
    mps = np.ascontiguousarray(mps, MAP_POINT_DTYPE); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    frame_desc = np.ascontiguousarray(frame_desc, np.uint8); u_right = np.ascontiguousarray(u_right, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    out = np.full(max(len(mps), 1), -1, np.int32)
    nm = C.c_int(0)
    _check(lib().pslot_search_by_projection_obj(ext._h, _p(mps), _p(mp_desc), len(mps), _p(frame_desc), _p(u_right),
                                            None if occ is None else _p(occ), len(u_right), _p(sf), len(sf), th, nn_ratio,
                                            _p(out), C.byref(nm)))
    return out[:len(mps)].copy(), nm.value, lib().pslot_debug_projection_sweeps(ext._h)

def search_by_projection_last(ext: OrbExtractor, pts, pt_desc, frame_desc, u_right, occupied, check_orientation=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) after the projection (reference
    src/ORBmatcher.cc:1671-1756) -> (match_of_pt, rejected, nmatches, device sweeps)."""
    pts = np.ascontiguousarray(pts, PROJ_POINT_DTYPE); pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
    frame_desc = np.ascontiguousarray(frame_desc, np.uint8); u_right = np.ascontiguousarray(u_right, np.float32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    out = np.full(max(len(pts), 1), -1, np.int32); rej = np.zeros(max(len(pts), 1), np.uint8)
    nm = C.c_int(0)
    _check(lib().pslot_search_by_projection_last(ext._h, _p(pts), _p(pt_desc), len(pts), _p(frame_desc), _p(u_right),
                                                 None if occ is None else _p(occ), len(u_right), int(check_orientation),
                                                 _p(out), _p(rej), C.byref(nm)))
    return out[:len(pts)].copy(), rej[:len(pts)].copy(), nm.value, lib().pslot_debug_projection_sweeps(ext._h)


def search_by_projection_kf(ext: OrbExtractor, pts, pt_desc, frame_desc, occupied, orb_dist, check_orientation=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) after the projection (reference
    src/ORBmatcher.cc:2583-2654) -> (match_of_pt, rejected, nmatches)."""
    pts = np.ascontiguousarray(pts, PROJ_POINT_DTYPE); pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
    frame_desc = np.ascontiguousarray(frame_desc, np.uint8).reshape(-1, 32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    out = np.full(max(len(pts), 1), -1, np.int32); rej = np.zeros(max(len(pts), 1), np.uint8)
    nm = C.c_int(0)
    _check(lib().pslot_search_by_projection_kf(ext._h, _p(pts), _p(pt_desc), len(pts), _p(frame_desc),
                                               None if occ is None else _p(occ), len(frame_desc), int(orb_dist),
                                               int(check_orientation), _p(out), _p(rej), C.byref(nm)))
    return out[:len(pts)].copy(), rej[:len(pts)].copy(), nm.value


def stereo_match(ext_l: OrbExtractor, ext_r: OrbExtractor, kL, dL, kR, dR, bf: float, max_d: float, debug=False):
    """Frame::ComputeStereoMatches (reference src/Frame.cc:2142-2316) -> (mvuRight, mvDepth)."""
    kL = np.ascontiguousarray(kL, KEYPOINT_DTYPE); kR = np.ascontiguousarray(kR, KEYPOINT_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    nL, nR = len(kL), len(kR)
    ur = np.full(max(nL, 1), -1, np.float32); dp = np.full(max(nL, 1), -1, np.float32)
    _check(lib().pslot_stereo_match(ext_l._h, ext_r._h, _p(kL), nL, _p(dL), _p(kR), nR, _p(dR),
                                    C.c_float(bf), C.c_float(max_d), _p(ur), _p(dp)))
    if debug:
        bd = np.zeros(max(nL, 1), np.int32); bi = np.zeros(max(nL, 1), np.int32); bs = np.zeros(max(nL, 1), np.int32)
        if nL:
            _check(lib().pslot_stereo_debug(ext_l._h, nL, _p(bd), _p(bi), _p(bs)))
        return ur[:nL], dp[:nL], bd[:nL], bi[:nL], bs[:nL]
    return ur[:nL], dp[:nL]


def extract_token(ext: OrbExtractor) -> int:
    """Identity of the feature set the last extraction left on the device (0: none)."""
    return int(lib().pslot_orb_extract_token(ext._h))


def stereo_match_resident(ext_l: OrbExtractor, ext_r: OrbExtractor, token_l: int, token_r: int, bf: float, max_d: float):
    """Frame::ComputeStereoMatches on the two feature sets the extractors' last extractions left on the device (nothing is
    uploaded) -> (mvuRight, mvDepth); PslotError (state) when either extractor has moved on since the tokens were taken."""
    cap = _check(lib().pslot_orb_max_keypoints(ext_l._h))
    ur = np.full(max(cap, 1), -1, np.float32); dp = np.full(max(cap, 1), -1, np.float32)
    n = C.c_int(0)
    _check(lib().pslot_stereo_match_resident(ext_l._h, ext_r._h, token_l, token_r, C.c_float(bf), C.c_float(max_d), _p(ur), _p(dp),
                                             cap, C.byref(n)))
    return ur[:n.value].copy(), dp[:n.value].copy()


class StereoBatch:
    """Throughput context: up to max_pairs independent stereo pairs per submit on one GPU stream."""

    def __init__(self, params, width: int, height: int, max_pairs: int, device: int = 0, stream: int = 0):
        h = C.c_void_p()
        nf, sf, nl, it, mt = params
        _check(lib().pslot_batch_create(C.byref(h), nf, sf, nl, it, mt, width, height, max_pairs, device,
                                        C.c_void_p(stream) if stream else None))
        self._h = h
        self.W, self.H, self.max_pairs = width, height, max_pairs
        self.K = _check(lib().pslot_batch_max_keypoints(h))

    def close(self):
        if getattr(self, "_h", None):
            lib().pslot_batch_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def kernel_launches(self) -> int:
        """Kernels launched by the most recent submit (before any submit: the count of the gray host path)."""
        return _check(lib().pslot_batch_kernel_launches(self._h))

    @property
    def stream(self) -> int:
        return lib().pslot_batch_stream(self._h) or 0

    def set_copy_stream(self, stream: int):
        """Share one upload stream between the batch contexts of a GPU (full-rate PCIe transfers)."""
        _check(lib().pslot_batch_set_copy_stream(self._h, C.c_void_p(stream) if stream else None))

    def alloc_out(self, npairs: int, pinned: bool = False):
        """Host output arrays in the layout pslot_batch_out documents."""
        K = self.K
        shapes = dict(counts=((2 * npairs,), np.int32), kps=((2 * npairs, K), KEYPOINT_DTYPE),
                      desc=((2 * npairs, K, 32), np.uint8), u_right=((npairs, K), np.float32),
                      depth=((npairs, K), np.float32))
        out = {}
        for k, (shp, dt) in shapes.items():
            if pinned:
                import torch
                nbytes = int(np.prod(shp)) * np.dtype(dt).itemsize
                t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
                out["_t_" + k] = t
                out[k] = t.numpy().view(dt).reshape(shp)
            else:
                out[k] = np.zeros(shp, dt)
        return out

    def submit_ptr(self, left_ptr: int, right_ptr: int, pitch: int, image_stride: int, npairs: int, host_input: bool,
                   bf: float, max_d: float, out=None):
        bo = None
        if out is not None:
            bo = BatchOut(*(out[k].ctypes.data for k in ("counts", "kps", "desc", "u_right", "depth")))
        _check(lib().pslot_batch_submit(self._h, C.c_void_p(left_ptr), C.c_void_p(right_ptr), pitch, image_stride,
                                        npairs, 1 if host_input else 0, C.c_float(bf), C.c_float(max_d),
                                        C.byref(bo) if bo is not None else None))

    def submit_color(self, left: np.ndarray, right: np.ndarray, rgb: bool, bf: float, max_d: float, out=None):
        """left/right: uint8 [npairs, H, W, 3|4] interleaved colour host arrays."""
        assert left.dtype == np.uint8 and left.shape == right.shape and left.ndim == 4 and left.strides == right.strides
        bo = None
        if out is not None:
            bo = BatchOut(*(out[k].ctypes.data for k in ("counts", "kps", "desc", "u_right", "depth")))
        _check(lib().pslot_batch_submit_color(self._h, C.c_void_p(left.ctypes.data), C.c_void_p(right.ctypes.data),
                                              left.strides[1], left.strides[0], left.shape[0], left.shape[3],
                                              1 if rgb else 0, C.c_float(bf), C.c_float(max_d),
                                              C.byref(bo) if bo is not None else None))

    def submit(self, left: np.ndarray, right: np.ndarray, bf: float, max_d: float, out=None):
        """left/right: uint8 [npairs, H, W] host arrays (C-contiguous rows)."""
        assert left.dtype == np.uint8 and left.shape == right.shape and left.ndim == 3
        assert left.strides[2] == 1 and right.strides == left.strides
        self.submit_ptr(left.ctypes.data, right.ctypes.data, left.strides[1], left.strides[0], left.shape[0], True,
                        bf, max_d, out)

    def assign_grid(self, spec: GridSpec, npairs: int):
        """Feature grids of the left keypoints of the last submit -> (cell_start [npairs, cells+1], indices [npairs, K]);
        asynchronous: valid after wait()."""
        cs = np.zeros((npairs, spec.cols * spec.rows + 1), np.int32); idx = np.zeros((npairs, self.K), np.int32)
        _check(lib().pslot_batch_assign_grid(self._h, C.byref(spec), _p(cs), _p(idx)))
        return cs, idx

    def assign_features(self, masks, object_ids):
        """Frame::AssignFeatures for the left keypoints of every pair of the last submit (they are still on the device):
        masks [npairs, H, W] u8, object_ids [npairs, J] -> (cls [npairs, K], order [npairs, K], class_start [npairs, J + 2]);
        asynchronous: valid after wait()."""
        masks = np.ascontiguousarray(masks, np.uint8)
        ids = np.ascontiguousarray(object_ids, np.int32)
        npairs, J = ids.shape
        assert masks.shape == (npairs, self.H, self.W)
        cls = np.zeros((npairs, self.K), np.int32); order = np.zeros((npairs, self.K), np.int32)
        cs = np.zeros((npairs, J + 2), np.int32)
        _check(lib().pslot_batch_assign_features(self._h, _p(masks), self.W, self.W * self.H, _p(ids), J, _p(cls), _p(order), _p(cs)))
        self._keep = (masks, ids)  # the copies run asynchronously
        return cls, order, cs

    def wait(self):
        _check(lib().pslot_batch_wait(self._h))

    STAGES = ("ingest_copy", "pyramid", "fast", "quadtree", "describe", "stereo_match", "stereo_reject", "result_copy")

    def set_profiling(self, enable: bool):
        _check(lib().pslot_batch_set_profiling(self._h, 1 if enable else 0))

    def stats(self) -> dict:
        """Work counters of the last batch (candidates, keypoints, SAD evaluations, final matches)."""
        st = np.zeros(4, np.int64)
        _check(lib().pslot_batch_stats(self._h, _p(st)))
        return dict(zip(("candidates", "keypoints", "sad_evals", "matches"), (int(v) for v in st)))

    def stereo_counters(self) -> dict:
        """Row-band candidates scanned / descriptors compared by the profiled submits since the last call."""
        c = np.zeros(2, np.int64)
        _check(lib().pslot_batch_stereo_counters(self._h, _p(c)))
        return {"row_candidates": int(c[0]), "hamming_evals": int(c[1])}

    def stage_times_ms(self) -> np.ndarray:
        ms = np.zeros(len(self.STAGES), np.float32)
        _check(lib().pslot_batch_stage_times(self._h, _p(ms)))
        return ms

    def run(self, left, right, bf, max_d):
        """Synchronous convenience: -> list of per-pair dicts (kL, dL, kR, dR, u_right, depth)."""
        n = left.shape[0]
        out = self.alloc_out(n)
        self.submit(left, right, bf, max_d, out)
        self.wait()
        res = []
        for i in range(n):
            nl, nr = int(out["counts"][i]), int(out["counts"][n + i])
            res.append(dict(kL=out["kps"][i, :nl].copy(), dL=out["desc"][i, :nl].copy(),
                            kR=out["kps"][n + i, :nr].copy(), dR=out["desc"][n + i, :nr].copy(),
                            u_right=out["u_right"][i, :nl].copy(), depth=out["depth"][i, :nl].copy()))
        return res
