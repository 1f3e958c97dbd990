"""The REFERENCE's own code (oracle/_ref/libref_orb{,_mono}.so, built by `make -C oracle ref` from the sources where
they lie under /root/reference) driven from Python.  Test infrastructure only: the parity pin of the oracle.

What is inside the libraries (oracle/Makefile, oracle/ref_compat/README.md):
  * src/ORBextractor.cc, unmodified, over stand-in cv::Mat containers;
  * Frame::ComputeStereoMatches / AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid (src/Frame.cc) and
    ORBmatcher::SearchByProjection (local map + last frame) / DescriptorDistance (src/ORBmatcher.cc), lifted verbatim at
    build time;
  * `mono` variant: an operator new that never reuses memory (heap addresses = creation order), which removes the
    address-dependent tie of ORBextractor.cc:684; `stock` variant: glibc malloc, the reference as anyone builds it.

The extractor's five OpenCV primitive calls are served either by the real OpenCV 4.13 inside cv2 (`primitives="cv2"`,
through ctypes callbacks) or by the C oracle's restatements, themselves pinned against cv2 (`primitives="native"`)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

REF_DIR = Path(__file__).resolve().parent / "_ref"
LIBS = {"stock": REF_DIR / "libref_orb.so", "mono": REF_DIR / "libref_orb_mono.so",
        "mono_fma": REF_DIR / "libref_orb_mono_fma.so"}
LIB = LIBS["stock"]
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])
MAP_POINT_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"),
                            ("level", "<i4"), ("flags", "<i4")])
_state: dict = {}


def available(variant: str = "stock") -> bool:
    return LIBS[variant].exists()


def _view(ptr, w, h, step):
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(h, step))
    return a[:, :w]


_u8p, _sz = C.c_void_p, C.c_size_t
_RESIZE = C.CFUNCTYPE(None, _u8p, C.c_int, C.c_int, _sz, _u8p, C.c_int, C.c_int, _sz, C.c_int)
_BORDER = C.CFUNCTYPE(None, _u8p, C.c_int, C.c_int, _sz, _u8p, _sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_FAST = C.CFUNCTYPE(C.c_int, _u8p, C.c_int, C.c_int, _sz, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int)
_BLUR = C.CFUNCTYPE(None, _u8p, C.c_int, C.c_int, _sz, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int)
_ATAN2 = C.CFUNCTYPE(C.c_float, C.c_float, C.c_float)


def _cv2_callbacks():
    import cv2

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

    return (_RESIZE(resize), _BORDER(border), _FAST(fast), _BLUR(blur), _ATAN2(atan2))


def _load(variant: str = "stock"):
    if variant in _state:
        return _state[variant]["lib"]
    lib = C.CDLL(str(LIBS[variant]))
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    lib.ref_set_callbacks.argtypes = [_RESIZE, _BORDER, _FAST, _BLUR, _ATAN2]
    lib.ref_create.restype = vp
    lib.ref_create.argtypes = [i32, f32, i32, i32, i32]
    lib.ref_destroy.argtypes = [vp]
    lib.ref_extract.argtypes = [vp, vp, i32, i32, C.c_size_t, vp, vp, i32]
    lib.ref_level.argtypes = [vp, i32, vp, C.c_size_t, C.POINTER(i32), C.POINTER(i32)]
    lib.reff_stereo.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, f32, f32, vp, vp]
    lib.reff_obj_stereo.argtypes = lib.reff_stereo.argtypes
    lib.reff_hamming256.argtypes = [vp, vp]
    lib.reff_assign_features.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, vp, i32, f32, f32, f32, f32, f32,
                                         f32, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.reff_frame_create.restype = vp
    lib.reff_frame_create.argtypes = [vp, i32, vp, vp, vp, i32, f32, f32, f32, f32, f32, f32]
    lib.reff_frame_destroy.argtypes = [vp]
    lib.reff_frame_grid.argtypes = [vp, vp, vp]
    lib.reff_frame_features_in_area.argtypes = [vp, f32, f32, f32, i32, i32, i32, vp]
    lib.reff_search_by_projection.argtypes = [vp, vp, vp, i32, vp, f32, f32, vp]
    lib.reff_search_by_projection_last.argtypes = [vp, vp, i32, vp, vp, vp, f32, f32, f32, f32, f32, f32, f32, vp, f32,
                                                   i32, i32, vp]
    lib.reff_search_by_projection_kf.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, f32, f32, f32, f32, vp, f32, i32, i32, vp]
    lib.reff_search_by_projection_obj.argtypes = [vp, vp, vp, i32, vp, vp, f32, f32, vp]
    lib.reff_search_by_projection_sim3.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, vp, i32, vp]
    _state[variant] = dict(lib=lib, cbs=None, primitives=None)
    return lib


def _set_primitives(variant: str, primitives: str):
    """The callback table is per library, not per extractor: switch it before every call that runs the extractor."""
    st = _state[variant]
    if st["primitives"] == primitives:
        return
    if primitives == "cv2":
        st["cbs"] = st["cbs"] or _cv2_callbacks()  # keep the callback objects alive
        st["lib"].ref_set_callbacks(*st["cbs"])
    elif primitives == "native":
        st["lib"].ref_set_callbacks(_RESIZE(0), _BORDER(0), _FAST(0), _BLUR(0), _ATAN2(0))
    else:
        raise ValueError(primitives)
    st["primitives"] = primitives


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class ReferenceExtractor:
    """ORB_SLAM2::ORBextractor of the reference tree (include/ORBextractor.h:45-111), as compiled from its own source."""

    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th, variant="stock", primitives="cv2"):
        self._lib = _load(variant)
        self.variant, self.primitives = variant, primitives
        self._h = self._lib.ref_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        self.cap = nfeatures + 16 * nlevels + 64

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_destroy(self._h)
            self._h = None

    def extract(self, img: np.ndarray):
        img = np.ascontiguousarray(img, np.uint8)
        _set_primitives(self.variant, self.primitives)
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


def stereo_match(eL: ReferenceExtractor, eR: ReferenceExtractor, kL, dL, kR, dR, mbf: float, mb: float):
    """Frame::ComputeStereoMatches (Frame.cc:2142-2316) on the pyramids eL / eR hold -> (mvuRight, mvDepth)."""
    assert eL._lib is eR._lib
    kL = np.ascontiguousarray(kL, KEYPOINT_DTYPE); kR = np.ascontiguousarray(kR, KEYPOINT_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    ur = np.zeros(max(len(kL), 1), np.float32); dp = np.zeros(max(len(kL), 1), np.float32)
    eL._lib.reff_stereo(eL._h, eR._h, _p(kL), len(kL), _p(dL), _p(kR), len(kR), _p(dR), mbf, mb, _p(ur), _p(dp))
    return ur[:len(kL)], dp[:len(kL)]


def obj_stereo_match(eL: ReferenceExtractor, eR: ReferenceExtractor, kL, dL, kR, dR, mbf: float, mb: float):
    """Frame::ComputeObjStereoMatches (Frame.cc:2318-2503) on an object feature set -> (uRight, depth)."""
    kL = np.ascontiguousarray(kL, KEYPOINT_DTYPE); kR = np.ascontiguousarray(kR, KEYPOINT_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    ur = np.zeros(max(len(kL), 1), np.float32); dp = np.zeros(max(len(kL), 1), np.float32)
    rc = eL._lib.reff_obj_stereo(eL._h, eR._h, _p(kL), len(kL), _p(dL), _p(kR), len(kR), _p(dR), mbf, mb, _p(ur), _p(dp))
    assert rc == 0
    return ur[:len(kL)], dp[:len(kL)]


def hamming256(a, b, variant="stock") -> int:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return int(_load(variant).reff_hamming256(_p(a), _p(b)))


def assign_features(kps, desc, u_right, depth, okps, odesc, ou_right, odepth, mask, object_ids, th_depth, min_x, max_x,
                    min_y, max_y, inv_w, inv_h, variant="stock"):
    """The reference's own Frame::AssignFeatures (src/Frame.cc:759-977, with AssignDetObjFeasToGrid :1863-1888) on a stub frame
    holding the given ORB features (kps ...) and dense object features (okps ...) -> dict of index lists:
    background, det_of (original index of each kept detection), ori[k], obj[k], init_n[k], grid[k] = (cell_start, indices)."""
    lib = _load(variant)
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE); okps = np.ascontiguousarray(okps, KEYPOINT_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8); odesc = np.ascontiguousarray(odesc, np.uint8)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    u_right, depth, ou_right, odepth = f(u_right), f(depth), f(ou_right), f(odepth)
    mask = np.ascontiguousarray(mask, np.uint8)
    ids = np.ascontiguousarray(object_ids, np.int32)
    n, on, J = len(kps), len(okps), len(ids)
    i32 = lambda m: np.zeros(max(m, 1), np.int32)
    bg, det_of, ori, ori_s, obj, obj_s, init_n = i32(n), i32(J), i32(n), i32(J + 1), i32(on), i32(J + 1), i32(J)
    gcs, gidx = i32(J * (64 * 48 + 1)), i32(on)
    bg_n, n_det = C.c_int(0), C.c_int(0)
    rc = lib.reff_assign_features(_p(kps), n, _p(desc), _p(u_right), _p(depth), _p(okps), on, _p(odesc), _p(ou_right), _p(odepth),
                                  _p(mask), mask.shape[1], mask.shape[0], _p(ids), J, th_depth, min_x, max_x, min_y, max_y,
                                  inv_w, inv_h, _p(bg), C.byref(bg_n), _p(det_of), C.byref(n_det), _p(ori), _p(ori_s), _p(obj),
                                  _p(obj_s), _p(init_n), _p(gcs), _p(gidx))
    if rc:
        raise RuntimeError(f"reference AssignFeatures moved an attribute away from its keypoint (rc {rc})")
    nd = n_det.value
    out = {"background": bg[:bg_n.value].copy(), "det_of": det_of[:nd].copy(), "init_n": init_n[:nd].copy(), "ori": [], "obj": [],
           "grid": []}
    g0 = 0
    for k in range(nd):
        out["ori"].append(ori[ori_s[k]:ori_s[k + 1]].copy())
        out["obj"].append(obj[obj_s[k]:obj_s[k + 1]].copy())
        cs = gcs[k * (64 * 48 + 1):(k + 1) * (64 * 48 + 1)].copy()
        out["grid"].append((cs, gidx[g0:g0 + cs[-1]].copy()))
        g0 += cs[-1]
    return out


class ReferenceFrame:
    """A Frame shell holding mvKeysUn / mDescriptors / mvuRight / mGrid, on which the reference's own
    AssignFeaturesToGrid, GetFeaturesInArea and ORBmatcher::SearchByProjection bodies run."""

    def __init__(self, kps_un, desc, u_right, scale_factors, min_x, max_x, min_y, max_y, inv_w, inv_h, variant="stock"):
        self._lib = _load(variant)
        self.kps = np.ascontiguousarray(kps_un, KEYPOINT_DTYPE)
        self.n = len(self.kps)
        desc = np.ascontiguousarray(desc, np.uint8); u_right = np.ascontiguousarray(u_right, np.float32)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        self._h = self._lib.reff_frame_create(_p(self.kps), self.n, _p(desc), _p(u_right), _p(sf), len(sf),
                                              min_x, max_x, min_y, max_y, inv_w, inv_h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.reff_frame_destroy(self._h)
            self._h = None

    def grid(self):
        cs = np.zeros(64 * 48 + 1, np.int32); idx = np.zeros(max(self.n, 1), np.int32)
        n = self._lib.reff_frame_grid(self._h, _p(cs), _p(idx))
        return cs, idx[:n].copy()

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(max(self.n, 1), np.int32)
        n = self._lib.reff_frame_features_in_area(self._h, x, y, r, min_level, max_level, len(out), _p(out))
        return out[:n].copy()

    def search_by_projection(self, mps, mp_desc, occupied, th, nn_ratio):
        """-> (nmatches, slot_of_feature): index of the map point each feature holds afterwards (-1 none, -2 prior)."""
        mps = np.ascontiguousarray(mps, MAP_POINT_DTYPE); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        occ = np.ascontiguousarray(occupied, np.uint8)
        slot = np.zeros(max(self.n, 1), np.int32)
        n = self._lib.reff_search_by_projection(self._h, _p(mps), _p(mp_desc), len(mps), _p(occ), th, nn_ratio, _p(slot))
        return n, slot[:self.n]

    def search_by_projection_kf(self, kf_keys_un, world_xyz, kf_desc, kf_flags, level, min_dist, max_dist, fx, fy, cx, cy,
                                occupied, th, orb_dist, check_orientation):
        kk = np.ascontiguousarray(kf_keys_un, KEYPOINT_DTYPE); xyz = np.ascontiguousarray(world_xyz, np.float32)
        kd = np.ascontiguousarray(kf_desc, np.uint8); kf = np.ascontiguousarray(kf_flags, np.uint8)
        lv = np.ascontiguousarray(level, np.int32); mn = np.ascontiguousarray(min_dist, np.float32)
        mx = np.ascontiguousarray(max_dist, np.float32); occ = np.ascontiguousarray(occupied, np.uint8)
        slot = np.zeros(max(self.n, 1), np.int32)
        n = self._lib.reff_search_by_projection_kf(self._h, _p(kk), len(kk), _p(xyz), _p(kd), _p(kf), _p(lv), _p(mn), _p(mx),
                                                   fx, fy, cx, cy, _p(occ), th, int(orb_dist), int(check_orientation), _p(slot))
        return n, slot[:self.n]

    def search_by_projection_obj(self, mps, mp_desc, occupied, bbox, th, nn_ratio):
        mps = np.ascontiguousarray(mps); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        occ = np.ascontiguousarray(occupied, np.uint8); bb = np.ascontiguousarray(bbox, np.int32)
        slot = np.zeros(max(self.n, 1), np.int32)
        n = self._lib.reff_search_by_projection_obj(self._h, _p(mps), _p(mp_desc), len(mps), _p(occ), _p(bb), th, nn_ratio, _p(slot))
        return n, slot[:self.n]

    def search_by_projection_sim3(self, scw, world_xyz, normals, pt_desc, pt_flags, level, min_dist, max_dist, fx, fy, cx, cy,
                                  pre_matched, th):
        S = np.ascontiguousarray(scw, np.float32).reshape(16); xyz = np.ascontiguousarray(world_xyz, np.float32)
        nr = np.ascontiguousarray(normals, np.float32); pd = np.ascontiguousarray(pt_desc, np.uint8)
        pf = np.ascontiguousarray(pt_flags, np.uint8); lv = np.ascontiguousarray(level, np.int32)
        mn = np.ascontiguousarray(min_dist, np.float32); mx = np.ascontiguousarray(max_dist, np.float32)
        pm = np.ascontiguousarray(pre_matched, np.int32)
        slot = np.zeros(max(self.n, 1), np.int32)
        n = self._lib.reff_search_by_projection_sim3(self._h, _p(S), _p(xyz), _p(nr), _p(pd), _p(pf), _p(lv), _p(mn), _p(mx),
                                                     len(xyz), fx, fy, cx, cy, _p(pm), int(th), _p(slot))
        return n, slot[:self.n]

    def search_by_projection_last(self, last_keys_un, world_xyz, last_desc, last_flags, last_tz, fx, fy, cx, cy, mbf, mb,
                                  occupied, th, mono, check_orientation):
        lk = np.ascontiguousarray(last_keys_un, KEYPOINT_DTYPE); xyz = np.ascontiguousarray(world_xyz, np.float32)
        ld = np.ascontiguousarray(last_desc, np.uint8); lf = np.ascontiguousarray(last_flags, np.uint8)
        occ = np.ascontiguousarray(occupied, np.uint8)
        slot = np.zeros(max(self.n, 1), np.int32)
        n = self._lib.reff_search_by_projection_last(self._h, _p(lk), len(lk), _p(xyz), _p(ld), _p(lf), last_tz, fx, fy,
                                                     cx, cy, mbf, mb, _p(occ), th, int(mono), int(check_orientation),
                                                     _p(slot))
        return n, slot[:self.n]
