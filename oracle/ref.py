"""ctypes wrapper around oracle/_ref/libref.so: the REFERENCE'S OWN sources for the hot path, compiled unmodified
from /root/reference against the test-only cv:: shim (oracle/ref_build/).

TEST INFRASTRUCTURE ONLY.  Used by tests/ (to pin the C oracle and the CUDA path against the reference's real code) and by
bench.py's --impl reference / cpu_baseline legs (kind "reference").  libref.so is built in THIS container (the only place
/root/reference exists) and travels to the GPU box inside the snapshot; nothing here reads /root/reference at run time.
The interface mirrors oracle/oracle.py so a test can run the same call against either.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from .oracle import KEYPOINT_DTYPE, KEYLINE_DTYPE, LineParams, _img, _p, build as _build_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SRC = "/root/reference"
_LIBS = {}


class FldParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("nlevels", C.c_int), ("scale", C.c_double), ("length_threshold", C.c_int),
                ("distance_threshold", C.c_float), ("canny_th1", C.c_double), ("canny_th2", C.c_double),
                ("canny_aperture_size", C.c_int), ("do_merge", C.c_int)]


def so_path(variant=""):
    return os.path.join(_HERE, "_ref", "libref%s.so" % variant)


def available(variant=""):
    return os.path.exists(so_path(variant))


def build(force=False):
    """Compile oracle/_ref/libref.so (-O3 -march=native, the reference's flags) and libref_generic.so (-O3) when the
    reference tree is present; a no-op on the GPU box, which uses the prebuilt files from the snapshot."""
    if not os.path.isdir(_REF_SRC):
        return available()
    _build_oracle()
    for variant, march in (("", "-march=native"), ("_generic", "")):
        args = ["make", "-C", os.path.join(_HERE, "ref_build"), "-s", "SUFFIX=" + variant, "REF_MARCH=" + march]
        if force:
            args.append("-B")
        subprocess.check_call(args)
    return True


def lib(variant=""):
    if variant not in _LIBS:
        if not available(variant):
            build()
        # liboracle.so (the primitives the cv:: shim forwards to) must be loaded first; libref.so carries an rpath to it
        L = C.CDLL(so_path(variant))
        vp = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_orb_create.restype = vp
        L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_orb_destroy.argtypes = [vp]
        L.ref_orb_features_per_level.argtypes = [vp, C.c_int]
        L.ref_orb_umax.argtypes = [vp, C.c_int]
        L.ref_orb_scale_factor.argtypes = [vp, C.c_int]
        L.ref_orb_scale_factor.restype = C.c_float
        L.ref_orb_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, vp, C.c_int]
        L.ref_orb_level.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), vp]
        L.ref_orb_level_keypoints.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, C.c_int, vp, C.c_int]
        L.ref_distribute_octree.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.ref_line_features_per_level.argtypes = [C.POINTER(LineParams), C.c_int]
        L.ref_lsd_detect_keylines.argtypes = [C.POINTER(LineParams), vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int]
        L.ref_lbd_compute.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp, vp]
        L.ref_line_extract.argtypes = [C.POINTER(LineParams), vp, C.c_int, C.c_int, C.c_size_t, vp, vp, vp, C.c_int]
        L.ref_fld_extract.argtypes = [C.POINTER(FldParams), vp, C.c_int, C.c_int, C.c_size_t, vp, vp, vp, C.c_int]
        L.ref_fld_detect.argtypes = [C.POINTER(FldParams), vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int]
        L.ref_orb_descriptor_distance.argtypes = [vp, vp]
        L.ref_line_descriptor_distance.argtypes = [vp, vp]
        L.ref_match_nnr.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, vp]
        L.ref_three_maxima.argtypes = [vp, C.c_int, vp]
        L.ref_set_heap_mode.argtypes = [C.c_int]
        L.ref_std_sort_desc.argtypes = [vp, C.c_int, vp]
        L.ref_orb_search_for_initialization.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_int, vp]
        L.ref_line_search_by_knn.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, C.c_float, vp]
        L.ref_line_search_for_triangulation.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, vp]
        _LIBS[variant] = L
    return _LIBS[variant]


def set_heap_mode(mode, variant=""):
    """0: operator new = glibc malloc (the reference as it runs; DistributeOctTree's address tie-break then depends on
    the allocation history).  1: monotone bump arena reset per call (fresh-heap tie order, deterministic)."""
    lib(variant).ref_set_heap_mode(int(mode))


def _check(n, L):
    if n == -1000:
        raise RuntimeError("reference threw: %s" % L.ref_last_error().decode())
    assert n >= 0, "capacity too small"
    return n


class ORBextractor:
    """The reference's PL_SLAM::ORBextractor itself (src/ORBextractor.cc)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, variant=""):
        self._L = lib(variant)
        self._h = self._L.ref_orb_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_orb_destroy(self._h)
            self._h = None

    def features_per_level(self):
        return [self._L.ref_orb_features_per_level(self._h, l) for l in range(self.nlevels)]

    def scale_factors(self):
        return [self._L.ref_orb_scale_factor(self._h, l) for l in range(self.nlevels)]

    def umax(self):
        return [self._L.ref_orb_umax(self._h, v) for v in range(16)]

    def __call__(self, image):
        img = _img(image)
        cap = self.nfeatures * 2 + 64 * self.nlevels
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = _check(self._L.ref_orb_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap), self._L)
        return kps[:n].copy(), desc[:n].copy()

    def level_image(self, level):
        w, h = C.c_int(), C.c_int()
        self._L.ref_orb_level(self._h, level, C.byref(w), C.byref(h), None)
        out = np.empty((h.value, w.value), np.uint8)
        self._L.ref_orb_level(self._h, level, C.byref(w), C.byref(h), _p(out))
        return out

    def level_keypoints(self, image, level):
        img = _img(image)
        cap = self.nfeatures * 2 + 64
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        n = _check(self._L.ref_orb_level_keypoints(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], level, _p(kps), cap), self._L)
        return kps[:n].copy()

    def distribute_octree(self, xs, ys, resp, minX, maxX, minY, maxY, N):
        xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); resp = np.ascontiguousarray(resp, np.int32)
        out = np.empty(max(1, len(xs)), np.int32)
        n = _check(self._L.ref_distribute_octree(self._h, _p(xs), _p(ys), _p(resp), len(xs), minX, maxX, minY, maxY, N, _p(out), len(out)), self._L)
        return out[:n].copy()


def features_per_level_lines(params, variant=""):
    return [lib(variant).ref_line_features_per_level(C.byref(params), l) for l in range(params.nlevels)]


def lsd_detect_keylines(params, img, variant=""):
    L = lib(variant)
    img = _img(img)
    cap = 1 << 16
    kl = np.zeros(cap, KEYLINE_DTYPE)
    n = _check(L.ref_lsd_detect_keylines(C.byref(params), _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), cap), L)
    return kl[:n].copy()


def lbd_compute(img, keylines, want_float=False, variant=""):
    L = lib(variant)
    img = _img(img)
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    n = len(kl)
    desc = np.zeros((n, 32), np.uint8)
    fdesc = np.zeros((n, 72), np.float32)
    _check(L.ref_lbd_compute(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), n, _p(desc), _p(fdesc) if want_float else None), L)
    return (desc, fdesc) if want_float else desc


def line_extract(params, img, variant=""):
    """The reference's Lineextractor::ComputeLsdWithLbd itself (src/Lineextractor.cc:112-212)."""
    L = lib(variant)
    img = _img(img)
    cap = max(16, params.nfeatures * 2 + 16)
    kl = np.zeros(cap, KEYLINE_DTYPE)
    mid = np.zeros(cap, KEYPOINT_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = _check(L.ref_line_extract(C.byref(params), _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), _p(mid), _p(desc), cap), L)
    return kl[:n].copy(), mid[:n].copy(), desc[:n].copy()


def fld_params(nfeatures=240, nlevels=1, scale=1.05, length_threshold=10, distance_threshold=1.414213562, canny_th1=50.0,
               canny_th2=100.0, canny_aperture_size=3, do_merge=False):
    return FldParams(nfeatures, nlevels, scale, length_threshold, distance_threshold, canny_th1, canny_th2, canny_aperture_size,
                     1 if do_merge else 0)


def fld_extract(params, img, variant=""):
    """The reference's Lineextractor::ComputeFldWithLbd itself (src/Lineextractor.cc:242-336)."""
    L = lib(variant)
    img = _img(img)
    cap = max(16, params.nfeatures * 2 + 16)
    kl = np.zeros(cap, KEYLINE_DTYPE)
    mid = np.zeros(cap, KEYPOINT_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = _check(L.ref_fld_extract(C.byref(params), _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), _p(mid), _p(desc), cap), L)
    return kl[:n].copy(), mid[:n].copy(), desc[:n].copy()


def fld_detect(params, img, variant=""):
    L = lib(variant)
    img = _img(img)
    cap = 1 << 16
    lines = np.empty((cap, 4), np.float32)
    n = _check(L.ref_fld_detect(C.byref(params), _p(img), img.shape[1], img.shape[0], img.strides[0], _p(lines), cap), L)
    return lines[:n].copy()


def descriptor_distance(a, b, line=False, variant=""):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    L = lib(variant)
    return (L.ref_line_descriptor_distance if line else L.ref_orb_descriptor_distance)(_p(a), _p(b))


def match_nnr(q, t, nnr, variant=""):
    L = lib(variant)
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    m = np.empty(max(len(q), 1), np.int32)
    n = _check(L.ref_match_nnr(_p(q), len(q), _p(t), len(t), nnr, _p(m)), L)
    return m[:len(q)], n


def three_maxima(sizes, variant=""):
    s = np.ascontiguousarray(sizes, np.int32)
    out = np.zeros(3, np.int32)
    lib(variant).ref_three_maxima(_p(s), len(s), _p(out))
    return tuple(int(v) for v in out)


def std_sort_desc(keys, variant=""):
    """The toolchain's real std::sort with `a.response > b.response` (include/Lineextractor.h:66-71) -> permutation."""
    k = np.ascontiguousarray(keys, np.float32)
    p = np.zeros(max(len(k), 1), np.int32)
    lib(variant).ref_std_sort_desc(_p(k), len(k), _p(p))
    return p[:len(k)]


def orb_search_for_initialization(k1, d1, k2, d2, bounds, prev, window, nnratio, check_ori, variant=""):
    """The reference's ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:406-521) on mock Frames -> (nmatches, matches12, prev)."""
    L = lib(variant)
    k1 = np.ascontiguousarray(k1, KEYPOINT_DTYPE); k2 = np.ascontiguousarray(k2, KEYPOINT_DTYPE)
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    b = np.ascontiguousarray(bounds, np.float32); pv = np.ascontiguousarray(prev, np.float32).copy()
    m = np.full(max(len(k1), 1), -1, np.int32)
    n = _check(L.ref_orb_search_for_initialization(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), _p(b), _p(pv), window, nnratio, int(check_ori), _p(m)), L) if True else 0
    return n, m[:len(k1)], pv


def line_search_by_knn(dkf, df, ml_state, ml_len, f_linelen, nnr, checklen, lengtherr, variant=""):
    """The reference's Linematcher::SearchByKNN (src/Linematcher.cc:437-517) -> (return value, per frame line: key-frame line index or -1)."""
    L = lib(variant)
    dkf = np.ascontiguousarray(dkf, np.uint8); df = np.ascontiguousarray(df, np.uint8)
    st = np.ascontiguousarray(ml_state, np.uint8); ln = np.ascontiguousarray(ml_len, np.float32); fl = np.ascontiguousarray(f_linelen, np.float32)
    out = np.full(max(len(df), 1), -1, np.int32)
    r = L.ref_line_search_by_knn(_p(dkf), len(dkf), _p(df), len(df), _p(st), _p(ln), _p(fl), nnr, int(checklen), lengtherr, _p(out))
    if r == -1000:
        raise RuntimeError("reference threw: %s" % L.ref_last_error().decode())
    return r, out[:len(df)]


def line_search_for_triangulation(d1, d2, has1, has2, mid1, mid2, scale, sigma2, cam, pose, F12, nnr, variant=""):
    """The reference's Linematcher::SearchForTriangulation (src/Linematcher.cc:804-879) -> matched pairs (n x 2)."""
    L = lib(variant)
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    h1 = np.ascontiguousarray(has1, np.uint8); h2 = np.ascontiguousarray(has2, np.uint8)
    m1 = np.ascontiguousarray(mid1, KEYPOINT_DTYPE); m2 = np.ascontiguousarray(mid2, KEYPOINT_DTYPE)
    sc = np.ascontiguousarray(scale, np.float32); sg = np.ascontiguousarray(sigma2, np.float32)
    cm = np.ascontiguousarray(cam, np.float32); ps = np.ascontiguousarray(pose, np.float32); f = np.ascontiguousarray(F12, np.float32)
    pairs = np.zeros((max(len(d1), 1), 2), np.int32)
    n = _check(L.ref_line_search_for_triangulation(_p(d1), len(d1), _p(d2), len(d2), _p(h1), _p(h2), _p(m1), _p(m2), _p(sc), _p(sg), len(sc),
                                                   _p(cm), _p(ps), _p(f), nnr, _p(pairs)), L)
    return pairs[:n].copy()
