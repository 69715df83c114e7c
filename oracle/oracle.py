"""ctypes wrapper around oracle/liboracle.so (the CPU oracle).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never import this from
spl_slam_b200/ (the product path has no CPU fallback).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
KEYLINE_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"),
                          ("pt_x", "<f4"), ("pt_y", "<f4"), ("response", "<f4"), ("size", "<f4"),
                          ("startPointX", "<f4"), ("startPointY", "<f4"),
                          ("endPointX", "<f4"), ("endPointY", "<f4"),
                          ("sPointInOctaveX", "<f4"), ("sPointInOctaveY", "<f4"),
                          ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
                          ("lineLength", "<f4"), ("numOfPixels", "<i4")])
assert KEYPOINT_DTYPE.itemsize == 28 and KEYLINE_DTYPE.itemsize == 68


class LineParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("nlevels", C.c_int), ("refine", C.c_int),
                ("scale", C.c_double), ("sigma_scale", C.c_double), ("quant", C.c_double),
                ("ang_th", C.c_double), ("log_eps", C.c_double), ("density_th", C.c_double),
                ("n_bins", C.c_int), ("min_line_length", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_prims.c", "orc_orb.c", "orc_lsd.c", "orc_lbd.c", "orc_bow.c", "orc_fld.c", "plf_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        u8p, i32p, f32p, vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_orb_create.restype = vp
        L.orc_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_orb_destroy.argtypes = [vp]
        L.orc_orb_features_per_level.argtypes = [vp, C.c_int]
        L.orc_orb_scale_factor.argtypes = [vp, C.c_int]
        L.orc_orb_scale_factor.restype = C.c_float
        L.orc_orb_umax.argtypes = [vp, C.c_int]
        L.orc_orb_extract.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_size_t, vp, u8p, C.c_int]
        L.orc_orb_level_size.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_orb_level_image.restype = vp
        L.orc_orb_level_image.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_orb_level_blurred.restype = vp
        L.orc_orb_level_blurred.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_orb_level_raw_count.argtypes = [vp, C.c_int]
        L.orc_orb_level_raw.argtypes = [vp, C.c_int, i32p, i32p, i32p]
        L.orc_orb_level_kept_count.argtypes = [vp, C.c_int]
        L.orc_grid_candidates.argtypes = [vp, vp, C.c_int, vp, f32p, f32p, f32p, i32p, i32p, C.c_int, i32p, i32p, C.c_int]
        L.orc_bow_transform.argtypes = [C.c_int, C.c_int, i32p, u8p, vp, u8p, u8p, C.c_int, C.c_int, i32p, vp, i32p]
        L.orc_undistort_points.argtypes = [f32p, f32p, C.c_int, f32p, C.c_int, f32p]
        L.orc_stereo_match.argtypes = [vp, vp, vp, u8p, C.c_int, vp, u8p, C.c_int, C.c_float, C.c_float, f32p, f32p]
        L.orc_distribute_octree.argtypes = [i32p, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, i32p, C.c_int]
        L.orc_resize_linear_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_int, C.c_int, C.c_size_t]
        L.orc_resize_linear_exact_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_int, C.c_int,
                                                 C.c_size_t, C.c_double, C.c_double]
        L.orc_border_reflect101_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_int, C.c_size_t]
        L.orc_gauss_kernel_q8.argtypes = [C.c_int, C.c_double, i32p]
        L.orc_gauss_blur_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_size_t, C.c_int, C.c_double]
        L.orc_pyrdown_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_size_t]
        L.orc_sobel3_s16.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, vp, vp]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_fast9.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, i32p, i32p, i32p, C.c_int]
        L.orc_lsd_detect.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_int, f32p, C.c_int]
        L.orc_line_features_per_level.argtypes = [C.POINTER(LineParams), C.c_int]
        L.orc_lsd_detect_keylines.argtypes = [C.POINTER(LineParams), u8p, C.c_int, C.c_int, C.c_size_t, vp, C.c_int]
        L.orc_lbd_compute.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, u8p, f32p]
        L.orc_line_extract.argtypes = [C.POINTER(LineParams), u8p, C.c_int, C.c_int, C.c_size_t, vp, vp, u8p, C.c_int]
        L.orc_std_sort_desc.argtypes = [f32p, C.c_int, i32p]
        L.orc_descriptor_distance.argtypes = [u8p, u8p]
        L.orc_knn2.argtypes = [u8p, C.c_int, u8p, C.c_long, i32p, i32p]
        L.orc_match_nnr.argtypes = [u8p, C.c_int, u8p, C.c_long, C.c_float, i32p]
        L.orc_hamming_candidates.argtypes = [u8p, C.c_int, u8p, i32p, i32p, i32p, i32p, i32p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


# ---------------- primitives ----------------
def resize_linear(img, dw, dh):
    img = _img(img)
    out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw)
    return out


def resize_linear_exact(img, fx, fy=None):
    img = _img(img)
    fy = fx if fy is None else fy
    dw, dh = int(np.rint(img.shape[1] * fx)), int(np.rint(img.shape[0] * fy))
    out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_exact_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw, fx, fy)
    return out


def border_reflect101(img, b):
    img = _img(img)
    h, w = img.shape
    out = np.empty((h + 2 * b, w + 2 * b), np.uint8)
    lib().orc_border_reflect101_u8(_p(img), w, h, img.strides[0], _p(out), b, w + 2 * b)
    return out


def gauss_kernel_q8(ksize, sigma):
    q = np.zeros(ksize, np.int32)
    rc = lib().orc_gauss_kernel_q8(ksize, sigma, _p(q))
    assert rc == 0
    return q


def gauss_blur(img, ksize, sigma):
    img = _img(img)
    out = np.empty_like(img)
    lib().orc_gauss_blur_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), out.strides[0], ksize, sigma)
    return out


def pyrdown(img):
    img = _img(img)
    h, w = img.shape
    out = np.empty((h // 2, w // 2), np.uint8)
    lib().orc_pyrdown_u8(_p(img), w, h, img.strides[0], _p(out), out.strides[0])
    return out


def sobel3(img):
    img = _img(img)
    h, w = img.shape
    dx = np.empty((h, w), np.int16)
    dy = np.empty((h, w), np.int16)
    lib().orc_sobel3_s16(_p(img), w, h, img.strides[0], _p(dx), _p(dy))
    return dx, dy


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(float(y), float(x)))


def fast9(img, th):
    img = _img(img)
    h, w = img.shape
    cap = max(1, w * h)
    xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); sc = np.empty(cap, np.int32)
    n = lib().orc_fast9(_p(img), w, h, img.strides[0], th, _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def distribute_octree(xs, ys, resp, minX, maxX, minY, maxY, N):
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    resp = np.ascontiguousarray(resp, np.int32)
    out = np.empty(max(1, len(xs)), np.int32)
    n = lib().orc_distribute_octree(_p(xs), _p(ys), _p(resp), len(xs), minX, maxX, minY, maxY, N, _p(out), len(out))
    return out[:n].copy()


# ---------------- ORB extractor ----------------
class ORBextractor:
    """Oracle mirror of PL_SLAM::ORBextractor (include/ORBextractor.h:45-113)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST):
        self._h = lib().orc_orb_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_orb_destroy(self._h)
            self._h = None

    def features_per_level(self):
        return [lib().orc_orb_features_per_level(self._h, l) for l in range(self.nlevels)]

    def scale_factors(self):
        return [lib().orc_orb_scale_factor(self._h, l) for l in range(self.nlevels)]

    def umax(self):
        return [lib().orc_orb_umax(self._h, v) for v in range(16)]

    def __call__(self, image):
        img = _img(image)
        cap = self.nfeatures * 2 + 64 * self.nlevels
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().orc_orb_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
        assert n >= 0, "oracle keypoint capacity too small"
        return kps[:n].copy(), desc[:n].copy()

    def level_image(self, level):
        w, h = C.c_int(), C.c_int()
        assert lib().orc_orb_level_size(self._h, level, C.byref(w), C.byref(h)) == 0
        st = C.c_size_t()
        p = lib().orc_orb_level_image(self._h, level, C.byref(st))
        buf = (C.c_uint8 * (st.value * h.value)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(h.value, st.value)[:, :w.value].copy()

    def level_blurred(self, level):
        w, h = C.c_int(), C.c_int()
        assert lib().orc_orb_level_size(self._h, level, C.byref(w), C.byref(h)) == 0
        st = C.c_size_t()
        p = lib().orc_orb_level_blurred(self._h, level, C.byref(st))
        if not p:
            return None
        buf = (C.c_uint8 * (w.value * h.value)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(h.value, w.value).copy()

    def level_raw(self, level):
        n = lib().orc_orb_level_raw_count(self._h, level)
        xs = np.empty(max(n, 1), np.int32); ys = np.empty(max(n, 1), np.int32); rr = np.empty(max(n, 1), np.int32)
        if n:
            lib().orc_orb_level_raw(self._h, level, _p(xs), _p(ys), _p(rr))
        return xs[:n], ys[:n], rr[:n]

    def level_kept_count(self, level):
        return lib().orc_orb_level_kept_count(self._h, level)


def stereo_match(orbL, orbR, kL, dL, kR, dR, mb, mbf):
    """Frame::ComputeStereoMatches (src/Frame.cc:881-1055): orbL/orbR are oracle extractors that just processed the
    left/right image (their pyramids are read).  Returns (mvuRight, mvDepth)."""
    kL = np.ascontiguousarray(kL, KEYPOINT_DTYPE); kR = np.ascontiguousarray(kR, KEYPOINT_DTYPE)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    u = np.empty(len(kL), np.float32); z = np.empty(len(kL), np.float32)
    lib().orc_stereo_match(orbL._h, orbR._h, _p(kL), _p(dL), len(kL), _p(kR), _p(dR), len(kR), mb, mbf, _p(u), _p(z))
    return u, z


def synth_vocabulary(k, L, seed, dup=True):
    """A random k-ary vocabulary tree of depth L in loadFromTextFile node order (breadth first): parent, descriptors,
    weights (some zero = stopped words), leaf flags.  With dup, some siblings share a descriptor (first-minimum ties)."""
    rng = np.random.default_rng(seed)
    parent, leaf = [0], [0]
    level_nodes = [0]
    for lv in range(1, L + 1):
        nxt = []
        for p in level_nodes:
            for _ in range(k):
                parent.append(p); leaf.append(1 if lv == L else 0); nxt.append(len(parent) - 1)
        level_nodes = nxt
    n = len(parent)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    if dup:
        for i in range(2, n, 7):
            if parent[i] == parent[i - 1]:
                desc[i] = desc[i - 1]
    weight = rng.uniform(0.0, 9.0, n)
    weight[rng.random(n) < 0.05] = 0.0
    weight[np.array(leaf) == 0] = 0.0
    return (np.array(parent, np.int32), np.ascontiguousarray(desc), np.ascontiguousarray(weight, np.float64), np.array(leaf, np.uint8))


def bow_transform(vocab, L, feats, levelsup):
    """Per-feature tree descent (TemplatedVocabulary.h:1218-1258) -> (word, weight, node)."""
    parent, desc, weight, leaf = vocab
    feats = np.ascontiguousarray(feats, np.uint8)
    n = len(feats)
    w = np.zeros(n, np.int32); wt = np.zeros(n, np.float64); nd = np.zeros(n, np.int32)
    lib().orc_bow_transform(L, len(parent), _p(parent), _p(desc), _p(weight), _p(leaf), _p(feats), n, levelsup, _p(w), _p(wt), _p(nd))
    return w, wt, nd


def bow_vectors(word, weight, node):
    """The BowVector / FeatureVector maps of transform(features, v, fv, levelsup) for TF_IDF + L1 (TemplatedVocabulary.h:1145-1190,
    BowVector.cpp:34-46, :62-84, FeatureVector.cpp:31-45): ordered dicts word -> value and node -> feature indices."""
    v, fv = {}, {}
    for i in range(len(word)):
        if weight[i] > 0:
            v[int(word[i])] = v.get(int(word[i]), 0.0) + float(weight[i])
            fv.setdefault(int(node[i]), []).append(i)
    norm = 0.0
    for k in sorted(v):
        norm += abs(v[k])
    if norm > 0.0:
        for k in v:
            v[k] /= norm
    return dict(sorted(v.items())), dict(sorted(fv.items()))


def undistort_points(pts, fx, fy, cx, cy, dist):
    """cv::undistortPoints(pts, pts, K, D, Mat(), K) on N x 2 float points (src/Frame.cc:750, :785, :814-815)."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    cam = np.array([fx, fy, cx, cy], np.float32); kd = np.ascontiguousarray(dist, np.float32)
    out = np.empty_like(pts)
    lib().orc_undistort_points(_p(cam), _p(kd), len(kd), _p(pts), len(pts), _p(out))
    return out


def undistort_keypoints(kps, fx, fy, cx, cy, dist):
    """Frame::UndistortKeyPoints (src/Frame.cc:733-763)."""
    out = np.ascontiguousarray(kps, KEYPOINT_DTYPE).copy()
    if np.float32(dist[0]) != 0:
        p = undistort_points(np.stack([out["x"], out["y"]], 1), fx, fy, cx, cy, dist)
        out["x"], out["y"] = p[:, 0], p[:, 1]
    return out


def undistort_keylines(kls, mids, fx, fy, cx, cy, dist):
    """Frame::UndistortKeyLines (src/Frame.cc:766-826): mid-points, start points, end points."""
    ok = np.ascontiguousarray(kls, KEYLINE_DTYPE).copy()
    om = undistort_keypoints(mids, fx, fy, cx, cy, dist)
    if np.float32(dist[0]) != 0:
        a = undistort_points(np.stack([ok["startPointX"], ok["startPointY"]], 1), fx, fy, cx, cy, dist)
        b = undistort_points(np.stack([ok["endPointX"], ok["endPointY"]], 1), fx, fy, cx, cy, dist)
        ok["startPointX"], ok["startPointY"], ok["endPointX"], ok["endPointY"] = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    return ok, om


class GridParams(C.Structure):
    _fields_ = [("cols", C.c_int32), ("rows", C.c_int32), ("min_x", C.c_float), ("min_y", C.c_float),
                ("inv_w", C.c_float), ("inv_h", C.c_float)]


def grid_params(cols, rows, min_x, max_x, min_y, max_y):
    """mfGridElementWidthInv = cols / (mnMaxX - mnMinX) in float (src/Frame.cc:139-140)."""
    f = np.float32
    return GridParams(cols, rows, f(min_x), f(min_y), f(cols) / (f(max_x) - f(min_x)), f(rows) / (f(max_y) - f(min_y)))


def grid_candidates(kps, g, qx, qy, qr, qminl=None, qmaxl=None, keylines=None):
    """Frame::AssignFeaturesToGrid[Lines] + GetFeaturesInArea[Lines] (src/Frame.cc:365-399, :562-722) -> (cand_off, cand_idx)."""
    kps = np.ascontiguousarray(kps, KEYPOINT_DTYPE)
    kls = None if keylines is None else np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    qx = np.ascontiguousarray(qx, np.float32); qy = np.ascontiguousarray(qy, np.float32); qr = np.ascontiguousarray(qr, np.float32)
    mn = None if qminl is None else np.ascontiguousarray(qminl, np.int32)
    mx = None if qmaxl is None else np.ascontiguousarray(qmaxl, np.int32)
    nq = len(qx)
    off = np.zeros(nq + 1, np.int32)
    cap = max(1, nq * max(len(kps), 1))
    idx = np.zeros(cap, np.int32)
    tot = lib().orc_grid_candidates(_p(kps), None if kls is None else _p(kls), len(kps), C.byref(g), _p(qx), _p(qy), _p(qr),
                                    None if mn is None else _p(mn), None if mx is None else _p(mx), nq, _p(off), _p(idx), cap)
    return off, idx[:tot].copy()


# ---------------- lines ----------------
def line_params(nfeatures=600, nlevels=2, refine=0, scale=1.1, sigma_scale=0.6, quant=2.2, ang_th=12.5,
                log_eps=1.0, density_th=0.6, n_bins=1024, min_line_length=0.0):
    return LineParams(nfeatures, nlevels, refine, scale, sigma_scale, quant, ang_th, log_eps, density_th,
                      n_bins, min_line_length)


def lsd_detect(img, scale=1.1, sigma_scale=0.6, quant=2.2, ang_th=12.5, n_bins=1024):
    img = _img(img)
    cap = 1 << 16
    lines = np.empty((cap, 4), np.float32)
    n = lib().orc_lsd_detect(_p(img), img.shape[1], img.shape[0], img.strides[0], scale, sigma_scale, quant,
                             ang_th, n_bins, _p(lines), cap)
    return lines[:n].copy()


def lsd_detect_keylines(params, img):
    img = _img(img)
    cap = 1 << 16
    kl = np.zeros(cap, KEYLINE_DTYPE)
    n = lib().orc_lsd_detect_keylines(C.byref(params), _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), cap)
    return kl[:n].copy()


def lbd_compute(img, keylines, want_float=False):
    img = _img(img)
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    n = len(kl)
    desc = np.zeros((n, 32), np.uint8)
    fdesc = np.zeros((n, 72), np.float32)
    lib().orc_lbd_compute(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), n, _p(desc), _p(fdesc))
    return (desc, fdesc) if want_float else desc


def line_extract(params, img):
    """Oracle mirror of Lineextractor::ComputeLsdWithLbd (src/Lineextractor.cc:112-212)."""
    img = _img(img)
    cap = max(16, params.nfeatures * 2 + 16)
    kl = np.zeros(cap, KEYLINE_DTYPE)
    mid = np.zeros(cap, KEYPOINT_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = lib().orc_line_extract(C.byref(params), _p(img), img.shape[1], img.shape[0], img.strides[0],
                               _p(kl), _p(mid), _p(desc), cap)
    assert n >= 0
    return kl[:n].copy(), mid[:n].copy(), desc[:n].copy()


def features_per_level_lines(params):
    return [lib().orc_line_features_per_level(C.byref(params), l) for l in range(params.nlevels)]


def std_sort_desc(keys):
    """libstdc++ std::sort with comparator k[a] > k[b] (unstable) as a permutation."""
    k = np.ascontiguousarray(keys, np.float32)
    p = np.zeros(max(len(k), 1), np.int32)
    lib().orc_std_sort_desc(_p(k), len(k), _p(p))
    return p[:len(k)]


# ---------------- matching ----------------
def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_p(a), _p(b))


def knn2(q, t):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), 2), np.int32); dist = np.empty((len(q), 2), np.int32)
    lib().orc_knn2(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist))
    return idx, dist


def candidates_top2(q, t, cand_lists):
    """Sequential top-2 over candidate lists (src/ORBmatcher.cc:430-456) -> (best_idx, best_dist, cand_dist)."""
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    off = np.zeros(len(q) + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in cand_lists])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32) for c in cand_lists]), np.int32)
    bi = np.empty((len(q), 2), np.int32); bd = np.empty((len(q), 2), np.int32); cd = np.empty(max(len(flat), 1), np.int32)
    lib().orc_hamming_candidates(_p(q), len(q), _p(t), _p(off), _p(flat), _p(bi), _p(bd), _p(cd))
    return bi, bd, cd[:len(flat)]


def match_nnr(q, t, nnr):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    m = np.empty(len(q), np.int32)
    n = lib().orc_match_nnr(_p(q), len(q), _p(t), len(t), nnr, _p(m))
    return m, n


# ---------------- synthetic images ----------------
def synth_image(w, h, seed):
    """uniform u8 noise -> Gaussian s=2 -> K filled rectangles -> 3x3 s=0.8 blur (pure numpy + oracle blur)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    img = gauss_blur(img, 13, 2.0)
    K = int(round(60.0 * (w * h) / (640.0 * 480.0)))
    for _ in range(K):
        x0 = int(rng.integers(0, w - 8)); y0 = int(rng.integers(0, h - 8))
        rw = int(rng.integers(8, max(9, w // 4))); rh = int(rng.integers(8, max(9, h // 4)))
        g = int(rng.integers(0, 256))
        img[y0:min(h, y0 + rh), x0:min(w, x0 + rw)] = g
    img = gauss_blur(img, 3, 0.8)
    return img
