"""Host-side mirror of the reference's extractor / matcher interfaces over the C ABI (include/plf.h).

The classes keep the reference's names, argument meaning and error behaviour:
  ORBextractor   -- include/ORBextractor.h:45-113   (operator(), scale getters, mvImagePyramid)
  Lineextractor  -- include/Lineextractor.h:44-181  (ComputeLsdWithLbd, scale getters)
  Linematcher    -- include/Linematcher.h:36-84     (DescriptorDistance, matchNNR, KNN mutual step)
numpy arrays stand in for cv::Mat / std::vector<cv::KeyPoint> / std::vector<KeyLine> with the same
memory layouts (28-byte KeyPoint, 68-byte KeyLine, N x 32 CV_8U descriptors).

There is no CPU fallback: if libplf.so is missing or no CUDA device is usable every call raises.
"""
import ctypes as C
import weakref
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libplf.so")

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
KEYLINE_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"),
                          ("pt_x", "<f4"), ("pt_y", "<f4"), ("response", "<f4"), ("size", "<f4"),
                          ("startPointX", "<f4"), ("startPointY", "<f4"),
                          ("endPointX", "<f4"), ("endPointY", "<f4"),
                          ("sPointInOctaveX", "<f4"), ("sPointInOctaveY", "<f4"),
                          ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
                          ("lineLength", "<f4"), ("numOfPixels", "<i4")])

PLF_OK, PLF_ERR_INVALID, PLF_ERR_CUDA, PLF_ERR_CAPACITY, PLF_ERR_STATE = range(5)


class PlfError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("plf status %d: %s" % (status, msg))
        self.status = status


class GridParams(C.Structure):
    """plf_grid_params: the Frame grid (64 x 48 for points, 16 x 12 for lines in the reference)."""
    _fields_ = [("cols", C.c_int32), ("rows", C.c_int32), ("min_x", C.c_float), ("min_y", C.c_float),
                ("inv_w", C.c_float), ("inv_h", C.c_float)]

    @classmethod
    def for_image(cls, cols, rows, min_x, max_x, min_y, max_y):
        f = np.float32
        return cls(cols, rows, f(min_x), f(min_y), f(cols) / (f(max_x) - f(min_x)), f(rows) / (f(max_y) - f(min_y)))


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int)]


class LineParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("nlevels", C.c_int), ("refine", C.c_int),
                ("scale", C.c_double), ("sigma_scale", C.c_double), ("quant", C.c_double),
                ("ang_th", C.c_double), ("log_eps", C.c_double), ("density_th", C.c_double),
                ("n_bins", C.c_int), ("min_line_length", C.c_double)]


class FldParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("nlevels", C.c_int), ("scale", C.c_double), ("length_threshold", C.c_int),
                ("distance_threshold", C.c_float), ("canny_th1", C.c_double), ("canny_th2", C.c_double),
                ("canny_aperture_size", C.c_int), ("do_merge", C.c_int)]


_libs = {}


def load(path=None):
    """Load the C-ABI library (libplf.so by default) and declare its prototypes."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise PlfError(PLF_ERR_CUDA, "C-ABI library %s is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)" % path)
    L = C.CDLL(path)
    vp, i32p, f32p = C.c_void_p, C.c_void_p, C.c_void_p
    P = C.POINTER
    sig = {
        "plf_ctx_create": (C.c_int, [C.c_int, P(vp)]),
        "plf_ctx_create_prio": (C.c_int, [C.c_int, C.c_int, P(vp)]),
        "plf_ctx_destroy": (None, [vp]),
        "plf_last_error": (C.c_char_p, [vp]),
        "plf_ctx_synchronize": (C.c_int, [vp]),
        "plf_ctx_stream": (vp, [vp]),
        "plf_timer_start": (C.c_int, [vp]),
        "plf_timer_stop": (C.c_int, [vp, P(C.c_float)]),
        "plf_ctx_launch_count": (C.c_uint64, [vp]),
        "plf_ctx_wait": (C.c_int, [vp, vp]),
        "plf_profile_enable": (C.c_int, [vp, C.c_int]),
        "plf_profile_report": (C.c_int, [vp, C.c_char_p, C.c_size_t]),
        "plf_profile_timeline": (C.c_int, [vp, vp, C.c_char_p, C.c_size_t]),
        "plf_orb_create": (C.c_int, [vp, P(OrbParams), P(vp)]),
        "plf_orb_destroy": (None, [vp]),
        "plf_orb_tables": (C.c_int, [vp, f32p, f32p, f32p, f32p, i32p]),
        "plf_orb_max_keypoints": (C.c_int, [vp]),
        "plf_orb_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, vp, C.c_int, P(C.c_int)]),
        "plf_orb_extract_batch": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, C.c_int, i32p]),
        "plf_orb_extract_batch_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, C.c_int, vp]),
        "plf_orb_pyramid_level": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t, P(C.c_int), P(C.c_int)]),
        "plf_orb_debug_blurred": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t]),
        "plf_orb_debug_raw_keys": (C.c_int, [vp, C.c_int, C.c_int, i32p, i32p, i32p, C.c_int, P(C.c_int)]),
        "plf_orb_distribute_octree": (C.c_int, [vp, i32p, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int, P(C.c_int)]),
        "plf_line_create": (C.c_int, [vp, P(LineParams), P(vp)]),
        "plf_line_destroy": (None, [vp]),
        "plf_line_tables": (C.c_int, [vp, f32p, f32p, f32p, f32p, i32p]),
        "plf_line_max_keylines": (C.c_int, [vp]),
        "plf_line_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, vp, vp, C.c_int, P(C.c_int)]),
        "plf_line_extract_batch": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int, i32p]),
        "plf_line_extract_batch_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int, vp]),
        "plf_lsd_detect": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, P(C.c_int)]),
        "plf_lbd_compute": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp, vp]),
        "plf_descriptor_distance": (C.c_int, [vp, vp, vp, C.c_int, i32p]),
        "plf_hamming_knn2": (C.c_int, [vp, vp, C.c_int, vp, C.c_int64, i32p, i32p]),
        "plf_hamming_knn2_device": (C.c_int, [vp, vp, C.c_int, vp, C.c_int64, C.c_int64, vp, vp]),
        "plf_knn2_merge_device": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp]),
        "plf_upload": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "plf_device_malloc": (C.c_int, [vp, C.c_size_t, P(vp)]),
        "plf_device_free": (None, [vp, vp]),
        "plf_orb_extract_batch_from_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, C.c_int, vp]),
        "plf_line_extract_batch_from_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int, vp]),
        "plf_fld_create": (C.c_int, [vp, P(FldParams), P(vp)]),
        "plf_fld_destroy": (None, [vp]),
        "plf_fld_features_per_level": (C.c_int, [vp, i32p]),
        "plf_fld_detect": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, P(C.c_int)]),
        "plf_fld_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, vp, vp, C.c_int, P(C.c_int)]),
        "plf_comm_unique_id": (C.c_int, [vp]),
        "plf_comm_create": (C.c_int, [vp, vp, C.c_int, C.c_int, P(vp)]),
        "plf_comm_destroy": (None, [vp]),
        "plf_comm_rank": (C.c_int, [vp]),
        "plf_comm_world": (C.c_int, [vp]),
        "plf_hamming_knn2_sharded_device": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_int64, C.c_int64, vp, vp]),
        "plf_match_nnr_sharded_device": (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_int64, C.c_int64, C.c_float, vp, vp, vp, vp]),
        "plf_match_nnr": (C.c_int, [vp, vp, C.c_int, vp, C.c_int64, C.c_float, i32p, P(C.c_int)]),
        "plf_nnr_from_knn2_device": (C.c_int, [vp, vp, vp, C.c_int, C.c_float, vp, vp]),
        "plf_popc_peak": (C.c_int, [vp, P(C.c_double)]),
        "plf_match_nnr_mutual": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, C.c_float, i32p, P(C.c_int)]),
        "plf_hamming_candidates": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, i32p, i32p, i32p, i32p, i32p]),
        "plf_hamming_candidates_device": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp]),
        "plf_grid_build_device": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]),
        "plf_grid_query_device": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, P(C.c_int)]),
        "plf_grid_candidates": (C.c_int, [vp, vp, vp, C.c_int, vp, f32p, f32p, f32p, i32p, i32p, C.c_int, i32p, i32p, C.c_int, P(C.c_int)]),
        "plf_undistort_keypoints": (C.c_int, [vp, vp, vp, C.c_int, vp]),
        "plf_undistort_keypoints_device": (C.c_int, [vp, vp, vp, C.c_int, vp]),
        "plf_undistort_keylines": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp]),
        "plf_undistort_keylines_device": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp]),
        "plf_vocab_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, vp, vp, vp, P(vp)]),
        "plf_vocab_load_text": (C.c_int, [vp, C.c_char_p, P(vp)]),
        "plf_vocab_destroy": (None, [vp]),
        "plf_vocab_info": (C.c_int, [vp, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int)]),
        "plf_bow_transform": (C.c_int, [vp, vp, C.c_int, C.c_int, i32p, vp, i32p]),
        "plf_bow_transform_device": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp]),
        "plf_stereo_match": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_float, f32p, f32p]),
        "plf_stereo_match_batch_device": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                                    C.c_int, C.c_float, C.c_float, vp, vp]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    L._plf_missing = missing
    L._plf_symbols = list(sig)
    _libs[path] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """plf_ctx: one per host thread / stream (the reference's concurrent std::threads each get one)."""

    def __init__(self, device=0, lib=None, priority=0):
        self.lib = load(lib)
        h = C.c_void_p()
        st = self.lib.plf_ctx_create_prio(device, priority, C.byref(h))
        if st != PLF_OK:
            raise PlfError(st, "plf_ctx_create(device=%d) failed: no usable CUDA device (there is no CPU fallback)" % device)
        self.h = h
        self.device = device

    def check(self, st):
        if st != PLF_OK:
            raise PlfError(st, (self.lib.plf_last_error(self.h) or b"").decode())

    def synchronize(self):
        self.check(self.lib.plf_ctx_synchronize(self.h))

    def timer_start(self):
        self.check(self.lib.plf_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.check(self.lib.plf_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.plf_ctx_launch_count(self.h))

    def stream(self):
        return self.lib.plf_ctx_stream(self.h)

    def popc_peak(self):
        v = C.c_double()
        self.check(self.lib.plf_popc_peak(self.h, C.byref(v)))
        return v.value

    def wait(self, other):
        self.check(self.lib.plf_ctx_wait(self.h, other.h))

    def profile_enable(self, on=True):
        self.check(self.lib.plf_profile_enable(self.h, 1 if on else 0))

    def profile_timeline(self, ref):
        buf = C.create_string_buffer(1 << 22)
        self.check(self.lib.plf_profile_timeline(self.h, ref.h, buf, len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            name, t0, t1 = line.rsplit(" ", 2)
            out.append((name, float(t0), float(t1)))
        return out

    def profile_report(self):
        """{kernel name: (total_ms, launches)} accumulated since profile_enable(True)."""
        buf = C.create_string_buffer(1 << 16)
        self.check(self.lib.plf_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out

    def _adopt(self, child):
        """Extractors register here so that the context is never destroyed before the objects that use its stream
        (garbage collection at interpreter exit finalises objects in no particular order)."""
        if not hasattr(self, "_children"):
            self._children = weakref.WeakSet()
        self._children.add(child)

    def close(self):
        for ch in list(getattr(self, "_children", ())):
            ch.close()
        if getattr(self, "h", None):
            self.lib.plf_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_image(image):
    if image is None:
        return None
    a = np.asarray(image)
    if a.size == 0:
        return None
    if a.dtype != np.uint8 or a.ndim != 2:
        raise AssertionError("image.type() == CV_8UC1")  # the reference asserts (src/ORBextractor.cc:1050)
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


class ORBextractor:
    """Mirror of PL_SLAM::ORBextractor (include/ORBextractor.h:45-113)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, ctx=None, device=0, lib=None):
        self.ctx = ctx or Context(device, lib)
        self.lib = self.ctx.lib
        self.params = OrbParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        h = C.c_void_p()
        self.ctx.check(self.lib.plf_orb_create(self.ctx.h, C.byref(self.params), C.byref(h)))
        self.h = h
        self.ctx._adopt(self)
        self.nlevels = nlevels
        self.scaleFactor = scaleFactor
        n = nlevels
        self._scale = np.zeros(n, np.float32); self._inv = np.zeros(n, np.float32)
        self._sig = np.zeros(n, np.float32); self._isig = np.zeros(n, np.float32)
        self._per = np.zeros(n, np.int32)
        self.ctx.check(self.lib.plf_orb_tables(self.h, _p(self._scale), _p(self._inv), _p(self._sig), _p(self._isig), _p(self._per)))
        self.max_keypoints = self.lib.plf_orb_max_keypoints(self.h)
        self._nframes = 0

    # getters, include/ORBextractor.h:63-83
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv.copy()
    def GetScaleSigmaSquares(self): return self._sig.copy()
    def GetInverseScaleSigmaSquares(self): return self._isig.copy()
    def features_per_level(self): return self._per.copy()

    def __call__(self, image, mask=None):
        """operator()(image, mask, keypoints, descriptors): returns (keypoints[28 B records], descriptors N x 32).
        Empty image -> empty outputs (silent return, src/ORBextractor.cc:1046-1047); mask is ignored."""
        img = _as_image(image)
        if img is None:
            return np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        cap = self.max_keypoints
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int()
        self.ctx.check(self.lib.plf_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                                _p(kps), _p(desc), cap, C.byref(n)))
        self._nframes = 1
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        """Batched operator() over a (B, H, W) uint8 array: returns lists of keypoints / descriptors."""
        imgs = np.ascontiguousarray(images, np.uint8)
        assert imgs.ndim == 3
        B, H, W = imgs.shape
        cap = self.max_keypoints
        kps = np.zeros((B, cap), KEYPOINT_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        self.ctx.check(self.lib.plf_orb_extract_batch(self.h, _p(imgs), B, W, H, imgs.strides[1], imgs.strides[0],
                                                      _p(kps), _p(desc), cap, _p(n)))
        self._nframes = B
        return [kps[b, :n[b]].copy() for b in range(B)], [desc[b, :n[b]].copy() for b in range(B)]

    def pyramid_level(self, level, frame=0):
        """mvImagePyramid[level] (include/ORBextractor.h:85) of the last call, as an (h, w) uint8 array."""
        w, h = C.c_int(), C.c_int()
        self.ctx.check(self.lib.plf_orb_pyramid_level(self.h, frame, level, None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        self.ctx.check(self.lib.plf_orb_pyramid_level(self.h, frame, level, _p(out), out.strides[0], C.byref(w), C.byref(h)))
        return out

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    def ComputeStereoMatches(self, right, keysL, descL, keysR, descR, mb, mbf, frame_l=0, frame_r=0):
        """Frame::ComputeStereoMatches (src/Frame.cc:881-1055) for the images this extractor (left) and `right`
        processed last: -> (mvuRight, mvDepth)."""
        kl = np.ascontiguousarray(keysL, KEYPOINT_DTYPE); kr = np.ascontiguousarray(keysR, KEYPOINT_DTYPE)
        dl = np.ascontiguousarray(descL, np.uint8); dr = np.ascontiguousarray(descR, np.uint8)
        u = np.full(len(kl), -1.0, np.float32); z = np.full(len(kl), -1.0, np.float32)
        self.ctx.check(self.lib.plf_stereo_match(self.h, frame_l, right.h, frame_r, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr),
                                                 mb, mbf, _p(u), _p(z)))
        return u, z

    def debug_blurred(self, level, frame=0):
        w, h = C.c_int(), C.c_int()
        self.ctx.check(self.lib.plf_orb_pyramid_level(self.h, frame, level, None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        self.ctx.check(self.lib.plf_orb_debug_blurred(self.h, frame, level, _p(out), out.strides[0]))
        return out

    def debug_raw_keys(self, level, frame=0, cap=1 << 18):
        xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); rr = np.empty(cap, np.int32)
        n = C.c_int()
        self.ctx.check(self.lib.plf_orb_debug_raw_keys(self.h, frame, level, _p(xs), _p(ys), _p(rr), cap, C.byref(n)))
        return xs[:n.value].copy(), ys[:n.value].copy(), rr[:n.value].copy()

    def close(self):
        if getattr(self, "h", None):
            # finalisation order at interpreter exit is arbitrary: never touch an extractor whose context is gone
            if getattr(self.ctx, "h", None):
                self.lib.plf_orb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def distribute_octree(ctx, xs, ys, resp, minX, maxX, minY, maxY, N):
    """ORBextractor::DistributeOctTree (src/ORBextractor.cc:539-763) on the GPU for one level."""
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    resp = np.ascontiguousarray(resp, np.int32)
    out = np.empty(max(len(xs), 1), np.int32)
    n = C.c_int()
    ctx.check(ctx.lib.plf_orb_distribute_octree(ctx.h, _p(xs), _p(ys), _p(resp), len(xs), minX, maxX, minY, maxY, N,
                                                _p(out), len(out), C.byref(n)))
    return out[:n.value].copy()


class Lineextractor:
    """Mirror of PL_SLAM::Lineextractor, LSD constructor (include/Lineextractor.h:49-51)."""

    def __init__(self, nfeatures=240, nlevels=3, refine=0, scale=1.05, sigma_scale=0.6, quant=2.0, ang_th=22.5,
                 log_eps=1.0, density_th=0.7, n_bins=1024, min_line_length=32.0, busingLSD=True,
                 ctx=None, device=0, lib=None):
        if not busingLSD:
            raise NotImplementedError("FLD branch (System.usingLsdFeature: 0) is outside the hot path (SURVEY.md 8f rank 4)")
        self.ctx = ctx or Context(device, lib)
        self.lib = self.ctx.lib
        self.busingLSD = True
        self.params = LineParams(nfeatures, nlevels, refine, scale, sigma_scale, quant, ang_th, log_eps, density_th,
                                 n_bins, min_line_length)
        h = C.c_void_p()
        self.ctx.check(self.lib.plf_line_create(self.ctx.h, C.byref(self.params), C.byref(h)))
        self.h = h
        self.ctx._adopt(self)
        self.nlevels = nlevels
        self.scale = scale
        n = nlevels
        self._scale = np.zeros(n, np.float32); self._inv = np.zeros(n, np.float32)
        self._sig = np.zeros(n, np.float32); self._isig = np.zeros(n, np.float32)
        self._per = np.zeros(n, np.int32)
        self.ctx.check(self.lib.plf_line_tables(self.h, _p(self._scale), _p(self._inv), _p(self._sig), _p(self._isig), _p(self._per)))
        self.max_keylines = self.lib.plf_line_max_keylines(self.h)

    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return np.float32(self.scale)
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv.copy()
    def GetScaleSigmaSquares(self): return self._sig.copy()
    def GetInverseScaleSigmaSquares(self): return self._isig.copy()
    def features_per_level(self): return self._per.copy()

    def ComputeLsdWithLbd(self, image):
        """ComputeLsdWithLbd(image, keyLines, keypoints, descriptors) (src/Lineextractor.cc:112-212):
        returns (keyLines[68 B records], mid-point keypoints, descriptors NL x 32).  Empty image -> empties."""
        img = _as_image(image)
        if img is None:
            return np.zeros(0, KEYLINE_DTYPE), np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        cap = self.max_keylines
        kl = np.zeros(cap, KEYLINE_DTYPE); mid = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int()
        self.ctx.check(self.lib.plf_line_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                                 _p(kl), _p(mid), _p(desc), cap, C.byref(n)))
        return kl[:n.value].copy(), mid[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        imgs = np.ascontiguousarray(images, np.uint8)
        B, H, W = imgs.shape
        cap = self.max_keylines
        kl = np.zeros((B, cap), KEYLINE_DTYPE); mid = np.zeros((B, cap), KEYPOINT_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8); n = np.zeros(B, np.int32)
        self.ctx.check(self.lib.plf_line_extract_batch(self.h, _p(imgs), B, W, H, imgs.strides[1], imgs.strides[0],
                                                       _p(kl), _p(mid), _p(desc), cap, _p(n)))
        return ([kl[b, :n[b]].copy() for b in range(B)], [mid[b, :n[b]].copy() for b in range(B)],
                [desc[b, :n[b]].copy() for b in range(B)])

    def lsd_detect(self, image, cap=1 << 16):
        """LSDDetectorC::detect(image, keylines, 2, nlevels, opts) alone."""
        img = _as_image(image)
        kl = np.zeros(cap, KEYLINE_DTYPE)
        n = C.c_int()
        self.ctx.check(self.lib.plf_lsd_detect(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), cap, C.byref(n)))
        return kl[:n.value].copy()

    def lbd_compute(self, image, keylines, want_float=False):
        """BinaryDescriptor::compute(image, keylines, descriptors[, returnFloatDescr])."""
        img = _as_image(image)
        kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
        if len(kl) == 0:
            # the reference prints "Error: keypoint list is empty" and returns (binary_descriptor_custom.cpp:556-560)
            return (np.zeros((0, 32), np.uint8), np.zeros((0, 72), np.float32)) if want_float else np.zeros((0, 32), np.uint8)
        desc = np.zeros((len(kl), 32), np.uint8)
        fd = np.zeros((len(kl), 72), np.float32)
        self.ctx.check(self.lib.plf_lbd_compute(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), len(kl), _p(desc), _p(fd)))
        return (desc, fd) if want_float else desc

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):
                self.lib.plf_line_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FldLineextractor:
    """Mirror of PL_SLAM::Lineextractor built with its FLD constructor (include/Lineextractor.h:55-57; System.usingLsdFeature: 0)."""

    def __init__(self, nfeatures=240, nlevels=1, scale=1.05, length_threshold=10, distance_threshold=1.414213562, canny_th1=50.0,
                 canny_th2=100.0, canny_aperture_size=3, do_merge=False, busingLSD=False, ctx=None, device=0, lib=None):
        self.ctx = ctx or Context(device, lib)
        self.lib = self.ctx.lib
        self.prm = FldParams(nfeatures, nlevels, scale, length_threshold, distance_threshold, canny_th1, canny_th2, canny_aperture_size,
                             1 if do_merge else 0)
        h = C.c_void_p()
        self.ctx.check(self.lib.plf_fld_create(self.ctx.h, C.byref(self.prm), C.byref(h)))
        self.h = h
        self.nfeatures, self.nlevels, self.busingLSD = nfeatures, nlevels, False
        self.ctx._adopt(self)

    def features_per_level(self):
        per = np.zeros(self.nlevels, np.int32)
        self.ctx.check(self.lib.plf_fld_features_per_level(self.h, _p(per)))
        return per

    def detect(self, image, cap=1 << 14):
        """Lineextractor::detect(image, lines): n x 4 floats."""
        img = _as_image(image)
        lines = np.zeros((cap, 4), np.float32)
        n = C.c_int()
        self.ctx.check(self.lib.plf_fld_detect(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(lines), cap, C.byref(n)))
        return lines[:n.value].copy()

    def ComputeFldWithLbd(self, image):
        """ComputeFldWithLbd(image, keyLines, keypoints, descriptors) (src/Lineextractor.cc:242-336)."""
        img = _as_image(image)
        if img is None:
            return np.zeros(0, KEYLINE_DTYPE), np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        cap = self.nfeatures * 2 + 16
        kl = np.zeros(cap, KEYLINE_DTYPE); mid = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int()
        self.ctx.check(self.lib.plf_fld_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kl), _p(mid), _p(desc), cap, C.byref(n)))
        return kl[:n.value].copy(), mid[:n.value].copy(), desc[:n.value].copy()

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):
                self.lib.plf_fld_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _desc(a):
    a = np.ascontiguousarray(a, np.uint8)
    if a.ndim != 2 or a.shape[1] != 32:
        raise ValueError("descriptors must be N x 32 CV_8U")
    return a


class Linematcher:
    """Mirror of the brute-force pieces of PL_SLAM::Linematcher (include/Linematcher.h:36-84); the same
    DescriptorDistance serves ORBmatcher (include/ORBmatcher.h:44)."""

    def __init__(self, nnratio=0.6, checkOri=True, checklen=False, lengtherr=0.1, ctx=None, device=0, lib=None):
        self.ctx = ctx or Context(device, lib)
        self.lib = self.ctx.lib
        self.mfNNratio = nnratio

    def DescriptorDistance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        d = np.empty(len(a), np.int32)
        self.ctx.check(self.lib.plf_descriptor_distance(self.ctx.h, _p(a), _p(b), len(a), _p(d)))
        return int(d[0]) if len(d) == 1 else d

    def knnMatch2(self, desc1, desc2):
        q, t = _desc(desc1), _desc(desc2)
        idx = np.empty((len(q), 2), np.int32); dist = np.empty((len(q), 2), np.int32)
        self.ctx.check(self.lib.plf_hamming_knn2(self.ctx.h, _p(q), len(q), _p(t), len(t), _p(idx), _p(dist)))
        return idx, dist

    def matchNNR(self, desc1, desc2, nnr=None):
        """matchNNR(desc1, desc2, nnr, matches_12, nmatches) (src/Linematcher.cc:520-541) -> (matches_12, nmatches)."""
        q, t = _desc(desc1), _desc(desc2)
        m = np.empty(len(q), np.int32)
        n = C.c_int()
        self.ctx.check(self.lib.plf_match_nnr(self.ctx.h, _p(q), len(q), _p(t), len(t),
                                              self.mfNNratio if nnr is None else nnr, _p(m), C.byref(n)))
        return m, n.value

    def matchNNRMutual(self, desc1, desc2, nnr=None):
        """Both matchNNR directions + the mutual-consistency filter of SearchByKNN (src/Linematcher.cc:454-471)."""
        a, b = _desc(desc1), _desc(desc2)
        m = np.empty(len(a), np.int32)
        n = C.c_int()
        self.ctx.check(self.lib.plf_match_nnr_mutual(self.ctx.h, _p(a), len(a), _p(b), len(b),
                                                     self.mfNNratio if nnr is None else nnr, _p(m), C.byref(n)))
        return m, n.value

    def candidates_top2(self, desc1, desc2, cand_lists, want_dist=False):
        """The distance / top-2 core of the candidate-list matchers (ORBmatcher::SearchForInitialization,
        src/ORBmatcher.cc:430-456, and the other Search* loops): cand_lists[q] = train rows query q scans, in
        order.  Returns (best_idx[nq,2], best_dist[nq,2]) (+ the flat per-candidate distances)."""
        q, t = _desc(desc1), _desc(desc2)
        off = np.zeros(len(q) + 1, np.int32)
        off[1:] = np.cumsum([len(c) for c in cand_lists])
        flat = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32) for c in cand_lists]) if len(cand_lists) else np.zeros(0, np.int32), np.int32)
        bi = np.empty((len(q), 2), np.int32); bd = np.empty((len(q), 2), np.int32)
        cd = np.empty(max(len(flat), 1), np.int32) if want_dist else None
        self.ctx.check(self.lib.plf_hamming_candidates(self.ctx.h, _p(q), len(q), _p(t), len(t), _p(off), _p(flat) if len(flat) else None,
                                                       _p(bi), _p(bd), _p(cd) if want_dist else None))
        return (bi, bd, cd[:len(flat)]) if want_dist else (bi, bd)


class Camera(C.Structure):
    """plf_camera: mK (fx, fy, cx, cy) and mDistCoef (k1, k2, p1, p2[, k3]) of a Frame."""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("k", C.c_float * 5), ("nk", C.c_int32)]

    @classmethod
    def make(cls, fx, fy, cx, cy, dist):
        k = (C.c_float * 5)(*([float(np.float32(v)) for v in dist] + [0.0] * (5 - len(dist))))
        return cls(fx, fy, cx, cy, k, len(dist))


def undistort_keypoints(ctx, cam, keys):
    """Frame::UndistortKeyPoints (src/Frame.cc:733-763)."""
    k = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    out = np.empty_like(k)
    ctx.check(ctx.lib.plf_undistort_keypoints(ctx.h, C.byref(cam), _p(k), len(k), _p(out)))
    return out


def undistort_keylines(ctx, cam, keylines, midpoints):
    """Frame::UndistortKeyLines (src/Frame.cc:766-826) -> (mvLinesUn, mvMidPointsUn)."""
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE); mid = np.ascontiguousarray(midpoints, KEYPOINT_DTYPE)
    okl = np.empty_like(kl); om = np.empty_like(mid)
    ctx.check(ctx.lib.plf_undistort_keylines(ctx.h, C.byref(cam), _p(kl), _p(mid), len(kl), _p(okl), _p(om)))
    return okl, om


class ORBVocabulary:
    """The device-resident DBoW2 vocabulary tree (ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB>,
    include/ORBVocabulary.h) and its transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1258)."""

    def __init__(self, ctx, k=None, L=None, parent=None, desc=None, weight=None, is_leaf=None, path=None, scoring=0, weighting=0):
        self.ctx = ctx
        self.lib = ctx.lib
        h = C.c_void_p()
        if path is not None:
            ctx.check(self.lib.plf_vocab_load_text(ctx.h, os.fsencode(path), C.byref(h)))
        else:
            parent = np.ascontiguousarray(parent, np.int32); desc = np.ascontiguousarray(desc, np.uint8)
            weight = np.ascontiguousarray(weight, np.float64); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
            ctx.check(self.lib.plf_vocab_create(ctx.h, k, L, scoring, weighting, len(parent), _p(parent), _p(desc), _p(weight), _p(is_leaf), C.byref(h)))
        self.h = h
        ctx._adopt(self)
        vals = [C.c_int() for _ in range(6)]
        self.lib.plf_vocab_info(self.h, *[C.byref(v) for v in vals])
        self.k, self.L, self.nnodes, self.nwords, self.scoring, self.weighting = [v.value for v in vals]

    def transform_features(self, descriptors, levelsup=4):
        """Per-feature (word id, weight, node id at level L - levelsup)."""
        d = _desc(descriptors)
        n = len(d)
        w = np.zeros(n, np.int32); wt = np.zeros(n, np.float64); nd = np.zeros(n, np.int32)
        self.ctx.check(self.lib.plf_bow_transform(self.h, _p(d), n, levelsup, _p(w), _p(wt), _p(nd)))
        return w, wt, nd

    def transform(self, descriptors, levelsup=4):
        """transform(features, BowVector, FeatureVector, levelsup) for TF_IDF / TF weighting + L1 norm: the ordered
        insertions (BowVector::addWeight, FeatureVector::addFeature, BowVector::normalize) replayed on the host."""
        w, wt, nd = self.transform_features(descriptors, levelsup)
        v, fv = {}, {}
        for i in range(len(w)):
            if wt[i] > 0:
                v[int(w[i])] = v.get(int(w[i]), 0.0) + float(wt[i])
                fv.setdefault(int(nd[i]), []).append(i)
        norm = 0.0
        for k in sorted(v):
            norm += abs(v[k])
        if norm > 0.0:
            for k in v:
                v[k] /= norm
        return dict(sorted(v.items())), dict(sorted(fv.items()))

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):
                self.lib.plf_vocab_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def features_in_area(ctx, keys, grid, qx, qy, qr, min_level=None, max_level=None, keylines=None, cand_cap=None):
    """Frame::AssignFeaturesToGrid[Lines] + GetFeaturesInArea[Lines] (src/Frame.cc:365-399, :562-722) for a set of queries:
    -> (cand_off[nq + 1], cand_idx) CSR candidate lists in the reference's order."""
    kps = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    kls = None if keylines is None else np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    qx = np.ascontiguousarray(qx, np.float32); qy = np.ascontiguousarray(qy, np.float32); qr = np.ascontiguousarray(qr, np.float32)
    mn = None if min_level is None else np.ascontiguousarray(min_level, np.int32)
    mx = None if max_level is None else np.ascontiguousarray(max_level, np.int32)
    nq = len(qx)
    cap = cand_cap if cand_cap is not None else max(1, nq * 64)
    while True:
        off = np.zeros(nq + 1, np.int32); idx = np.zeros(cap, np.int32)
        tot = C.c_int()
        st = ctx.lib.plf_grid_candidates(ctx.h, _p(kps), None if kls is None else _p(kls), len(kps), C.byref(grid), _p(qx), _p(qy), _p(qr),
                                         None if mn is None else _p(mn), None if mx is None else _p(mx), nq, _p(off), _p(idx), cap, C.byref(tot))
        if st == PLF_ERR_CAPACITY and cand_cap is None and tot.value > cap:
            cap = tot.value
            continue
        ctx.check(st)
        return off, idx[:tot.value].copy()


ORBmatcher = Linematcher  # ORBmatcher::DescriptorDistance is the same function (src/ORBmatcher.cc:1656-1672)
