"""Second, independent CPU restatement: the reference's glue logic in pure Python over the
same OpenCV primitives the reference calls (cv2 4.13.0).  TEST INFRASTRUCTURE ONLY.

Used by tests/golden/make_golden.py to produce the committed golden vectors and by the
CPU tests to cross-check the C oracle (oracle/*.c).  Follows src/ORBextractor.cc:410-853,
1034-1132, LSDDetector_custom.cpp:56-73/227-324, src/Lineextractor.cc:112-212 and
src/Linematcher.cc:520-541.  Same documented tie-break/float choices as the C oracle.
"""
import math
import numpy as np
import cv2


# the float overloads the reference's C++ picks (libm through ctypes: numpy may route float32 trig through SVML)
import ctypes as _C
_libm = _C.CDLL("libm.so.6")
for _n in ("cosf", "sinf"):
    getattr(_libm, _n).restype = _C.c_float
    getattr(_libm, _n).argtypes = [_C.c_float]
_libm.atan2f.restype = _C.c_float
_libm.atan2f.argtypes = [_C.c_float, _C.c_float]


def _cosf(x):
    return np.float32(_libm.cosf(float(x)))


def _sinf(x):
    return np.float32(_libm.sinf(float(x)))


def _atan2f(y, x):
    return np.float32(_libm.atan2f(float(y), float(x)))


EDGE = 19
HALF = 15
f32 = np.float32


def orb_tables(nfeatures, scaleFactor, nlevels):
    sf = float(f32(scaleFactor))  # float param stored in a double member
    scale = [f32(1.0)]
    for i in range(1, nlevels):
        scale.append(f32(float(scale[-1]) * sf))
    inv = [f32(1.0) / s for s in scale]
    factor = f32(1.0 / sf)
    nd = f32(f32(nfeatures) * (f32(1) - factor)) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    nd = f32(nd)
    per = []
    tot = 0
    for _ in range(nlevels - 1):
        v = int(np.rint(nd))
        per.append(v)
        tot += v
        nd = f32(nd * factor)
    per.append(max(nfeatures - tot, 0))
    umax = [0] * (HALF + 2)
    vmax = int(math.floor(HALF * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(HALF * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(math.sqrt(HALF * HALF - v * v)))
    v0 = 0
    for v in range(HALF, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return scale, inv, per, umax[:16]


def compute_pyramid(image, inv):
    levels = []
    for l, s in enumerate(inv):
        w = int(np.rint(f32(image.shape[1]) * s))
        h = int(np.rint(f32(image.shape[0]) * s))
        if l == 0:
            roi = image
        else:
            roi = cv2.resize(levels[l - 1][EDGE:-EDGE, EDGE:-EDGE], (w, h), interpolation=cv2.INTER_LINEAR)
        levels.append(cv2.copyMakeBorder(roi, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101))
    return levels


class _Node:
    __slots__ = ("ulx", "uly", "brx", "bry", "keys", "nomore", "cid")


def _divide(n, xs, ys, counter):
    halfX = int(math.ceil(float(f32(n.brx - n.ulx) / f32(2))))
    halfY = int(math.ceil(float(f32(n.bry - n.uly) / f32(2))))
    mx, my = n.ulx + halfX, n.uly + halfY
    ch = []
    for (a, b, c, d) in ((n.ulx, n.uly, mx, my), (mx, n.uly, n.brx, my), (n.ulx, my, mx, n.bry), (mx, my, n.brx, n.bry)):
        c_ = _Node()
        c_.ulx, c_.uly, c_.brx, c_.bry, c_.keys, c_.nomore = a, b, c, d, [], False
        c_.cid = counter[0]
        counter[0] += 1
        ch.append(c_)
    for k in n.keys:
        q = (0 if xs[k] < mx else 1) + (0 if ys[k] < my else 2)
        ch[q].keys.append(k)
    for c_ in ch:
        if len(c_.keys) == 1:
            c_.nomore = True
    return ch


def distribute_octree(xs, ys, resp, minX, maxX, minY, maxY, N):
    """ORBextractor::DistributeOctTree, src/ORBextractor.cc:539-763 (python list emulation)."""
    nIni = max(1, int(np.round(float(f32(maxX - minX) / f32(maxY - minY)))))
    # C++ round() is half away from zero; values here are positive and np.round is half-even:
    r = float(f32(maxX - minX) / f32(maxY - minY))
    nIni = max(1, int(math.floor(r + 0.5)))
    hX = f32(maxX - minX) / f32(nIni)
    counter = [0]
    ini = []
    for i in range(nIni):
        n = _Node()
        n.ulx, n.uly = int(hX * f32(i)), 0
        n.brx, n.bry = int(hX * f32(i + 1)), maxY - minY
        n.keys, n.nomore = [], False
        n.cid = counter[0]
        counter[0] += 1
        ini.append(n)
    for k in range(len(xs)):
        ini[min(nIni - 1, int(f32(xs[k]) / hX))].keys.append(k)
    nodes = []
    for n in ini:
        if len(n.keys) == 1:
            n.nomore = True
            nodes.append(n)
        elif len(n.keys) > 1:
            nodes.append(n)
    finish = False
    while not finish:
        prev = len(nodes)
        new_front = []  # children in push order; final list = reversed(new_front) + survivors
        survivors = []
        vsz = []
        nexp = 0
        for n in nodes:
            if n.nomore:
                survivors.append(n)
                continue
            for c in _divide(n, xs, ys, counter):
                if c.keys:
                    new_front.append(c)
                    if len(c.keys) > 1:
                        nexp += 1
                        vsz.append(c)
        nodes = new_front[::-1] + survivors
        if len(nodes) >= N or len(nodes) == prev:
            finish = True
        elif len(nodes) + nexp * 3 > N:
            while not finish:
                prev = len(nodes)
                vprev = sorted(vsz, key=lambda c: (len(c.keys), c.cid))
                vsz = []
                for n in reversed(vprev):
                    front = []
                    for c in _divide(n, xs, ys, counter):
                        if c.keys:
                            front.append(c)
                            if len(c.keys) > 1:
                                vsz.append(c)
                    nodes.remove(n)
                    nodes = front[::-1] + nodes
                    if len(nodes) >= N:
                        break
                if len(nodes) >= N or len(nodes) == prev:
                    finish = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if resp[k] > resp[best]:
                best = k
        out.append(best)
    return out


_fd_cache = {}


def _fast(cell, th):
    fd = _fd_cache.get(th)
    if fd is None:
        fd = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        _fd_cache[th] = fd
    return fd.detect(np.ascontiguousarray(cell))


def raw_keypoints_level(level_img, iniTh, minTh):
    """Per-cell FAST with retry, src/ORBextractor.cc:771-829. level_img is the ROI (no border)."""
    H, Wd = level_img.shape
    minBX = minBY = EDGE - 3
    maxBX, maxBY = Wd - EDGE + 3, H - EDGE + 3
    width, height = f32(maxBX - minBX), f32(maxBY - minBY)
    nCols, nRows = int(width / f32(30)), int(height / f32(30))
    xs, ys, rs = [], [], []
    if nCols <= 0 or nRows <= 0:
        return xs, ys, rs
    wCell = int(math.ceil(float(width / f32(nCols))))
    hCell = int(math.ceil(float(height / f32(nRows))))
    for i in range(nRows):
        iniY = minBY + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minBX + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            cell = level_img[iniY:maxY, iniX:maxX]
            k = _fast(cell, iniTh)
            if not k:
                k = _fast(cell, minTh)
            for p in k:
                xs.append(int(p.pt[0]) + j * wCell)
                ys.append(int(p.pt[1]) + i * hCell)
                rs.append(int(p.response))
    return xs, ys, rs


def ic_angle(img, x, y, umax):
    m01 = m10 = 0
    for u in range(-HALF, HALF + 1):
        m10 += u * int(img[y, x + u])
    for v in range(1, HALF + 1):
        d = umax[v]
        vs = 0
        for u in range(-d, d + 1):
            p, m = int(img[y + v, x + u]), int(img[y - v, x + u])
            vs += p - m
            m10 += u * (p + m)
        m01 += v * vs
    return f32(cv2.fastAtan2(float(m01), float(m10)))


def orb_descriptor(blur, x, y, angle_deg, pattern):
    factorPI = f32(math.pi / 180.0)
    ang = f32(angle_deg) * factorPI
    a, b = _cosf(ang), _sinf(ang)    # cos(float) under `using namespace std` = cosf (observed by compiling the reference)
    px = pattern[:, 0].astype(f32)
    py = pattern[:, 1].astype(f32)
    yy = np.rint((px * b).astype(f32) + (py * a).astype(f32)).astype(np.int64)
    xx = np.rint((px * a).astype(f32) - (py * b).astype(f32)).astype(np.int64)
    vals = blur[y + yy, x + xx].astype(np.int32)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
    return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)


def load_pattern():
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "spl_slam_b200", "csrc", "orb_pattern.inc")
    txt = open(p).read()
    txt = txt[txt.index("*/") + 2:]
    v = np.array([int(t) for t in txt.replace("\n", "").split(",") if t.strip()], np.int32)
    return v.reshape(512, 2)


def orb_extract(image, nfeatures, scaleFactor, nlevels, iniTh, minTh, stages=None):
    """ORBextractor::operator(), src/ORBextractor.cc:1043-1105."""
    scale, inv, per, umax = orb_tables(nfeatures, scaleFactor, nlevels)
    pat = load_pattern()
    pyr = compute_pyramid(image, inv)
    kps, descs = [], []
    for l in range(nlevels):
        roi = pyr[l][EDGE:-EDGE, EDGE:-EDGE]
        xs, ys, rs = raw_keypoints_level(roi, iniTh, minTh)
        if stages is not None:
            stages.setdefault("raw", []).append((xs, ys, rs))
            stages.setdefault("level", []).append(roi.copy())
        if not xs:
            continue
        keep = distribute_octree(xs, ys, rs, EDGE - 3, roi.shape[1] - EDGE + 3, EDGE - 3, roi.shape[0] - EDGE + 3, per[l])
        blur = cv2.GaussianBlur(roi.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        sz = f32(int(f32(31) * scale[l]))
        for k in keep:
            x, y = xs[k] + EDGE - 3, ys[k] + EDGE - 3
            ang = ic_angle(roi, x, y, umax)
            descs.append(orb_descriptor(blur, x, y, ang, pat))
            fx, fy = f32(x), f32(y)
            if l:
                fx, fy = f32(fx * scale[l]), f32(fy * scale[l])
            kps.append((fx, fy, sz, ang, f32(rs[k]), l, -1))
    from .oracle import KEYPOINT_DTYPE
    ka = np.array(kps, dtype=KEYPOINT_DTYPE) if kps else np.zeros(0, KEYPOINT_DTYPE)
    da = np.array(descs, np.uint8).reshape(-1, 32)
    return ka, da


def lsd_keylines(image, nlevels, opts, min_length):
    """LSDDetectorC::detect(image, kl, 2, nlevels, opts): pyrDown pyramid + cv2 LSD + KeyLine assembly."""
    from .oracle import KEYLINE_DTYPE
    out = []
    cur = image.copy()
    cid = -1
    for o in range(nlevels):
        if o:
            cur = cv2.pyrDown(cur, dstsize=(cur.shape[1] // 2, cur.shape[0] // 2))
        lsd = cv2.createLineSegmentDetector(*opts)
        L = lsd.detect(cur)[0]
        L = np.zeros((0, 4), f32) if L is None else L.reshape(-1, 4)
        h, w = cur.shape
        osc = f32(2.0 ** o)
        for e in L:
            e = e.copy()
            for i, lim in ((0, w), (2, w), (1, h), (3, h)):
                if e[i] < 0:
                    e[i] = 0
                if e[i] >= lim:
                    e[i] = f32(lim) - f32(1)
            length = float(f32(math.sqrt(float(e[0] - e[2]) ** 2 + float(e[1] - e[3]) ** 2)))
            if not length > min_length:
                continue
            cid += 1
            sx, sy, ex, ey = e[0] * osc, e[1] * osc, e[2] * osc, e[3] * osc
            ax, ay, bx, by = (int(np.rint(v)) for v in e)
            out.append((_atan2f(ey - sy, ex - sx), cid, o, (ex + sx) / f32(2), (ey + sy) / f32(2),
                        f32(length) / f32(max(w, h)), (ex - sx) * (ey - sy), sx, sy, ex, ey,
                        e[0], e[1], e[2], e[3], f32(length), max(abs(bx - ax), abs(by - ay)) + 1))
    return np.array(out, dtype=KEYLINE_DTYPE) if out else np.zeros(0, KEYLINE_DTYPE)


def knn2(q, t):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) as (idx, dist) nq x 2 arrays (-1 = missing)."""
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    m = bf.knnMatch(np.ascontiguousarray(q), np.ascontiguousarray(t), 2)
    idx = -np.ones((len(q), 2), np.int32)
    dist = -np.ones((len(q), 2), np.int32)
    for i, r in enumerate(m):
        for j, d in enumerate(r[:2]):
            idx[i, j] = d.trainIdx
            dist[i, j] = int(d.distance)
    return idx, dist
