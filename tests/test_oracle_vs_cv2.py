"""CPU: pin the oracle (oracle/*.c) bit-for-bit against cv2 4.13 primitives and against the independent
python/cv2 restatement of the reference glue (oracle/cv2_pipeline.py)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
cv2.setNumThreads(1)


def rnd(rng, h, w):
    return rng.integers(0, 256, (h, w), dtype=np.uint8)


def test_resize_linear(oracle):
    rng = np.random.default_rng(0)
    for (w, h) in [(640, 480), (752, 480), (1241, 376), (101, 77)]:
        img = rnd(rng, h, w)
        for s in [1 / 1.2, 0.77]:
            dw, dh = int(round(w * s)), int(round(h * s))
            assert np.array_equal(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR), oracle.resize_linear(img, dw, dh))


def test_border_blur_pyrdown_sobel(oracle):
    rng = np.random.default_rng(1)
    img = rnd(rng, 97, 133)
    assert np.array_equal(cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101), oracle.border_reflect101(img, 19))
    for (k, s) in [(7, 2.0), (5, 1.0), (7, 0.6), (7, 0.8), (13, 2.0), (3, 0.8)]:
        assert np.array_equal(cv2.GaussianBlur(img, (k, k), s, borderType=cv2.BORDER_REFLECT_101), oracle.gauss_blur(img, k, s))
    assert list(oracle.gauss_kernel_q8(7, 2.0)) == [18, 34, 48, 56, 48, 34, 18]
    assert list(oracle.gauss_kernel_q8(5, 1.0)) == [14, 62, 104, 62, 14]
    for (w, h) in [(640, 480), (641, 481), (97, 33)]:
        im = rnd(rng, h, w)
        assert np.array_equal(cv2.pyrDown(im, dstsize=(w // 2, h // 2)), oracle.pyrdown(im))
        dx, dy = oracle.sobel3(im)
        assert np.array_equal(cv2.Sobel(im, cv2.CV_16S, 1, 0, ksize=3), dx)
        assert np.array_equal(cv2.Sobel(im, cv2.CV_16S, 0, 1, ksize=3), dy)


def test_resize_linear_exact(oracle):
    rng = np.random.default_rng(2)
    for (w, h) in [(640, 480), (321, 243)]:
        img = rnd(rng, h, w)
        for s in [1.1, 1.05, 0.8]:
            assert np.array_equal(cv2.resize(img, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR_EXACT), oracle.resize_linear_exact(img, s))


def test_fast_atan2(oracle):
    rng = np.random.default_rng(3)
    ys = (rng.normal(size=20000) * rng.choice([1, 100, 1e4], 20000)).astype(np.float32)
    xs = (rng.normal(size=20000) * rng.choice([1, 100, 1e4], 20000)).astype(np.float32)
    for y, x in list(zip(ys, xs)) + [(0, 0), (0, 1), (1, 0), (0, -1), (-1, 0), (1, 1), (-1, -1)]:
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(oracle.fast_atan2(y, x))


def test_fast9_cells(oracle):
    rng = np.random.default_rng(4)
    base = oracle.synth_image(640, 480, 1)
    fds = {th: cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16) for th in (20, 7)}
    total = 0
    for it in range(300):
        w, h = int(rng.integers(7, 42)), int(rng.integers(7, 42))
        if it % 3 == 0:
            cell = rnd(rng, h, w)
        else:
            x0, y0 = int(rng.integers(0, 640 - w)), int(rng.integers(0, 480 - h))
            cell = np.ascontiguousarray(base[y0:y0 + h, x0:x0 + w])
        for th, fd in fds.items():
            a = [(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in fd.detect(cell)]
            xs, ys, sc = oracle.fast9(cell, th)
            assert a == list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
            total += len(a)
    assert total > 1000


@pytest.mark.parametrize("S", [1.1, 1.0, 0.8])
def test_lsd_identical_to_cv2(oracle, S):
    for (w, h, seed) in [(320, 240, 2), (376, 240, 4), (640, 480, 0)]:
        img = oracle.synth_image(w, h, seed)
        a = cv2.createLineSegmentDetector(0, S, 0.6, 2.2, 12.5, 1.0, 0.6, 1024).detect(img)[0]
        a = np.zeros((0, 4), np.float32) if a is None else a.reshape(-1, 4)
        b = oracle.lsd_detect(img, S, 0.6, 2.2, 12.5, 1024)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert len(a) > 20


def test_orb_pipeline_vs_python_restatement(oracle):
    from oracle import cv2_pipeline as P
    for (w, h, seed, nf) in [(640, 480, 0, 1000), (376, 240, 3, 500)]:
        img = oracle.synth_image(w, h, seed)
        k, d = oracle.ORBextractor(nf, 1.2, 8 if w > 400 else 5, 20, 7)(img)
        k2, d2 = P.orb_extract(img, nf, 1.2, 8 if w > 400 else 5, 20, 7)
        assert len(k) == len(k2) and np.array_equal(k.view(np.uint8), k2.view(np.uint8))
        assert np.array_equal(d, d2)


def test_keylines_vs_python_restatement(oracle):
    from oracle import cv2_pipeline as P
    img = oracle.synth_image(640, 480, 5)
    kl = oracle.lsd_detect_keylines(oracle.line_params(600), img)
    kl2 = P.lsd_keylines(img, 2, (0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024), 0.0)
    assert len(kl) == len(kl2) > 100 and np.array_equal(kl.view(np.uint8), kl2.view(np.uint8))


def test_knn2_vs_bfmatcher(oracle):
    from oracle import cv2_pipeline as P
    rng = np.random.default_rng(5)
    for hi in (256, 4):   # uniform and tie-heavy
        q = rng.integers(0, hi, (300, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (500, 32), dtype=np.uint8)
        for a, b in zip(oracle.knn2(q, t), P.knn2(q, t)):
            assert np.array_equal(a, b)
    for a, b in zip(oracle.knn2(q[:3], t[:1]), P.knn2(q[:3], t[:1])):
        assert np.array_equal(a, b)


CAMERAS = [   # (fx, fy, cx, cy), (k1, k2, p1, p2[, k3]): EuRoC cam0, TUM1, a mild 4-coefficient set
    ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
    ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
    ((718.856, 718.856, 607.1928, 185.2157), (0.1, -0.05, 0.001, -0.002)),
]


def test_undistort_points_vs_cv2(oracle):
    """Frame::UndistortKeyPoints / UndistortKeyLines call cv::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK)
    (src/Frame.cc:750, :785, :814-815): the restatement is bit-identical to cv2 4.13 on float points."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (fx, fy, cx, cy), dist in CAMERAS:
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        D = np.array(dist, np.float32)
        pts = np.stack([rng.uniform(-20, 1260, 20000), rng.uniform(-20, 500, 20000)], 1).astype(np.float32)
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
        got = oracle.undistort_points(pts, K[0, 0], K[1, 1], K[0, 2], K[1, 2], D)
        assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
