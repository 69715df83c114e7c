"""CPU: the product kernels, compiled for the CPU with the TEST-ONLY CUDA emulation (tests/emu), against the
oracle on small inputs.  This exercises the same kernel and host-driver source the GPU runs (logic, not speed)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def S():
    import spl_slam_b200 as S
    return S


def test_emu_matching(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    m = S.Linematcher(0.75, ctx=ctx)
    rng = np.random.default_rng(1)
    for (nq, nt, hi) in [(300, 500, 256), (260, 1025, 3), (5, 1, 256), (3, 0, 256)]:
        q = rng.integers(0, hi, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (nt, 32), dtype=np.uint8)
        i, d = m.knnMatch2(q, t)
        oi, od = oracle.knn2(q, t)
        assert np.array_equal(i, oi) and np.array_equal(d, od)
        mm, n = m.matchNNR(q, t)
        om, on = oracle.match_nnr(q, t, 0.75)
        assert np.array_equal(mm, om) and n == on
        mu, nu = m.matchNNRMutual(q, t)
        om21, _ = oracle.match_nnr(t, q, 0.75) if nt else (np.zeros(0, np.int32), 0)
        ref = om.copy()
        for a in range(nq):
            if ref[a] >= 0 and om21[ref[a]] != a:
                ref[a] = -1
        assert np.array_equal(mu, ref) and nu == int((ref >= 0).sum())
    assert m.DescriptorDistance(q[0], q[1]) == oracle.descriptor_distance(q[0], q[1])


@pytest.mark.parametrize("w,h,nf,nl,seed", [(160, 120, 100, 3, 0), (320, 240, 300, 5, 1)])
def test_emu_orb(S, oracle, emu_lib, w, h, nf, nl, seed):
    ctx = S.Context(0, emu_lib)
    ex = S.ORBextractor(nf, 1.2, nl, 20, 7, ctx=ctx)
    ox = oracle.ORBextractor(nf, 1.2, nl, 20, 7)
    assert list(ex.features_per_level()) == ox.features_per_level()
    assert np.array_equal(ex.GetScaleFactors(), np.array(ox.scale_factors(), np.float32))
    img = oracle.synth_image(w, h, seed)
    k, d = ex(img)
    ok, od = ox(img)
    for l in range(nl):
        assert np.array_equal(ex.pyramid_level(l), ox.level_image(l))
        if ox.level_blurred(l) is not None:
            assert np.array_equal(ex.debug_blurred(l), ox.level_blurred(l))
        xs, ys, rr = ex.debug_raw_keys(l)
        oxs, oys, orr = ox.level_raw(l)
        assert sorted(zip(xs.tolist(), ys.tolist(), rr.tolist())) == sorted(zip(oxs.tolist(), oys.tolist(), orr.tolist()))
    assert len(k) == len(ok) and np.array_equal(k.view(np.uint8), ok.view(np.uint8))
    assert np.array_equal(d, od)


def test_emu_orb_batch_and_empty(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    ex = S.ORBextractor(100, 1.2, 3, 20, 7, ctx=ctx)
    ox = oracle.ORBextractor(100, 1.2, 3, 20, 7)
    imgs = np.stack([oracle.synth_image(160, 120, s) for s in (3, 4)])
    ks, ds = ex.extract_batch(imgs)
    for b in range(2):
        ok, od = ox(imgs[b])
        assert np.array_equal(ks[b].view(np.uint8), ok.view(np.uint8)) and np.array_equal(ds[b], od)
    k, d = ex(np.zeros((0, 0), np.uint8))       # empty image: silent return
    assert len(k) == 0 and d.shape == (0, 32)
    flat = np.full((120, 160), 128, np.uint8)   # no corners at all
    k, d = ex(flat)
    assert len(k) == 0
    with pytest.raises(AssertionError):
        ex(np.zeros((10, 10, 3), np.uint8))      # the reference asserts CV_8UC1


def test_emu_octree_standalone(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    rng = np.random.default_rng(9)
    for (W, H, n, N) in [(608, 448, 900, 217), (1209, 344, 1500, 434), (300, 200, 50, 60), (300, 200, 5, 0)]:
        pts = set()
        while len(pts) < n:
            pts.add((int(rng.integers(3, W - 3)), int(rng.integers(3, H - 3))))
        pts = sorted(pts, key=lambda p: (p[1], p[0]))
        xs = np.array([p[0] for p in pts], np.int32)
        ys = np.array([p[1] for p in pts], np.int32)
        rr = rng.integers(7, 40, n).astype(np.int32)   # many response ties
        # the oracle takes keys in the reference's push_back order (cell-major); reorder accordingly
        nC, nR = int(np.float32(W) / np.float32(30)), int(np.float32(H) / np.float32(30))
        wC, hC = int(np.ceil(np.float32(W) / nC)), int(np.ceil(np.float32(H) / nR))
        order = sorted(range(n), key=lambda i: ((ys[i] - 3) // hC, (xs[i] - 3) // wC, ys[i], xs[i]))
        xo, yo, ro = xs[order], ys[order], rr[order]
        ref = oracle.distribute_octree(xo, yo, ro, 16, 16 + W, 16, 16 + H, N)
        got = S.distribute_octree(ctx, xo, yo, ro, 16, 16 + W, 16, 16 + H, N)
        assert np.array_equal(ref, got), (W, H, n, N)


@pytest.mark.parametrize("w,h,nf,sc,seed", [(128, 96, 40, 1.1, 0), (192, 128, 20, 1.1, 3)])
def test_emu_lines(S, oracle, emu_lib, w, h, nf, sc, seed):
    ctx = S.Context(0, emu_lib)
    le = S.Lineextractor(nf, 2, 0, sc, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=ctx)
    prm = oracle.line_params(nf, 2, 0, sc, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
    assert list(le.features_per_level()) == oracle.features_per_level_lines(prm)
    img = oracle.synth_image(w, h, seed)
    kl = le.lsd_detect(img)
    okl = oracle.lsd_detect_keylines(prm, img)
    assert len(kl) == len(okl) > 5 and np.array_equal(kl.view(np.uint8), okl.view(np.uint8))
    d, fd = le.lbd_compute(img, okl, want_float=True)
    od, ofd = oracle.lbd_compute(img, okl, want_float=True)
    assert np.array_equal(d, od) and np.array_equal(fd.view(np.uint32), ofd.view(np.uint32))
    K, M, D = le.ComputeLsdWithLbd(img)
    oK, oM, oD = oracle.line_extract(prm, img)
    assert len(K) == len(oK) and np.array_equal(K.view(np.uint8), oK.view(np.uint8))
    assert np.array_equal(M.view(np.uint8), oM.view(np.uint8)) and np.array_equal(D, oD)


def test_emu_lines_empty_and_flat(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    le = S.Lineextractor(40, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=ctx)
    K, M, D = le.ComputeLsdWithLbd(np.zeros((0, 0), np.uint8))
    assert len(K) == 0 and D.shape == (0, 32)
    K, M, D = le.ComputeLsdWithLbd(np.full((120, 160), 90, np.uint8))   # no gradient anywhere
    assert len(K) == 0


def test_emu_lines_batch_with_flat_frame(S, oracle, emu_lib):
    """Regions live in per-frame slot blocks: a batch whose middle frame has no region at all."""
    ctx = S.Context(0, emu_lib)
    le = S.Lineextractor(40, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=ctx)
    prm = oracle.line_params(40, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
    imgs = np.stack([oracle.synth_image(128, 96, 5), np.full((96, 128), 90, np.uint8), oracle.synth_image(128, 96, 6)])
    Ks, Ms, Ds = le.extract_batch(imgs)
    for b in range(3):
        oK, oM, oD = oracle.line_extract(prm, imgs[b])
        assert len(Ks[b]) == len(oK) and np.array_equal(Ks[b].view(np.uint8), oK.view(np.uint8)) and np.array_equal(Ds[b], oD)
    assert len(Ks[1]) == 0 and len(Ks[0]) > 5 and len(Ks[2]) > 5


def _stereo_pair(oracle, w, h, seed, shift):
    left = oracle.synth_image(w, h, seed)
    rng = np.random.default_rng(seed + 1000)
    right = np.roll(left, -shift, axis=1).astype(np.int16)
    right[h // 2:] = np.roll(left, -(shift + 6), axis=1)[h // 2:]      # two disparity bands
    right = np.clip(right + rng.integers(-2, 3, right.shape), 0, 255).astype(np.uint8)
    return left, right


def test_emu_stereo_match(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    left, right = _stereo_pair(oracle, 256, 192, 3, 9)
    exL = S.ORBextractor(300, 1.2, 4, 20, 7, ctx=ctx); exR = S.ORBextractor(300, 1.2, 4, 20, 7, ctx=ctx)
    oxL = oracle.ORBextractor(300, 1.2, 4, 20, 7); oxR = oracle.ORBextractor(300, 1.2, 4, 20, 7)
    kL, dL = exL(left); kR, dR = exR(right)
    okL, odL = oxL(left); okR, odR = oxR(right)
    assert np.array_equal(kL.view(np.uint8), okL.view(np.uint8)) and np.array_equal(kR.view(np.uint8), okR.view(np.uint8))
    u, z = exL.ComputeStereoMatches(exR, kL, dL, kR, dR, 0.11, 0.11 * 200.0)
    ou, oz = oracle.stereo_match(oxL, oxR, okL, odL, okR, odR, 0.11, 0.11 * 200.0)
    assert (ou >= 0).sum() > 20                      # the case is not vacuous
    assert np.array_equal(u.view(np.uint32), ou.view(np.uint32)) and np.array_equal(z.view(np.uint32), oz.view(np.uint32))
    # one extractor, the pair as two frames of one batch
    ks, ds = exL.extract_batch(np.stack([left, right]))
    u2, z2 = exL.ComputeStereoMatches(exL, ks[0], ds[0], ks[1], ds[1], 0.11, 0.11 * 200.0, frame_l=0, frame_r=1)
    assert np.array_equal(u2.view(np.uint32), ou.view(np.uint32)) and np.array_equal(z2.view(np.uint32), oz.view(np.uint32))
    # no right keypoints -> nothing matched
    u3, z3 = exL.ComputeStereoMatches(exL, ks[0], ds[0], ks[1][:0], ds[1][:0], 0.11, 22.0, frame_l=0, frame_r=1)
    assert (u3 == -1).all() and (z3 == -1).all()


def test_emu_candidates(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    m = S.ORBmatcher(0.9, ctx=ctx)
    rng = np.random.default_rng(5)
    for hi in (256, 3):                                # hi = 3: tie-heavy
        q = rng.integers(0, hi, (70, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (150, 32), dtype=np.uint8)
        lists = [rng.integers(0, 150, int(n)).astype(np.int32) for n in rng.integers(0, 90, 70)]
        lists[0] = np.zeros(0, np.int32); lists[1] = np.array([7], np.int32)
        bi, bd, cd = m.candidates_top2(q, t, lists, want_dist=True)
        obi, obd, ocd = oracle.candidates_top2(q, t, lists)
        assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(cd, ocd)


def _grid_case(oracle, n, nq, seed, w=752, h=480):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, oracle.KEYPOINT_DTYPE)
    k["x"] = rng.uniform(-6, w + 6, n).astype(np.float32); k["y"] = rng.uniform(-6, h + 6, n).astype(np.float32)   # some fall outside the grid
    k["octave"] = rng.integers(0, 8, n)
    qx = rng.uniform(-20, w + 20, nq).astype(np.float32); qy = rng.uniform(-20, h + 20, nq).astype(np.float32)
    qr = rng.uniform(1, 60, nq).astype(np.float32)
    qx[:3] = k["x"][:3]; qy[:3] = k["y"][:3]                       # exact hits: |d| < r boundary cases
    mn = rng.integers(-1, 4, nq).astype(np.int32); mx = rng.integers(-1, 8, nq).astype(np.int32)
    return k, qx, qy, qr, mn, mx


def test_emu_grid_area_queries(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    k, qx, qy, qr, mn, mx = _grid_case(oracle, 700, 300, 11)
    og = oracle.grid_params(64, 48, 0, 752, 0, 480)
    g = S.GridParams.for_image(64, 48, 0, 752, 0, 480)
    for lv in (False, True):
        off, idx = S.features_in_area(ctx, k, g, qx, qy, qr, mn if lv else None, mx if lv else None)
        ooff, oidx = oracle.grid_candidates(k, og, qx, qy, qr, mn if lv else None, mx if lv else None)
        assert np.array_equal(off, ooff) and np.array_equal(idx, oidx) and len(idx) > 1000
    # line grid: mid-points + end points must be inside (PosInGridLines)
    kl = np.zeros(len(k), oracle.KEYLINE_DTYPE)
    rng = np.random.default_rng(3)
    kl["startPointX"] = k["x"] + rng.uniform(-30, 30, len(k)).astype(np.float32); kl["startPointY"] = k["y"] + rng.uniform(-30, 30, len(k)).astype(np.float32)
    kl["endPointX"] = 2 * k["x"] - kl["startPointX"]; kl["endPointY"] = 2 * k["y"] - kl["startPointY"]
    ogl = oracle.grid_params(16, 12, 0, 752, 0, 480); gl = S.GridParams.for_image(16, 12, 0, 752, 0, 480)
    off, idx = S.features_in_area(ctx, k, gl, qx, qy, qr, keylines=kl)
    ooff, oidx = oracle.grid_candidates(k, ogl, qx, qy, qr, keylines=kl)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx)
    # no features / no queries
    off, idx = S.features_in_area(ctx, k[:0], g, qx, qy, qr)
    assert (off == 0).all() and len(idx) == 0


def _write_vocab_text(path, k, L, vocab):
    parent, desc, weight, leaf = vocab
    with open(path, "w") as f:
        f.write("%d %d  0 0\n" % (k, L))                      # "k L scoring weighting" (saveToTextFile writes two spaces)
        for i in range(1, len(parent)):
            f.write("%d %d %s %r\n" % (parent[i], leaf[i], " ".join(str(int(b)) for b in desc[i]), float(weight[i])))


def test_emu_bow_transform(S, oracle, emu_lib, tmp_path):
    ctx = S.Context(0, emu_lib)
    rng = np.random.default_rng(8)
    for (k, L, levelsup) in [(10, 3, 1), (4, 5, 4), (3, 4, 9)]:
        vocab = oracle.synth_vocabulary(k, L, seed=k * 10 + L)
        voc = S.ORBVocabulary(ctx, k, L, *vocab)
        assert voc.nwords == k ** L and voc.nnodes == len(vocab[0])
        feats = rng.integers(0, 256, (300, 32), dtype=np.uint8)
        feats[:40] = vocab[1][rng.integers(1, len(vocab[0]), 40)]          # exact node descriptors: distance 0, sibling ties
        w, wt, nd = voc.transform_features(feats, levelsup)
        ow, owt, ond = oracle.bow_transform(vocab, L, feats, levelsup)
        assert np.array_equal(w, ow) and np.array_equal(wt.view(np.uint64), owt.view(np.uint64)) and np.array_equal(nd, ond)
        v, fv = voc.transform(feats, levelsup)
        ov, ofv = oracle.bow_vectors(ow, owt, ond)
        assert v == ov and fv == ofv and abs(sum(v.values()) - 1.0) < 1e-12
        # the same tree through the ORBvoc.txt text format
        path = str(tmp_path / ("voc_%d_%d.txt" % (k, L)))
        _write_vocab_text(path, k, L, vocab)
        voc2 = S.ORBVocabulary(ctx, path=path)
        w2, wt2, nd2 = voc2.transform_features(feats, levelsup)
        assert np.array_equal(w2, ow) and np.array_equal(wt2.view(np.uint64), owt.view(np.uint64)) and np.array_equal(nd2, ond)
    assert len(voc.transform_features(feats[:0])[0]) == 0


def test_emu_undistort(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    rng = np.random.default_rng(4)
    n = 500
    k = np.zeros(n, oracle.KEYPOINT_DTYPE)
    k["x"] = rng.uniform(0, 752, n).astype(np.float32); k["y"] = rng.uniform(0, 480, n).astype(np.float32)
    k["octave"] = rng.integers(0, 8, n); k["response"] = rng.uniform(7, 90, n).astype(np.float32); k["class_id"] = -1
    kl = np.zeros(n, oracle.KEYLINE_DTYPE)
    for f in ("startPointX", "endPointX"): kl[f] = rng.uniform(0, 752, n).astype(np.float32)
    for f in ("startPointY", "endPointY"): kl[f] = rng.uniform(0, 480, n).astype(np.float32)
    kl["lineLength"] = 33.0; kl["octave"] = 1
    for (fx, fy, cx, cy), dist in [((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
                                   ((517.3, 516.5, 318.6, 255.3), (0.2624, -0.9531, -0.0054, 0.0026, 1.1633)),
                                   ((435.2, 435.2, 367.4, 252.2), (0.0, 0.0, 0.0, 0.0))]:      # k1 == 0: plain copy
        cam = S.Camera.make(fx, fy, cx, cy, dist)
        out = S.undistort_keypoints(ctx, cam, k)
        ref = oracle.undistort_keypoints(k, np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy), dist)
        assert np.array_equal(out.view(np.uint8), ref.view(np.uint8))
        okl, om = S.undistort_keylines(ctx, cam, kl, k)
        rkl, rm = oracle.undistort_keylines(kl, k, np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy), dist)
        assert np.array_equal(okl.view(np.uint8), rkl.view(np.uint8)) and np.array_equal(om.view(np.uint8), rm.view(np.uint8))
        if dist[0] != 0:
            assert not np.array_equal(out["x"], k["x"])
