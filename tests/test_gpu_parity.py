"""GPU (B200): parity of the CUDA path, called through the C ABI, against the oracle.
Bit-exact for keypoints, octaves, descriptors and match indices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import spl_slam_b200 as S
    return S


def _check_orb(ex, ox, img):
    k, d = ex(img)
    ok, od = ox(img)
    assert len(k) == len(ok), (len(k), len(ok))
    for f in k.dtype.names:
        assert np.array_equal(k[f], ok[f]), f     # bit-exact incl. the float angle (same float ops, no FMA)
    assert np.array_equal(d, od)
    return len(k)


@pytest.mark.parametrize("w,h,nf,seed", [(640, 480, 1000, 0), (752, 480, 1200, 1), (1241, 376, 2000, 2), (1920, 1080, 2000, 3)])
def test_orb_configs(S, oracle, gpu_ctx, w, h, nf, seed):
    ex = S.ORBextractor(nf, 1.2, 8, 20, 7, ctx=gpu_ctx)
    ox = oracle.ORBextractor(nf, 1.2, 8, 20, 7)
    img = oracle.synth_image(w, h, seed)
    n = _check_orb(ex, ox, img)
    assert n >= nf * 0.9
    for l in range(8):
        assert np.array_equal(ex.pyramid_level(l), ox.level_image(l))
        if ox.level_blurred(l) is not None:
            assert np.array_equal(ex.debug_blurred(l), ox.level_blurred(l))
        xs, ys, rr = ex.debug_raw_keys(l)
        oxs, oys, orr = ox.level_raw(l)
        # the reference appends cell by cell (rows, then columns) and row-major inside a cell (ORBextractor.cc:783-830); the kernel
        # appends in arrival order and carries that order as an explicit key, so: put the kernel's list into the reference's order
        # and compare the two SEQUENCES (not sets)
        lh, lw = ox.level_image(l).shape
        width, height = lw - 32, lh - 32
        ncols, nrows = width // 30, height // 30
        wcell, hcell = -(-width // ncols), -(-height // nrows)
        key = lambda x, y: ((y - 3) // hcell, (x - 3) // wcell, y, x)
        got = sorted(zip(xs.tolist(), ys.tolist(), rr.tolist()), key=lambda t: key(t[0], t[1]))
        assert got == list(zip(oxs.tolist(), oys.tolist(), orr.tolist())), "raw FAST list of level %d" % l


def test_orb_batch_equals_single_and_stereo_pair(S, oracle, gpu_ctx):
    ex = S.ORBextractor(1200, 1.2, 8, 20, 7, ctx=gpu_ctx)
    ox = oracle.ORBextractor(1200, 1.2, 8, 20, 7)
    left = oracle.synth_image(752, 480, 10)
    right = np.roll(left, -7, axis=1).copy()          # EuRoC-style pair: right = left shifted by a disparity
    imgs = np.stack([left, right] + [oracle.synth_image(752, 480, 20 + i) for i in range(6)])
    ks, ds = ex.extract_batch(imgs)
    for b in range(len(imgs)):
        ok, od = ox(imgs[b])
        assert np.array_equal(ks[b].view(np.uint8), ok.view(np.uint8)) and np.array_equal(ds[b], od)


def test_orb_edge_cases(S, oracle, gpu_ctx):
    ex = S.ORBextractor(500, 1.2, 8, 20, 7, ctx=gpu_ctx)
    ox = oracle.ORBextractor(500, 1.2, 8, 20, 7)
    k, d = ex(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.full((480, 640), 77, np.uint8))
    assert len(k) == 0
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)    # dense corners: stresses the raw key lists
    _check_orb(ex, ox, noise)
    strided = np.zeros((480, 700), np.uint8)
    strided[:, :640] = oracle.synth_image(640, 480, 5)
    _check_orb(ex, ox, strided[:, :640])                          # non-contiguous rows (stride != width)
    with pytest.raises(S.PlfError):
        ex(np.zeros((40, 40), np.uint8))                          # too small for 8 levels: explicit error


def test_octree_standalone_ties(S, oracle, gpu_ctx):
    rng = np.random.default_rng(9)
    for (W, H, n, N) in [(608, 448, 3000, 217), (1209, 344, 6000, 434), (1888, 1048, 20000, 434), (300, 200, 5, 0)]:
        pts = set()
        while len(pts) < n:
            pts.add((int(rng.integers(3, W - 3)), int(rng.integers(3, H - 3))))
        pts = list(pts)
        xs = np.array([p[0] for p in pts], np.int32)
        ys = np.array([p[1] for p in pts], np.int32)
        rr = rng.integers(7, 30, n).astype(np.int32)
        nC, nR = int(np.float32(W) / np.float32(30)), int(np.float32(H) / np.float32(30))
        wC, hC = int(np.ceil(np.float32(W) / nC)), int(np.ceil(np.float32(H) / nR))
        order = sorted(range(n), key=lambda i: ((ys[i] - 3) // hC, (xs[i] - 3) // wC, ys[i], xs[i]))
        xo, yo, ro = xs[order], ys[order], rr[order]
        ref = oracle.distribute_octree(xo, yo, ro, 16, 16 + W, 16, 16 + H, N)
        got = S.distribute_octree(gpu_ctx, xo, yo, ro, 16, 16 + W, 16, 16 + H, N)
        assert np.array_equal(ref, got), (W, H, n, N)


@pytest.mark.parametrize("nq,nt,hi", [(600, 600, 256), (2000, 2000, 256), (1000, 5000, 4), (10000, 10000, 256), (7, 1, 256), (7, 0, 256)])
def test_knn2_and_nnr(S, oracle, gpu_ctx, nq, nt, hi):
    m = S.Linematcher(0.75, ctx=gpu_ctx)
    rng = np.random.default_rng(1234)
    q = rng.integers(0, hi, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, hi, (nt, 32), dtype=np.uint8)
    i, d = m.knnMatch2(q, t)
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(i, oi) and np.array_equal(d, od)
    mm, n = m.matchNNR(q, t)
    om, on = oracle.match_nnr(q, t, 0.75)
    assert np.array_equal(mm, om) and n == on
    if nt and nq <= 2000:
        mu, nu = m.matchNNRMutual(q, t)
        om21, _ = oracle.match_nnr(t, q, 0.75)
        ref = om.copy()
        for a in range(nq):
            if ref[a] >= 0 and om21[ref[a]] != a:
                ref[a] = -1
        assert np.array_equal(mu, ref) and nu == int((ref >= 0).sum())


@pytest.mark.parametrize("nt,nsample,hi", [(1000000, 16, 256), (10000000, 4, 256), (1000000, 8, 2)])
def test_knn2_large_properties(S, oracle, gpu_ctx, nt, nsample, hi):
    """1e4 x 1e6 / 1e7 (BASELINE config 5, incl. a tie-heavy variant): too slow for the scalar oracle in full, so check
    a query sample against the oracle and the whole result through the shard/merge property (sharded == unsharded)."""
    import torch
    rng = np.random.default_rng(1234)
    nq = 10000
    q = rng.integers(0, hi, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, hi, (nt, 32), dtype=np.uint8)
    lib, h = gpu_ctx.lib, gpu_ctx.h
    dq = torch.from_numpy(q).cuda(); dt = torch.from_numpy(t).cuda()
    idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda"); dist = torch.empty_like(idx)
    torch.cuda.synchronize()
    gpu_ctx.check(lib.plf_hamming_knn2_device(h, dq.data_ptr(), nq, dt.data_ptr(), nt, 0, idx.data_ptr(), dist.data_ptr()))
    gpu_ctx.synchronize()
    sample = rng.choice(nq, nsample, replace=False)
    oi, od = oracle.knn2(q[sample], t)
    assert np.array_equal(idx.cpu().numpy()[sample], oi) and np.array_equal(dist.cpu().numpy()[sample], od)
    shards = 4
    pidx = torch.empty((shards, nq, 2), dtype=torch.int32, device="cuda"); pdist = torch.empty_like(pidx)
    per = nt // shards
    for s in range(shards):
        gpu_ctx.check(lib.plf_hamming_knn2_device(h, dq.data_ptr(), nq, dt[s * per:].data_ptr(), per, s * per,
                                                  pidx[s].data_ptr(), pdist[s].data_ptr()))
    midx = torch.empty_like(idx); mdist = torch.empty_like(dist)
    gpu_ctx.check(lib.plf_knn2_merge_device(h, pidx.data_ptr(), pdist.data_ptr(), shards, nq, midx.data_ptr(), mdist.data_ptr()))
    gpu_ctx.synchronize()
    assert torch.equal(midx, idx) and torch.equal(mdist, dist)


LSD_TUM = dict(refine=0, scale=1.1, sigma_scale=0.6, quant=2.2, ang_th=12.5, log_eps=1.0, density_th=0.6, n_bins=1024)


def _line_objs(S, oracle, ctx, nf, min_len=0.0, **kw):
    o = dict(LSD_TUM); o.update(kw)
    le = S.Lineextractor(nf, 2, o["refine"], o["scale"], o["sigma_scale"], o["quant"], o["ang_th"], o["log_eps"],
                         o["density_th"], o["n_bins"], min_len, ctx=ctx)
    prm = oracle.line_params(nf, 2, o["refine"], o["scale"], o["sigma_scale"], o["quant"], o["ang_th"], o["log_eps"],
                             o["density_th"], o["n_bins"], min_len)
    return le, prm


def _check_lines(le, oracle, prm, img):
    kl = le.lsd_detect(img)
    okl = oracle.lsd_detect_keylines(prm, img)
    assert len(kl) == len(okl), (len(kl), len(okl))
    for f in ("octave", "class_id", "numOfPixels"):
        assert np.array_equal(kl[f], okl[f]), f
    # line end points / angles: stated tolerance 1e-3 px / 1e-5 rad (double cos/sin differ by <= 1 ulp between
    # libm and the CUDA math library); in practice they come out identical
    for f in ("startPointX", "startPointY", "endPointX", "endPointY", "sPointInOctaveX", "sPointInOctaveY",
              "ePointInOctaveX", "ePointInOctaveY", "lineLength", "pt_x", "pt_y"):
        assert np.allclose(kl[f], okl[f], rtol=0, atol=1e-3), f
    assert np.allclose(kl["angle"], okl["angle"], rtol=0, atol=1e-5)
    K, M, D = le.ComputeLsdWithLbd(img)
    oK, oM, oD = oracle.line_extract(prm, img)
    assert len(K) == len(oK)
    assert np.array_equal(K["octave"], oK["octave"]) and np.array_equal(K["class_id"], oK["class_id"])
    assert np.allclose(K["startPointX"], oK["startPointX"], atol=1e-3) and np.allclose(M["x"], oM["x"], atol=1e-3)
    assert np.array_equal(D, oD)          # LBD binary descriptors bit-exact
    # LBD alone on the oracle's keylines: float descriptor bit-exact
    d, fd = le.lbd_compute(img, okl, want_float=True)
    od, ofd = oracle.lbd_compute(img, okl, want_float=True)
    assert np.array_equal(d, od)
    assert np.array_equal(fd.view(np.uint32), ofd.view(np.uint32))
    return len(okl), bool(np.array_equal(kl.view(np.uint8), okl.view(np.uint8)))


@pytest.mark.parametrize("w,h,nf,seed,kw", [
    (640, 480, 600, 0, {}),
    (752, 480, 200, 1, dict(sigma_scale=0.8, density_th=0.8)),      # EuRoC mono line settings
    (752, 480, 600, 2, {}),
    (1241, 376, 800, 3, {}),
    (640, 480, 100, 4, dict(scale=1.0)),                              # SCALE == 1: no pre-blur / resize
    (640, 480, 300, 5, dict(scale=0.8)),
])
def test_line_configs(S, oracle, gpu_ctx, w, h, nf, seed, kw):
    le, prm = _line_objs(S, oracle, gpu_ctx, nf, **kw)
    img = oracle.synth_image(w, h, seed)
    n, identical = _check_lines(le, oracle, prm, img)
    assert n > 100
    assert identical, "keylines are within tolerance but not bit-identical"


def test_line_min_length_batch_and_edges(S, oracle, gpu_ctx):
    le, prm = _line_objs(S, oracle, gpu_ctx, 200, min_len=0.02 * 752)
    imgs = np.stack([oracle.synth_image(752, 480, 30 + i) for i in range(4)])
    Ks, Ms, Ds = le.extract_batch(imgs)
    for b in range(len(imgs)):
        oK, oM, oD = oracle.line_extract(prm, imgs[b])
        assert len(Ks[b]) == len(oK) and np.array_equal(Ds[b], oD)
        assert np.array_equal(Ks[b].view(np.uint8), oK.view(np.uint8))
        assert np.array_equal(Ms[b].view(np.uint8), oM.view(np.uint8))
    K, M, D = le.ComputeLsdWithLbd(np.zeros((0, 0), np.uint8))
    assert len(K) == 0
    K, M, D = le.ComputeLsdWithLbd(np.full((480, 752), 200, np.uint8))
    assert len(K) == 0
    # a frame without any region between two ordinary ones: its block of region slots stays empty
    mixed = np.stack([imgs[0], np.full((480, 752), 200, np.uint8), imgs[1]])
    Ks2, Ms2, Ds2 = le.extract_batch(mixed)
    assert len(Ks2[1]) == 0 and np.array_equal(Ds2[0], Ds[0]) and np.array_equal(Ds2[2], Ds[1])
    assert np.array_equal(Ks2[0].view(np.uint8), Ks[0].view(np.uint8)) and np.array_equal(Ks2[2].view(np.uint8), Ks[1].view(np.uint8))
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (240, 320), dtype=np.uint8)   # every pixel has a defined gradient
    oK, oM, oD = oracle.line_extract(prm, noise)
    K, M, D = le.ComputeLsdWithLbd(noise)
    assert len(K) == len(oK) and np.array_equal(D, oD)


def test_line_1080p(S, oracle, gpu_ctx):
    le, prm = _line_objs(S, oracle, gpu_ctx, 800)
    img = oracle.synth_image(1920, 1080, 11)
    oK, oM, oD = oracle.line_extract(prm, img)
    K, M, D = le.ComputeLsdWithLbd(img)
    assert len(K) == len(oK) == 800 and np.array_equal(D, oD)
    assert np.array_equal(K.view(np.uint8), oK.view(np.uint8))


def _stereo_pair(oracle, w, h, seed, shift):
    left = oracle.synth_image(w, h, seed)
    rng = np.random.default_rng(seed + 1000)
    right = np.roll(left, -shift, axis=1).astype(np.int16)
    right[h // 2:] = np.roll(left, -(shift + 6), axis=1)[h // 2:]      # two disparity bands
    right = np.clip(right + rng.integers(-2, 3, right.shape), 0, 255).astype(np.uint8)
    return left, right


@pytest.mark.parametrize("w,h,nf,mb,fx,seed", [(752, 480, 1200, 0.11, 435.2, 40), (1241, 376, 2000, 0.537, 718.9, 41)])
def test_stereo_matches(S, oracle, gpu_ctx, w, h, nf, mb, fx, seed):
    """BASELINE configs 2 / 3: Frame::ComputeStereoMatches on an EuRoC- / KITTI-style pair, bit-exact mvuRight, mvDepth."""
    import torch
    left, right = _stereo_pair(oracle, w, h, seed, 11)
    ctx2 = S.Context(0)                                   # the right extractor on its own stream, as in Frame.cc:116-119
    exL = S.ORBextractor(nf, 1.2, 8, 20, 7, ctx=gpu_ctx); exR = S.ORBextractor(nf, 1.2, 8, 20, 7, ctx=ctx2)
    oxL = oracle.ORBextractor(nf, 1.2, 8, 20, 7); oxR = oracle.ORBextractor(nf, 1.2, 8, 20, 7)
    kL, dL = exL(left); kR, dR = exR(right)
    okL, odL = oxL(left); okR, odR = oxR(right)
    assert np.array_equal(kL.view(np.uint8), okL.view(np.uint8)) and np.array_equal(kR.view(np.uint8), okR.view(np.uint8))
    mbf = np.float32(mb * fx)
    u, z = exL.ComputeStereoMatches(exR, kL, dL, kR, dR, mb, mbf)
    ou, oz = oracle.stereo_match(oxL, oxR, okL, odL, okR, odR, mb, mbf)
    assert (ou >= 0).sum() > nf // 10
    assert np.array_equal(u.view(np.uint32), ou.view(np.uint32)) and np.array_equal(z.view(np.uint32), oz.view(np.uint32))
    # batched device-resident variant: 3 pairs interleaved (L, R, L, R, ...) in one extractor batch
    pairs = [(left, right), _stereo_pair(oracle, w, h, seed + 1, 5), _stereo_pair(oracle, w, h, seed + 2, 20)]
    imgs = torch.from_numpy(np.stack([im for p in pairs for im in p])).cuda()
    cap = exL.max_keypoints
    nfr = len(imgs)
    dk = torch.zeros((nfr, cap, 7), dtype=torch.float32, device="cuda"); dd = torch.zeros((nfr, cap, 32), dtype=torch.uint8, device="cuda")
    dn = torch.zeros(nfr, dtype=torch.int32, device="cuda")
    du = torch.zeros((3, cap), dtype=torch.float32, device="cuda"); dz = torch.zeros_like(du)
    torch.cuda.synchronize()
    lib = gpu_ctx.lib
    gpu_ctx.check(lib.plf_orb_extract_batch_device(exL.h, imgs.data_ptr(), nfr, w, h, w, w * h, dk.data_ptr(), dd.data_ptr(), cap, dn.data_ptr()))
    gpu_ctx.check(lib.plf_stereo_match_batch_device(exL.h, exL.h, 3, 0, 2, 1, 2, dk.data_ptr(), dd.data_ptr(), dn.data_ptr(),
                                                    dk.data_ptr(), dd.data_ptr(), dn.data_ptr(), cap, mb, mbf, du.data_ptr(), dz.data_ptr()))
    gpu_ctx.synchronize()
    for p, (l, r) in enumerate(pairs):
        okL, odL = oxL(l); okR, odR = oxR(r)
        ou, oz = oracle.stereo_match(oxL, oxR, okL, odL, okR, odR, mb, mbf)
        n = len(okL)
        assert int(dn[2 * p]) == n
        assert np.array_equal(du[p, :n].cpu().numpy().view(np.uint32), ou.view(np.uint32))
        assert np.array_equal(dz[p, :n].cpu().numpy().view(np.uint32), oz.view(np.uint32))


def test_candidate_lists(S, oracle, gpu_ctx):
    m = S.ORBmatcher(0.9, ctx=gpu_ctx)
    rng = np.random.default_rng(5)
    for hi, nq, nt, maxc in [(256, 2000, 2000, 60), (3, 500, 800, 300), (256, 3, 10, 5000)]:
        q = rng.integers(0, hi, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (nt, 32), dtype=np.uint8)
        lists = [rng.integers(0, nt, int(n)).astype(np.int32) for n in rng.integers(0, maxc, nq)]
        lists[0] = np.zeros(0, np.int32); lists[1] = np.array([7], np.int32)
        bi, bd, cd = m.candidates_top2(q, t, lists, want_dist=True)
        obi, obd, ocd = oracle.candidates_top2(q, t, lists)
        assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(cd, ocd)
    with pytest.raises(S.PlfError):
        m.candidates_top2(q, t, [np.array([nt], np.int32)] + [np.zeros(0, np.int32)] * (nq - 1))


def test_grid_area_queries_and_projection_matching(S, oracle, gpu_ctx):
    """The data path of the tracking matchers (ORBmatcher::SearchByProjection / SearchForInitialization): Frame grid ->
    GetFeaturesInArea candidate lists -> top-2 over each list, all bit-identical to the sequential reference code."""
    ox = oracle.ORBextractor(2000, 1.2, 8, 20, 7)
    ex = S.ORBextractor(2000, 1.2, 8, 20, 7, ctx=gpu_ctx)
    img1 = oracle.synth_image(1241, 376, 50)
    img2 = np.roll(img1, (2, -5), axis=(0, 1)).copy()
    k1, d1 = ex(img1); k2, d2 = ex(img2)
    g = S.GridParams.for_image(64, 48, 0, 1241, 0, 376); og = oracle.grid_params(64, 48, 0, 1241, 0, 376)
    # SearchForInitialization (src/ORBmatcher.cc:406-456): level-0 keypoints of frame 1 search a 100-px window in frame 2
    sel = np.nonzero(k1["octave"] == 0)[0]
    qx, qy = k1["x"][sel], k1["y"][sel]
    qr = np.full(len(sel), 100, np.float32); lv = np.zeros(len(sel), np.int32)
    off, idx = S.features_in_area(gpu_ctx, k2, g, qx, qy, qr, lv, lv)
    ooff, oidx = oracle.grid_candidates(k2, og, qx, qy, qr, lv, lv)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx) and len(idx) > 10 * len(sel)
    lists = [idx[off[i]:off[i + 1]] for i in range(len(sel))]
    m = S.ORBmatcher(0.9, ctx=gpu_ctx)
    bi, bd, cd = m.candidates_top2(d1[sel], d2, lists, want_dist=True)
    obi, obd, ocd = oracle.candidates_top2(d1[sel], d2, lists)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and np.array_equal(cd, ocd)
    good = (bd[:, 0] <= 50) & (bd[:, 0] < bd[:, 1].astype(np.float32) * np.float32(0.9))
    assert good.sum() > len(sel) // 4            # the shifted frame is really being matched
    # random queries with level windows and points outside the grid
    rng = np.random.default_rng(2)
    nq = 5000
    qx = rng.uniform(-30, 1270, nq).astype(np.float32); qy = rng.uniform(-30, 400, nq).astype(np.float32)
    qr = rng.uniform(1, 40, nq).astype(np.float32)
    mn = rng.integers(-1, 5, nq).astype(np.int32); mx = rng.integers(-1, 8, nq).astype(np.int32)
    off, idx = S.features_in_area(gpu_ctx, k2, g, qx, qy, qr, mn, mx)
    ooff, oidx = oracle.grid_candidates(k2, og, qx, qy, qr, mn, mx)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx)
    # line grid (16 x 12, PosInGridLines)
    le, prm = _line_objs(S, oracle, gpu_ctx, 800)
    K, M, D = le.ComputeLsdWithLbd(img1)
    gl = S.GridParams.for_image(16, 12, 0, 1241, 0, 376); ogl = oracle.grid_params(16, 12, 0, 1241, 0, 376)
    qx, qy = M["x"], M["y"]; qr = np.full(len(M), 60, np.float32)
    off, idx = S.features_in_area(gpu_ctx, M, gl, qx, qy, qr, keylines=K)
    ooff, oidx = oracle.grid_candidates(M, ogl, qx, qy, qr, keylines=K)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx) and len(idx) > len(M)


def test_1080p_batch_equals_single_and_frame_sharding(S, oracle, gpu_ctx):
    """BASELINE config 4 (1920x1080 frames, frame i -> GPU i mod G) through its size-independent property: a frame's
    result does not depend on the batch it travels in, so any frame sharding reproduces the single-frame results."""
    from spl_slam_b200 import sharded
    ex = S.ORBextractor(2000, 1.2, 8, 20, 7, ctx=gpu_ctx)
    le, prm = _line_objs(S, oracle, gpu_ctx, 800)
    imgs = np.stack([oracle.synth_image(1920, 1080, 60 + i) for i in range(6)])
    ks, ds = ex.extract_batch(imgs)
    Ks, Ms, Ds = le.extract_batch(imgs)
    ox = oracle.ORBextractor(2000, 1.2, 8, 20, 7)
    for b in (0, 5):                                           # two frames against the oracle (1080p LSD is ~0.5 s each on the CPU)
        ok, od = ox(imgs[b])
        assert np.array_equal(ks[b].view(np.uint8), ok.view(np.uint8)) and np.array_equal(ds[b], od)
        oK, oM, oD = oracle.line_extract(prm, imgs[b])
        assert np.array_equal(Ks[b].view(np.uint8), oK.view(np.uint8)) and np.array_equal(Ds[b], oD)
    for world in (2, 4):                                       # every shard of every world size gives the same per-frame output
        for rank in range(world):
            mine = list(sharded.frame_shard(len(imgs), rank, world))
            assert mine == [i for i in range(len(imgs)) if i % world == rank]
            k2, d2 = ex.extract_batch(imgs[mine])
            K2, M2, D2 = le.extract_batch(imgs[mine])
            for j, i in enumerate(mine):
                assert np.array_equal(k2[j].view(np.uint8), ks[i].view(np.uint8)) and np.array_equal(d2[j], ds[i])
                assert np.array_equal(K2[j].view(np.uint8), Ks[i].view(np.uint8)) and np.array_equal(D2[j], Ds[i])


@pytest.mark.parametrize("k,L,levelsup", [(10, 5, 4), (10, 3, 1), (6, 4, 7)])
def test_bow_transform(S, oracle, gpu_ctx, k, L, levelsup):
    """Frame::ComputeBoW (src/Frame.cc:724-731): DBoW2 tree descent for every ORB descriptor of a frame, then the ordered
    BowVector / FeatureVector maps, against the sequential restatement (bit-exact word ids, weights, node ids, values)."""
    vocab = oracle.synth_vocabulary(k, L, seed=100 + k + L)
    voc = S.ORBVocabulary(gpu_ctx, k, L, *vocab)
    ex = S.ORBextractor(2000, 1.2, 8, 20, 7, ctx=gpu_ctx)
    kp, d = ex(oracle.synth_image(1241, 376, 70))
    rng = np.random.default_rng(0)
    d = d.copy()
    d[:100] = vocab[1][rng.integers(1, len(vocab[0]), 100)]       # exact node descriptors: zero distances and sibling ties
    w, wt, nd = voc.transform_features(d, levelsup)
    ow, owt, ond = oracle.bow_transform(vocab, L, d, levelsup)
    assert np.array_equal(w, ow) and np.array_equal(wt.view(np.uint64), owt.view(np.uint64)) and np.array_equal(nd, ond)
    v, fv = voc.transform(d, levelsup)
    ov, ofv = oracle.bow_vectors(ow, owt, ond)
    assert v == ov and fv == ofv and len(v) > 100
    assert sum(len(x) for x in fv.values()) == int((owt > 0).sum())


def test_frame_pipeline_after_extraction(S, oracle, gpu_ctx):
    """What Frame::Frame does after the extractors (src/Frame.cc:296-362): UndistortKeyPoints / UndistortKeyLines,
    ComputeImageBounds, AssignFeaturesToGrid[Lines] -- then a tracking-style GetFeaturesInArea sweep.  Every stage
    bit-identical to the CPU restatement (cv::undistortPoints itself is pinned against cv2 in test_oracle_vs_cv2)."""
    (fx, fy, cx, cy), dist = (458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)   # EuRoC cam0
    f32 = np.float32
    cam = S.Camera.make(fx, fy, cx, cy, dist)
    img = oracle.synth_image(752, 480, 80)
    ex = S.ORBextractor(1200, 1.2, 8, 20, 7, ctx=gpu_ctx)
    le, prm = _line_objs(S, oracle, gpu_ctx, 200, sigma_scale=0.8, density_th=0.8)
    k, d = ex(img)
    K, M, D = le.ComputeLsdWithLbd(img)
    ku = S.undistort_keypoints(gpu_ctx, cam, k)
    oku = oracle.undistort_keypoints(k, f32(fx), f32(fy), f32(cx), f32(cy), dist)
    assert np.array_equal(ku.view(np.uint8), oku.view(np.uint8)) and not np.array_equal(ku["x"], k["x"])
    Ku, Mu = S.undistort_keylines(gpu_ctx, cam, K, M)
    oKu, oMu = oracle.undistort_keylines(K, M, f32(fx), f32(fy), f32(cx), f32(cy), dist)
    assert np.array_equal(Ku.view(np.uint8), oKu.view(np.uint8)) and np.array_equal(Mu.view(np.uint8), oMu.view(np.uint8))
    # ComputeImageBounds (src/Frame.cc:851-878): undistorted image corners
    corners = np.zeros(4, oracle.KEYPOINT_DTYPE)
    corners["x"] = [0, 752, 0, 752]; corners["y"] = [0, 0, 480, 480]
    cu = S.undistort_keypoints(gpu_ctx, cam, corners)
    ocu = oracle.undistort_keypoints(corners, f32(fx), f32(fy), f32(cx), f32(cy), dist)
    assert np.array_equal(cu.view(np.uint8), ocu.view(np.uint8))
    minx, maxx = min(cu["x"][0], cu["x"][2]), max(cu["x"][1], cu["x"][3])
    miny, maxy = min(cu["y"][0], cu["y"][1]), max(cu["y"][2], cu["y"][3])
    g = S.GridParams.for_image(64, 48, minx, maxx, miny, maxy); og = oracle.grid_params(64, 48, minx, maxx, miny, maxy)
    gl = S.GridParams.for_image(16, 12, minx, maxx, miny, maxy); ogl = oracle.grid_params(16, 12, minx, maxx, miny, maxy)
    rng = np.random.default_rng(6)
    qx = (ku["x"] + rng.uniform(-3, 3, len(ku))).astype(f32); qy = (ku["y"] + rng.uniform(-3, 3, len(ku))).astype(f32)
    qr = (15 * np.power(f32(1.2), ku["octave"])).astype(f32)
    mn = (ku["octave"] - 1).astype(np.int32); mx = ku["octave"].astype(np.int32)
    off, idx = S.features_in_area(gpu_ctx, ku, g, qx, qy, qr, mn, mx)
    ooff, oidx = oracle.grid_candidates(oku, og, qx, qy, qr, mn, mx)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx) and len(idx) >= len(ku)
    off, idx = S.features_in_area(gpu_ctx, Mu, gl, Mu["x"], Mu["y"], np.full(len(Mu), 40, f32), keylines=Ku)
    ooff, oidx = oracle.grid_candidates(oMu, ogl, oMu["x"], oMu["y"], np.full(len(Mu), 40, f32), keylines=oKu)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx)


def test_error_conventions(S, oracle, gpu_ctx):
    """The ABI never truncates silently and never throws across the boundary: capacity / argument problems come back as
    status codes with a message (translated to PlfError by the Python mirror, to std::runtime_error by the C++ shim)."""
    import ctypes as C
    lib, h = gpu_ctx.lib, gpu_ctx.h
    img = oracle.synth_image(640, 480, 2)
    ex = S.ORBextractor(1000, 1.2, 8, 20, 7, ctx=gpu_ctx)
    kps = np.zeros(100, S.KEYPOINT_DTYPE); desc = np.zeros((100, 32), np.uint8); n = C.c_int()
    st = lib.plf_orb_extract(ex.h, img.ctypes.data, 640, 480, 640, kps.ctypes.data, desc.ctypes.data, 100, C.byref(n))
    assert st == S.api.PLF_ERR_CAPACITY and b"capacity" in lib.plf_last_error(h)          # 1000 keypoints do not fit 100 slots
    st = lib.plf_orb_extract(ex.h, img.ctypes.data, 640, 480, 100, kps.ctypes.data, desc.ctypes.data, 100, C.byref(n))
    assert st == S.api.PLF_ERR_INVALID                                                     # stride < width
    k, d = ex(img)                                                                         # the extractor is still usable
    assert len(k) > 900
    # stereo before any extraction on the right extractor: call-order error, not a crash
    ex2 = S.ORBextractor(1000, 1.2, 8, 20, 7, ctx=gpu_ctx)
    with pytest.raises(S.PlfError) as e:
        ex.ComputeStereoMatches(ex2, k, d, k, d, 0.1, 40.0)
    assert e.value.status == S.api.PLF_ERR_STATE
    # lines: unsupported parameters are rejected at construction (refine != 0, > 2 octaves)
    with pytest.raises(S.PlfError):
        S.Lineextractor(200, 2, 1, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=gpu_ctx)
    with pytest.raises(S.PlfError):
        S.Lineextractor(200, 3, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=gpu_ctx)
    # keyline capacity too small for the per-octave quota
    le, prm = _line_objs(S, oracle, gpu_ctx, 200)
    kl = np.zeros(10, S.KEYLINE_DTYPE); mid = np.zeros(10, S.KEYPOINT_DTYPE); ld = np.zeros((10, 32), np.uint8)
    st = lib.plf_line_extract(le.h, img.ctypes.data, 640, 480, 640, kl.ctypes.data, mid.ctypes.data, ld.ctypes.data, 10, C.byref(n))
    assert st == S.api.PLF_ERR_CAPACITY
    K, M, D = le.ComputeLsdWithLbd(img)
    assert len(K) == 200
    # candidate lists with an out-of-range train index, grid with a bad geometry
    m = S.ORBmatcher(0.9, ctx=gpu_ctx)
    with pytest.raises(S.PlfError):
        m.candidates_top2(d[:2], d[:5], [np.array([5], np.int32), np.zeros(0, np.int32)])
    with pytest.raises(S.PlfError):
        S.features_in_area(gpu_ctx, k, S.GridParams(0, 48, 0, 0, 1, 1), k["x"], k["y"], np.ones(len(k), np.float32))


def test_descriptor_distance_batch(S, oracle, gpu_ctx):
    """plf_descriptor_distance (ORBmatcher / Linematcher::DescriptorDistance for n pairs) on the device."""
    import ctypes as C
    rng = np.random.default_rng(9)
    for n, hi in ((1, 256), (1000, 256), (4097, 4)):
        a = rng.integers(0, hi, (n, 32), dtype=np.uint8); b = rng.integers(0, hi, (n, 32), dtype=np.uint8)
        out = np.empty(n, np.int32)
        gpu_ctx.check(gpu_ctx.lib.plf_descriptor_distance(gpu_ctx.h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)))
        want = np.unpackbits(a ^ b, axis=1).sum(1).astype(np.int32)
        assert np.array_equal(out, want)
        for i in range(0, n, max(1, n // 7)):
            assert int(out[i]) == oracle.descriptor_distance(a[i], b[i])


def test_vocab_load_text(S, oracle, gpu_ctx, tmp_path):
    """plf_vocab_load_text (ORBVocabulary::loadFromTextFile, ORBvoc.txt format) + the tree descent on the device."""
    voc = oracle.synth_vocabulary(6, 3, 77)
    parent, desc, weight, leaf = voc
    from test_emu_parity import _write_vocab_text
    path = str(tmp_path / "voc.txt")
    _write_vocab_text(path, 6, 3, voc)
    v = S.ORBVocabulary(gpu_ctx, path=path)
    feats = np.random.default_rng(5).integers(0, 256, (500, 32), dtype=np.uint8)
    w, wt, nd = v.transform_features(feats, 2)
    ow, owt, ond = oracle.bow_transform(voc, 3, feats, 2)
    assert np.array_equal(w, ow) and np.array_equal(wt, owt) and np.array_equal(nd, ond)


def test_line_batch_key_workspace_retry(S, oracle, gpu_ctx):
    """Batches reserve seed-key room for a quarter of the pixels; frames where far more pixels have a defined gradient (noise)
    make the call re-prepare with more room and run again -- the results must not depend on it."""
    rng = np.random.default_rng(12)
    imgs = np.stack([oracle.gauss_blur(rng.integers(0, 256, (120, 160), dtype=np.uint8), 3, 0.8) for _ in range(6)] +
                    [oracle.synth_image(160, 120, 40 + i) for i in range(4)])
    le = S.Lineextractor(60, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=gpu_ctx)
    prm = oracle.line_params(60, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
    for rep in range(2):       # the second call runs in the enlarged workspace
        K, M, D = le.extract_batch(imgs)
        for b in range(len(imgs)):
            oK, oM, oD = oracle.line_extract(prm, imgs[b])
            assert len(K[b]) == len(oK) and np.array_equal(K[b].view(np.uint8), oK.view(np.uint8)) and np.array_equal(D[b], oD)


def test_orb_register_staged_kernels_without_tma():
    """k_fast_cells / k_describe (the kernels taken when a caller's level-0 pitch is not a multiple of 16 bytes, or with
    PLF_NO_TMA=1) give the same keypoints and descriptors as the TMA-fed ones: the variable is read once per process, so the
    comparison runs in a child process."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "import spl_slam_b200 as S\n"
        "from oracle import oracle as O\n"
        "O.build(); ctx = S.Context(0)\n"
        "for (w, h, nf, seed) in ((752, 480, 1200, 1), (640, 480, 1000, 0)):\n"
        "    img = O.synth_image(w, h, seed)\n"
        "    k, d = S.ORBextractor(nf, 1.2, 8, 20, 7, ctx=ctx)(img)\n"
        "    ok, od = O.ORBextractor(nf, 1.2, 8, 20, 7)(img)\n"
        "    assert len(k) == len(ok) and np.array_equal(k.view(np.uint8), ok.view(np.uint8)) and np.array_equal(d, od)\n"
        "rep = ctx.profile_report() if hasattr(ctx, 'profile_report') else {}\n"
        "print('ok')\n" % root)
    env = dict(os.environ, PLF_NO_TMA="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
