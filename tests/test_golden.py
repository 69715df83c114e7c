"""Golden vectors (tests/golden/golden_v2.npz, made by tests/golden/make_golden.py): outputs of the REFERENCE'S OWN CODE
(compiled unmodified into oracle/_ref in the build container) on seeded synthetic images at the BASELINE shapes, plus cv2
BFMatcher tables.  The oracle is checked on CPU, the CUDA path on the GPU box (where /root/reference does not exist)."""
import hashlib
import os
import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.npz"))
CASES = ["tum640", "small320", "euroc752", "kitti1241"]


def _lsd(n):
    return dict(nfeatures=n, nlevels=2, refine=0, scale=1.1, sigma_scale=0.6, quant=2.2, ang_th=12.5, log_eps=1.0,
                density_th=0.6, n_bins=1024, min_line_length=0.0)


def _case(oracle, tag):
    w, h, seed, nf, nl, nlines = (int(v) for v in G[tag + "_case"])
    img = oracle.synth_image(w, h, seed)
    assert hashlib.sha1(img.tobytes()).digest() == G[tag + "_sha1"].tobytes(), "synthetic image generator changed"
    return img, nf, nl, nlines


def _same(a, b):
    return len(a) == len(b) and np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_golden(oracle, tag):
    img, nf, nl, nlines = _case(oracle, tag)
    k, d = oracle.ORBextractor(nf, 1.2, nl, 20, 7)(img)
    assert _same(k, G[tag + "_ref_kps"]) and np.array_equal(d, G[tag + "_ref_desc"])
    prm = oracle.line_params(**_lsd(nlines))
    assert _same(oracle.lsd_detect_keylines(prm, img), G[tag + "_ref_keylines"])
    if tag + "_keylines_cv2" in G:
        assert _same(oracle.lsd_detect_keylines(oracle.line_params(**_lsd(nlines)), img), G[tag + "_keylines_cv2"])
    K, M, D = oracle.line_extract(prm, img)
    assert _same(K, G[tag + "_ref_line_kl"]) and _same(M, G[tag + "_ref_line_mid"]) and np.array_equal(D, G[tag + "_ref_line_desc"])
    _, F = oracle.lbd_compute(img, K, want_float=True)
    assert np.array_equal(F.view(np.uint32), G[tag + "_ref_lbd_float"].view(np.uint32))    # 72 floats per line, bit for bit


@pytest.mark.parametrize("quota", [30, 100])
def test_oracle_response_ties_golden(oracle, quota):
    """std::sort's unstable order among lines of equal response (Lineextractor.cc:175)."""
    K, M, D = oracle.line_extract(oracle.line_params(**_lsd(quota)), G["ties480_img"])
    assert _same(K, G["ties480_q%d_ref_line_kl" % quota]) and np.array_equal(D, G["ties480_q%d_ref_line_desc" % quota])


@pytest.mark.parametrize("tag", ["uniform", "ties"])
def test_oracle_knn_matches_golden(oracle, tag):
    i, d = oracle.knn2(G["knn_%s_q" % tag], G["knn_%s_t" % tag])
    assert np.array_equal(i, G["knn_%s_idx" % tag]) and np.array_equal(d, G["knn_%s_dist" % tag])
    m, n = oracle.match_nnr(G["knn_%s_q" % tag], G["knn_%s_t" % tag], 0.75)
    assert np.array_equal(m, G["knn_%s_ref_nnr" % tag])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_gpu_matches_reference_golden(gpu_ctx, oracle, tag):
    import spl_slam_b200 as S
    img, nf, nl, nlines = _case(oracle, tag)
    k, d = S.ORBextractor(nf, 1.2, nl, 20, 7, ctx=gpu_ctx)(img)
    assert _same(k, G[tag + "_ref_kps"]) and np.array_equal(d, G[tag + "_ref_desc"])
    le = S.Lineextractor(nlines, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=gpu_ctx)
    kl = le.lsd_detect(img)
    g = G[tag + "_ref_keylines"]
    assert len(kl) == len(g) and np.array_equal(kl["numOfPixels"], g["numOfPixels"]) and np.array_equal(kl["octave"], g["octave"])
    for f in ("startPointX", "startPointY", "endPointX", "endPointY", "sPointInOctaveX", "sPointInOctaveY", "ePointInOctaveX",
              "ePointInOctaveY", "lineLength"):
        assert np.allclose(kl[f], g[f], rtol=0, atol=1e-3)       # stated tolerance for line end points
    assert np.allclose(kl["angle"], g["angle"], rtol=0, atol=1e-5)   # stated tolerance for KeyLine.angle
    K, M, D = le.ComputeLsdWithLbd(img)
    assert np.array_equal(D, G[tag + "_ref_line_desc"])              # binary LBD descriptors: bit-exact
    # in fact everything is bit-identical on these images; keep that visible
    assert _same(kl, g) and _same(K, G[tag + "_ref_line_kl"]) and _same(M, G[tag + "_ref_line_mid"])
    D2, F = le.lbd_compute(img, K, want_float=True)
    assert np.array_equal(F.view(np.uint32), G[tag + "_ref_lbd_float"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("quota", [30, 100])
def test_gpu_response_ties_golden(gpu_ctx, quota):
    import spl_slam_b200 as S
    le = S.Lineextractor(quota, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=gpu_ctx)
    K, M, D = le.ComputeLsdWithLbd(G["ties480_img"])
    assert _same(K, G["ties480_q%d_ref_line_kl" % quota]) and np.array_equal(D, G["ties480_q%d_ref_line_desc" % quota])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["uniform", "ties"])
def test_gpu_knn_matches_golden(gpu_ctx, tag):
    import spl_slam_b200 as S
    lm = S.Linematcher(0.75, ctx=gpu_ctx)
    i, d = lm.knnMatch2(G["knn_%s_q" % tag], G["knn_%s_t" % tag])
    assert np.array_equal(i, G["knn_%s_idx" % tag]) and np.array_equal(d, G["knn_%s_dist" % tag])
    m, n = lm.matchNNR(G["knn_%s_q" % tag], G["knn_%s_t" % tag])
    assert np.array_equal(m, G["knn_%s_ref_nnr" % tag])
