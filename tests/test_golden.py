"""Golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from cv2 4.13 primitives +
the Python restatement of the reference glue): the oracle on CPU, and the CUDA path on the GPU box."""
import os
import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
LSD = dict(nfeatures=600, nlevels=2, refine=0, scale=1.1, sigma_scale=0.6, quant=2.2, ang_th=12.5, log_eps=1.0,
           density_th=0.6, n_bins=1024, min_line_length=0.0)


@pytest.mark.parametrize("tag", ["tum640", "small320"])
def test_oracle_matches_golden(oracle, tag):
    img = G[tag + "_img"]
    nf, nl = (int(v) for v in G[tag + "_orb_params"])
    k, d = oracle.ORBextractor(nf, 1.2, nl, 20, 7)(img)
    assert np.array_equal(k.view(np.uint8), G[tag + "_kps"].view(np.uint8)) and np.array_equal(d, G[tag + "_desc"])
    kl = oracle.lsd_detect_keylines(oracle.line_params(**LSD), img)
    assert np.array_equal(kl.view(np.uint8), G[tag + "_keylines_cv2"].view(np.uint8))
    K, M, D = oracle.line_extract(oracle.line_params(**LSD), img)
    assert np.array_equal(K.view(np.uint8), G[tag + "_line_kl_oraclepin"].view(np.uint8))
    assert np.array_equal(D, G[tag + "_line_desc_oraclepin"])


@pytest.mark.parametrize("tag", ["uniform", "ties"])
def test_oracle_knn_matches_golden(oracle, tag):
    i, d = oracle.knn2(G["knn_%s_q" % tag], G["knn_%s_t" % tag])
    assert np.array_equal(i, G["knn_%s_idx" % tag]) and np.array_equal(d, G["knn_%s_dist" % tag])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["tum640", "small320"])
def test_gpu_matches_golden(gpu_ctx, tag):
    import spl_slam_b200 as S
    img = G[tag + "_img"]
    nf, nl = (int(v) for v in G[tag + "_orb_params"])
    k, d = S.ORBextractor(nf, 1.2, nl, 20, 7, ctx=gpu_ctx)(img)
    assert np.array_equal(k.view(np.uint8), G[tag + "_kps"].view(np.uint8)) and np.array_equal(d, G[tag + "_desc"])
    le = S.Lineextractor(LSD["nfeatures"], 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0, ctx=gpu_ctx)
    kl = le.lsd_detect(img)
    g = G[tag + "_keylines_cv2"]
    assert len(kl) == len(g) and np.array_equal(kl["numOfPixels"], g["numOfPixels"])
    for f in ("startPointX", "startPointY", "endPointX", "endPointY"):
        assert np.allclose(kl[f], g[f], rtol=0, atol=1e-3)       # stated tolerance for line end points
    assert np.allclose(kl["angle"], g["angle"], rtol=0, atol=1e-5)
    K, M, D = le.ComputeLsdWithLbd(img)
    assert np.array_equal(D, G[tag + "_line_desc_oraclepin"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["uniform", "ties"])
def test_gpu_knn_matches_golden(gpu_ctx, tag):
    import spl_slam_b200 as S
    i, d = S.Linematcher(0.75, ctx=gpu_ctx).knnMatch2(G["knn_%s_q" % tag], G["knn_%s_t" % tag])
    assert np.array_equal(i, G["knn_%s_idx" % tag]) and np.array_equal(d, G["knn_%s_dist" % tag])
