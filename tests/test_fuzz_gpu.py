"""GPU (B200): randomized differential tests, CUDA path (through the C ABI) against the C oracle on many small images.

The region-growing kernels take shortcuts that are argued to be exact (batched acceptance under a drift bound, speculative
growth with in-order commit, plf_line_kernels.cuh / plf_lsd_grow_cta.cuh); the named BASELINE shapes exercise them on a few
dozen images.  Here: >= 20 000 small images of five families (noise at several amplitudes and smoothness, straight edges at
random orientation, concentric rings -- curved edges keep the region angle near the tolerance --, rectangles, gratings) through
the batch entry point, a few hundred mid-size images through the single-frame call (the giant-component CTA path), and ORB on
small images.  Everything is compared byte for byte (keyline records, mid points, LBD / rBRIEF descriptors)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LSD_TUM = dict(refine=0, scale=1.1, sigma_scale=0.6, quant=2.2, ang_th=12.5, log_eps=1.0, density_th=0.6, n_bins=1024)


def _smooth(batch, sigma):
    if sigma <= 0:
        return batch
    from scipy.ndimage import gaussian_filter
    out = gaussian_filter(batch.astype(np.float32), sigma=(0, sigma, sigma), mode="nearest")
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def make_family(kind, n, w, h, rng):
    """n images (n, h, w) u8 of one family; vectorised over the batch."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    if kind == "noise":
        amp = rng.choice([8, 32, 128, 255], size=(n, 1, 1))
        base = rng.integers(0, 256, size=(n, 1, 1))
        img = np.clip(base + (rng.random((n, h, w)) - 0.5) * amp, 0, 255).astype(np.uint8)
        return _smooth(img, float(rng.choice([0.0, 0.7, 1.5])))
    if kind == "edges":        # two or three half planes at random orientation and contrast, light noise
        img = np.full((n, h, w), 0.0, np.float32)
        for _ in range(3):
            th = rng.random((n, 1, 1)) * np.pi
            c = rng.random((n, 1, 1)) * (w + h) / 2
            lvl = rng.integers(-120, 121, size=(n, 1, 1))
            img += np.where(xx[None] * np.cos(th) + yy[None] * np.sin(th) > c, lvl, 0)
        img = np.clip(128 + img + (rng.random((n, h, w)) - 0.5) * rng.choice([0, 6, 20]), 0, 255).astype(np.uint8)
        return _smooth(img, float(rng.choice([0.0, 0.8])))
    if kind == "rings":        # concentric rings: curved edges, the region angle drifts to the tolerance
        cx = rng.random((n, 1, 1)) * w; cy = rng.random((n, 1, 1)) * h
        per = 6 + rng.random((n, 1, 1)) * 30
        r = np.sqrt((xx[None] - cx) ** 2 + (yy[None] - cy) ** 2)
        img = 128 + rng.integers(30, 127, size=(n, 1, 1)) * np.sign(np.sin(2 * np.pi * r / per))
        img = np.clip(img + (rng.random((n, h, w)) - 0.5) * 8, 0, 255).astype(np.uint8)
        return _smooth(img, float(rng.choice([0.6, 1.2])))
    if kind == "rects":
        img = rng.integers(0, 256, size=(n, 1, 1)).repeat(h, 1).repeat(w, 2).astype(np.uint8)
        for i in range(n):
            for _ in range(int(rng.integers(1, 8))):
                x0, y0 = int(rng.integers(0, w - 4)), int(rng.integers(0, h - 4))
                img[i, y0:y0 + int(rng.integers(3, h // 2)), x0:x0 + int(rng.integers(3, w // 2))] = int(rng.integers(0, 256))
        return _smooth(img, float(rng.choice([0.0, 0.6, 1.0])))
    if kind == "gratings":     # oriented sine gratings of two frequencies on a ramp
        th = rng.random((n, 1, 1)) * np.pi
        f1 = 0.05 + rng.random((n, 1, 1)) * 0.4
        u = xx[None] * np.cos(th) + yy[None] * np.sin(th)
        img = 128 + 60 * np.sin(u * f1) + 40 * np.sin((xx[None] * np.sin(th) - yy[None] * np.cos(th)) * f1 * 0.37) + 0.3 * (xx[None] - w / 2)
        return np.clip(img, 0, 255).astype(np.uint8)
    raise ValueError(kind)


FAMILIES = ("noise", "edges", "rings", "rects", "gratings")


@pytest.fixture(scope="module")
def S_mod():
    import spl_slam_b200 as S
    return S


def _line_objs(S, oracle, ctx, nf, **kw):
    o = dict(LSD_TUM); o.update(kw)
    le = S.Lineextractor(nf, 2, o["refine"], o["scale"], o["sigma_scale"], o["quant"], o["ang_th"], o["log_eps"], o["density_th"], o["n_bins"], 0.0, ctx=ctx)
    prm = oracle.line_params(nf, 2, o["refine"], o["scale"], o["sigma_scale"], o["quant"], o["ang_th"], o["log_eps"], o["density_th"], o["n_bins"], 0.0)
    return le, prm


def _oracle_lines(oracle, prm, imgs):
    with ThreadPoolExecutor(os.cpu_count() or 4) as pool:        # the C oracle releases the GIL
        return list(pool.map(lambda im: oracle.line_extract(prm, im), imgs))


def _compare_lines(got, want, tag):
    K, M, D = got
    oK, oM, oD = want
    assert len(K) == len(oK), "%s: %d lines, oracle %d" % (tag, len(K), len(oK))
    assert np.array_equal(K.view(np.uint8), oK.view(np.uint8)), "%s: keyline records differ" % tag
    assert np.array_equal(M.view(np.uint8), oM.view(np.uint8)), "%s: mid points differ" % tag
    assert np.array_equal(D, oD), "%s: LBD descriptors differ" % tag
    return len(K)


def test_fuzz_lines_10k_small_images(S_mod, oracle, gpu_ctx):
    w, h, per_family = 96, 80, 4096
    le, prm = _line_objs(S_mod, oracle, gpu_ctx, 100)
    rng = np.random.default_rng(20261019)
    total_imgs = total_lines = 0
    for kind in FAMILIES:
        imgs = make_family(kind, per_family, w, h, rng)
        want = _oracle_lines(oracle, prm, imgs)
        for b0 in range(0, per_family, 1024):
            Ks, Ms, Ds = le.extract_batch(imgs[b0:b0 + 1024])
            for i in range(len(Ks)):
                total_lines += _compare_lines((Ks[i], Ms[i], Ds[i]), want[b0 + i], "%s image %d" % (kind, b0 + i))
        total_imgs += per_family
    assert total_imgs >= 20000 and total_lines > 40000, (total_imgs, total_lines)


def test_fuzz_lines_single_frame_giant_components(S_mod, oracle, gpu_ctx):
    """320x240 frames one at a time: connected edge networks of thousands of pixels -> k_lsd_grow_cta (speculation + commit)."""
    w, h = 320, 240
    rng = np.random.default_rng(7)
    total = 0
    for kw in ({}, dict(sigma_scale=0.8, density_th=0.8), dict(scale=1.0)):
        le, prm = _line_objs(S_mod, oracle, gpu_ctx, 300, **kw)
        imgs = np.concatenate([make_family(k, 32, w, h, rng) for k in FAMILIES] + [np.stack([oracle.synth_image(w, h, 100 + i) for i in range(40)])])
        want = _oracle_lines(oracle, prm, imgs)
        for i, im in enumerate(imgs):
            total += _compare_lines(le.ComputeLsdWithLbd(im), want[i], "settings %s image %d" % (kw, i))
    assert total > 10000


def test_fuzz_orb_small_images(S_mod, oracle, gpu_ctx):
    w, h, n = 160, 120, 256
    ex = S_mod.ORBextractor(150, 1.2, 4, 20, 7, ctx=gpu_ctx)
    ox = oracle.ORBextractor(150, 1.2, 4, 20, 7)
    rng = np.random.default_rng(11)
    total = 0
    for kind in FAMILIES:
        imgs = make_family(kind, n, w, h, rng)
        ks, ds = ex.extract_batch(imgs)
        for i in range(n):
            ok, od = ox(imgs[i])
            assert len(ks[i]) == len(ok), "%s image %d: %d keypoints, oracle %d" % (kind, i, len(ks[i]), len(ok))
            assert np.array_equal(ks[i].view(np.uint8), ok.view(np.uint8)) and np.array_equal(ds[i], od), "%s image %d" % (kind, i)
            total += len(ok)
    assert total > 10000
