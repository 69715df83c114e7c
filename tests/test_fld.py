"""FLD branch (System.usingLsdFeature: 0): plf_fld_* against the REFERENCE'S OWN Lineextractor::detect / ComputeFldWithLbd
(src/Lineextractor.cc:242-336, 413-905, compiled unmodified into oracle/_ref over the cv2-pinned Canny / fitLine models).
CPU: the kernels through the test-only emulation.  GPU: the product library."""
import os
import sys
import numpy as np
import pytest

from oracle import ref as R

@pytest.fixture(scope="module")
def S():
    import spl_slam_b200 as S
    return S


CASES = [(320, 240, 7, 1, 10, 1.414213562, 240), (640, 480, 0, 1, 15, 1.732, 240), (640, 480, 3, 2, 15, 1.732, 240),
         (752, 480, 1, 2, 15, 1.732, 100), (1241, 376, 2, 2, 15, 1.732, 240)]


def _check(S, oracle, ctx, cases, exact_endpoints):
    assert R.available() or R.build(), "oracle/_ref is not built"
    for (w, h, seed, nl, lt, dt, nf) in cases:
        img = oracle.synth_image(w, h, seed)
        fe = S.FldLineextractor(nf, nl, 1.05, lt, dt, 50.0, 100.0, 3, False, ctx=ctx)
        P = R.fld_params(nf, nl, 1.05, lt, dt, 50.0, 100.0, 3, False)
        a, b = fe.detect(img), R.fld_detect(P, img)
        assert a.shape == b.shape and len(a) > 20
        assert np.allclose(a, b, rtol=0, atol=1e-3)                      # stated tolerance for line end points
        if exact_endpoints:
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        K, M, D = fe.ComputeFldWithLbd(img)
        rK, rM, rD = R.fld_extract(P, img)
        assert len(K) == len(rK)
        for f in K.dtype.names:
            if f in ("pt_x", "pt_y"):
                continue                # left uninitialised by the reference (a local KeyLine whose pt is never set, :303-322)
            if K[f].dtype.kind == "f" and not exact_endpoints:
                assert np.allclose(K[f], rK[f], rtol=1e-6, atol=1e-3), f
            else:
                assert np.array_equal(K[f], rK[f]), f
        assert np.array_equal(M["octave"], rM["octave"]) and np.allclose(M["x"], rM["x"], atol=1e-3) and np.allclose(M["y"], rM["y"], atol=1e-3)
        assert np.array_equal(D, rD), "LBD descriptors of the FLD lines"
        fe.close()


def test_emu_fld_equals_reference(S, oracle, emu_lib):
    ctx = S.Context(0, emu_lib)
    # the emulation runs every CUDA thread as a pthread: small images keep the CPU suite short; the full shapes run on the GPU
    _check(S, oracle, ctx, [(192, 144, 9, 2, 10, 1.414213562, 40)], exact_endpoints=True)


def test_fld_rejects_what_it_does_not_implement(S, emu_lib):
    ctx = S.Context(0, emu_lib)
    with pytest.raises(S.PlfError):
        S.FldLineextractor(240, 1, 1.05, 10, 1.414, 50.0, 100.0, 3, True, ctx=ctx)      # do_merge
    with pytest.raises(S.PlfError):
        S.FldLineextractor(240, 1, 1.05, 10, 1.414, 50.0, 100.0, 5, False, ctx=ctx)     # aperture 5
    with pytest.raises(S.PlfError):
        S.FldLineextractor(240, 1, 1.05, 0, 1.414, 50.0, 100.0, 3, False, ctx=ctx)      # CV_Assert(_length_threshold > 0)


@pytest.mark.gpu
def test_gpu_fld_equals_reference(S, oracle, gpu_ctx):
    # double cos / sin / atan2 of CUDA and glibc may differ in the last ulp: end points within the stated tolerance (they are
    # bit-identical on these images, which the second pass asserts as well)
    _check(S, oracle, gpu_ctx, CASES, exact_endpoints=False)
    _check(S, oracle, gpu_ctx, CASES[:2], exact_endpoints=True)
