"""The C oracle against the REFERENCE'S OWN CODE (oracle/_ref/libref.so: src/ORBextractor.cc, src/Lineextractor.cc,
LSDDetector_custom.cpp and the compute path of binary_descriptor_custom.cpp, compiled unmodified against the test-only cv:: shim
over cv2-pinned primitives; see oracle/ref_build/).  This is what moves parity from "our reading of the source" to "what the
source computes": every comparison below is byte-for-byte unless it says otherwise.

libref.so is built in the build container (the only place /root/reference exists) and travels with the snapshot."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from oracle import ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not (R.available() or os.path.isdir("/root/reference")), reason="oracle/_ref is not built")

SHAPES = [(640, 480, 1000, 0), (752, 480, 1200, 1), (1241, 376, 2000, 2), (320, 240, 300, 7), (97, 83, 120, 5)]


@pytest.fixture(scope="module")
def ref():
    assert R.build() or R.available()
    R.set_heap_mode(1)
    R.set_heap_mode(1, "_generic")
    return R


def _bits(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def _same(a, b):
    return len(a) == len(b) and np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("w,h,nf,seed", SHAPES)
@pytest.mark.parametrize("variant", ["", "_generic"])
def test_orb_extractor(oracle, ref, w, h, nf, seed, variant):
    """ORBextractor::operator(): tables, every pyramid level, keypoints (position, octave, size, angle, response, ORDER) and
    descriptors; the reference runs with monotone heap addresses (heap mode 1)."""
    img = oracle.synth_image(w, h, seed)
    eo, er = oracle.ORBextractor(nf, 1.2, 8, 20, 7), ref.ORBextractor(nf, 1.2, 8, 20, 7, variant=variant)
    assert eo.features_per_level() == er.features_per_level() and eo.umax() == er.umax()
    assert np.array_equal(np.float32(eo.scale_factors()), np.float32(er.scale_factors()))
    ko, do = eo(img)
    kr, dr = er(img)
    for l in range(8):
        assert np.array_equal(eo.level_image(l), er.level_image(l)), "pyramid level %d" % l
    assert _same(ko, kr) and np.array_equal(do, dr)


def test_orb_1080p(oracle, ref):
    img = oracle.synth_image(1920, 1080, 4)
    ko, do = oracle.ORBextractor(2000, 1.2, 8, 20, 7)(img)
    kr, dr = ref.ORBextractor(2000, 1.2, 8, 20, 7)(img)
    assert _same(ko, kr) and np.array_equal(do, dr)


def test_orb_other_parameters(oracle, ref):
    img = oracle.synth_image(480, 360, 9)
    for nf, sf, nl, ini, mn in ((500, 1.1, 4, 30, 5), (2000, 1.5, 3, 12, 7), (50, 1.2, 8, 20, 7), (4000, 1.2, 8, 20, 7)):
        ko, do = oracle.ORBextractor(nf, sf, nl, ini, mn)(img)
        kr, dr = ref.ORBextractor(nf, sf, nl, ini, mn)(img)
        assert _same(ko, kr) and np.array_equal(do, dr), (nf, sf, nl, ini, mn)


def test_octree_tie_order_is_heap_dependent_in_the_reference(oracle, ref):
    """DistributeOctTree sorts (size, node ADDRESS) pairs (ORBextractor.cc:684).  With monotone heap addresses (mode 1) the
    reference is deterministic and equals the oracle's "later created first" rule, order included.  Under glibc malloc
    (mode 0) the same call on the same input returns different keypoint sets depending on what the heap went through before --
    recorded here so that nobody mistakes it for a parity failure."""
    img = oracle.synth_image(752, 480, 1)
    eo = oracle.ORBextractor(1200, 1.2, 8, 20, 7)
    eo(img)
    er = ref.ORBextractor(1200, 1.2, 8, 20, 7)
    libc = C.CDLL("libc.so.6")
    libc.malloc.restype = C.c_void_p
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(5)
    fpl = eo.features_per_level()
    differing_levels = 0
    for l in range(8):
        xs, ys, rr = eo.level_raw(l)
        lh, lw = eo.level_image(l).shape
        args = (xs, ys, rr, 16, lw - 16, 16, lh - 16, fpl[l])
        want = oracle.distribute_octree(*args)
        ref.set_heap_mode(1)
        for _ in range(3):
            assert np.array_equal(er.distribute_octree(*args), want)      # deterministic and equal to the oracle, order included
        ref.set_heap_mode(0)
        outs = []
        for _ in range(4):
            outs.append(er.distribute_octree(*args))
            ps = [libc.malloc(int(s)) for s in rng.integers(16, 400, 200)]     # disturb the free lists
            for i in rng.permutation(len(ps))[:120]:
                libc.free(ps[i])
        ref.set_heap_mode(1)
        for o in outs:
            # whatever the tie order, the result has the same size class and overwhelmingly the same keypoints
            assert abs(len(o) - len(want)) <= 2 and len(set(o) ^ set(want)) <= max(12, len(want) // 8)
        differing_levels += any(set(o) != set(outs[0]) for o in outs)
    assert differing_levels > 0, "glibc-malloc runs agreed everywhere; the heap-dependence note in DESIGN.md needs another look"


LINE_CASES = [(640, 480, 200, 0, 2), (752, 480, 600, 1, 2), (1241, 376, 800, 2, 2), (320, 240, 100, 7, 2), (752, 480, 40, 3, 2),
              (640, 480, 300, 4, 1)]


@pytest.mark.parametrize("w,h,nf,seed,nl", LINE_CASES)
def test_lines(oracle, ref, w, h, nf, seed, nl):
    """LSDDetectorC::detect, BinaryDescriptor::compute and Lineextractor::ComputeLsdWithLbd."""
    img = oracle.synth_image(w, h, seed)
    prm = oracle.line_params(nf, nl, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
    assert oracle.features_per_level_lines(prm) == ref.features_per_level_lines(prm)
    ko = oracle.lsd_detect_keylines(prm, img)
    for variant in ("", "_generic"):
        assert _same(ko, ref.lsd_detect_keylines(prm, img, variant)), "KeyLines (all 17 fields, angle included)"
    do, fo = oracle.lbd_compute(img, ko, True)
    dg, fg = ref.lbd_compute(img, ko, True, "_generic")
    assert np.array_equal(do, dg) and np.array_equal(fo.view(np.uint32), fg.view(np.uint32)), "-O3 build: 72 floats per line, bit for bit"
    dn, fn = ref.lbd_compute(img, ko, True, "")
    # the reference's own flags (-O3 -march=native) let GCC fuse multiply-adds into FMAs: 45 % of the floats move by one ulp and
    # a descriptor bit flips about once in 2e5 bits (test_lines_1080p) -- compiler- and CPU-dependent in the reference itself
    assert np.max(np.abs(fo - fn)) < 1e-6 and _bits(do, dn) <= 2
    for variant in ("", "_generic"):
        a, b = oracle.line_extract(prm, img), ref.line_extract(prm, img, variant)
        assert _same(a[0], b[0]) and _same(a[1], b[1])
        assert np.array_equal(a[2], b[2]) if variant else _bits(a[2], b[2]) <= 2


def test_lines_1080p(oracle, ref):
    img = oracle.synth_image(1920, 1080, 6)
    prm = oracle.line_params(800, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
    a, b = oracle.line_extract(prm, img), ref.line_extract(prm, img, "_generic")
    assert _same(a[0], b[0]) and _same(a[1], b[1]) and np.array_equal(a[2], b[2])
    n = ref.line_extract(prm, img, "")           # -march=native: FMA contraction flips 1 bit of 204 800 on this image
    assert _same(a[0], n[0]) and _bits(a[2], n[2]) <= 4


def test_lines_min_length_and_other_options(oracle, ref):
    img = oracle.synth_image(640, 480, 21)
    for nf, scale, quant, ang, nb, minlen in ((240, 1.05, 2.0, 22.5, 1024, 32.0), (100, 1.0, 2.0, 22.5, 512, 10.0), (600, 0.8, 2.2, 12.5, 1024, 0.0)):
        prm = oracle.line_params(nf, 2, 0, scale, 0.6, quant, ang, 1.0, 0.6, nb, minlen)
        a, b = oracle.line_extract(prm, img), ref.line_extract(prm, img)
        assert _same(a[0], b[0]) and np.array_equal(a[2], b[2]), (nf, scale, quant, ang, nb, minlen)


def test_line_response_ties_follow_std_sort(oracle, ref):
    """Lineextractor.cc:175 uses std::sort (unstable).  On an image with dozens of equal-response lines the selection and
    the output order follow libstdc++'s introsort exactly."""
    img = np.full((480, 640), 30, np.uint8)
    for y in range(20, 440, 60):
        for x in range(20, 600, 80):
            img[y:y + 30, x:x + 50] = 200
    img = oracle.gauss_blur(img, 3, 0.8)
    for nf in (30, 60, 100, 150):
        prm = oracle.line_params(nf, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
        a, b = oracle.line_extract(prm, img), ref.line_extract(prm, img)
        assert _same(a[0], b[0]) and np.array_equal(a[2], b[2])


def test_std_sort_restatement(oracle, ref):
    """oracle.std_sort_desc == the toolchain's std::sort, as permutations: random, tie-heavy, presorted, constant, and
    median-of-three killers that drive introsort into its heapsort fallback."""
    rng = np.random.default_rng(3)
    for trial in range(600):
        n = int(rng.integers(0, 3000)) if trial % 3 else int(rng.integers(0, 40))
        kind = trial % 6
        k = [rng.random(n), rng.integers(0, 4, n), rng.integers(0, max(2, n // 8 + 1), n), np.sort(rng.integers(0, 50, n)),
             np.sort(rng.integers(0, 50, n))[::-1], np.ones(n)][kind]
        k = np.ascontiguousarray(k, np.float32)
        assert np.array_equal(oracle.std_sort_desc(k), ref.std_sort_desc(k))
    heap0 = oracle.lib().orc_std_sort_heap_calls()
    for n in (500, 2000, 4096):
        k = np.zeros(n, np.float32)
        half = n // 2
        for i in range(half):
            k[i] = i + 1 if i % 2 == 0 else half + i + (1 if half % 2 == 0 else 0)
            k[half + i] = 2 * (i + 1)
        assert np.array_equal(oracle.std_sort_desc(-k), ref.std_sort_desc(-k))
    assert oracle.lib().orc_std_sort_heap_calls() > heap0, "the heapsort fallback was not exercised"


def test_matcher_pieces(oracle, ref):
    """DescriptorDistance of both matchers and Linematcher::matchNNR (BFMatcher modelled by the cv2-pinned knn2)."""
    rng = np.random.default_rng(11)
    a = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    for i in range(300):
        d = oracle.descriptor_distance(a[i], b[i])
        assert d == ref.descriptor_distance(a[i], b[i]) == ref.descriptor_distance(a[i], b[i], line=True)
    for hi in (256, 4):
        q = rng.integers(0, hi, (150, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (400, 32), dtype=np.uint8)
        for nnr in (0.6, 0.75, 1.0):
            mo, no = oracle.match_nnr(q, t, nnr)
            mr, nr = ref.match_nnr(q, t, nnr)
            assert no == nr and np.array_equal(mo, mr)


def test_device_libm_models_equal_glibc():
    """spl_slam_b200/csrc/plf_libm.cuh (cosf, sinf, atanf, atan2f as the device computes them) compiled as host code against
    this machine's libm: strided sample here; `libm_check full` (exhaustive, 36 s) is recorded in profiles/r2_libm_check_full.txt."""
    exe = os.path.join(ROOT, "oracle", "_ref", "libm_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fopenmp", os.path.join(ROOT, "oracle", "libm_check.cpp"), "-o", exe, "-lm"])
    out = subprocess.run([exe, "quick"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
