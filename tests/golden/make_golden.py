"""Generate the committed golden vectors (run in the BUILD container: it needs /root/reference for oracle/_ref and cv2 4.13).

Sources of truth, strongest first:
  * `*_ref_*` entries come from the REFERENCE'S OWN CODE: src/ORBextractor.cc, src/Lineextractor.cc, LSDDetector_custom.cpp and
    the compute path of binary_descriptor_custom.cpp compiled unmodified into oracle/_ref/libref.so (oracle/ref_build/), run
    on seeded synthetic images.  ORB runs in heap mode 1 (monotone heap addresses; DistributeOctTree's address tie-break is
    otherwise not reproducible even by the reference itself); the LBD float descriptor comes from the -O3 build without FMA
    contraction (libref_generic.so), the binary descriptor is the same in both builds.
  * `*_cv2` entries come from oracle/cv2_pipeline.py: the reference's glue restated in Python over the real cv2 primitives.
  * knn2 tables come from cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) (what Linematcher::matchNNR calls).
Images are regenerated from their seed by oracle.synth_image (numpy + the cv2-pinned blur); their SHA-1 is stored and checked,
so the GPU box needs neither cv2 nor /root/reference.

  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cv2  # noqa: E402
from oracle import oracle as O, cv2_pipeline as P, ref as R  # noqa: E402

cv2.setNumThreads(1)
LSD_OPTS = (0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024)

# tag: (w, h, seed, ORB nfeatures, ORB levels, line nfeatures)  -- the BASELINE.json shapes
CASES = {"tum640": (640, 480, 0, 1000, 8, 200), "small320": (320, 240, 3, 300, 5, 60), "euroc752": (752, 480, 11, 1200, 8, 200),
         "kitti1241": (1241, 376, 12, 2000, 8, 800)}


def ties_image():
    """A grid of identical rectangles: dozens of lines with EQUAL response, which makes std::sort's instability visible."""
    img = np.full((480, 640), 30, np.uint8)
    for y in range(20, 440, 60):
        for x in range(20, 600, 80):
            img[y:y + 30, x:x + 50] = 200
    return O.gauss_blur(img, 3, 0.8)


def main():
    assert R.build(), "oracle/_ref needs /root/reference"
    R.set_heap_mode(1)
    out = {}
    for tag, (w, h, seed, nf, nl, nlines) in CASES.items():
        img = O.synth_image(w, h, seed)
        out[tag + "_case"] = np.array([w, h, seed, nf, nl, nlines], np.int32)
        out[tag + "_sha1"] = np.frombuffer(hashlib.sha1(img.tobytes()).digest(), np.uint8)
        k, d = R.ORBextractor(nf, 1.2, nl, 20, 7)(img)
        out[tag + "_ref_kps"], out[tag + "_ref_desc"] = k, d
        if w <= 640:
            k2, d2 = P.orb_extract(img, nf, 1.2, nl, 20, 7)
            assert np.array_equal(k2.view(np.uint8), k.view(np.uint8)) and np.array_equal(d2, d), "cv2 pipeline != compiled reference"
            kl = P.lsd_keylines(img, 2, LSD_OPTS, 0.0)
            out[tag + "_keylines_cv2"] = kl
        prm = O.line_params(nlines, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
        out[tag + "_ref_keylines"] = R.lsd_detect_keylines(prm, img)
        K, M, D = R.line_extract(prm, img)
        Kg, Mg, Dg = R.line_extract(prm, img, variant="_generic")
        assert np.array_equal(K.view(np.uint8), Kg.view(np.uint8)) and np.array_equal(D, Dg)
        out[tag + "_ref_line_kl"], out[tag + "_ref_line_mid"], out[tag + "_ref_line_desc"] = K, M, D
        out[tag + "_ref_lbd_float"] = R.lbd_compute(img, K, want_float=True, variant="_generic")[1]
        print(tag, len(k), "keypoints", len(out[tag + "_ref_keylines"]), "keylines", len(K), "selected lines")
    img = ties_image()
    out["ties480_img"] = img
    for q in (30, 100):
        prm = O.line_params(q, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
        K, M, D = R.line_extract(prm, img)
        out["ties480_q%d_ref_line_kl" % q], out["ties480_q%d_ref_line_desc" % q] = K, D
    rng = np.random.default_rng(1234)
    for tag, hi in (("uniform", 256), ("ties", 4)):
        q = rng.integers(0, hi, (200, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (777, 32), dtype=np.uint8)
        idx, dist = P.knn2(q, t)
        out["knn_%s_q" % tag] = q
        out["knn_%s_t" % tag] = t
        out["knn_%s_idx" % tag] = idx
        out["knn_%s_dist" % tag] = dist
        m, n = R.match_nnr(q, t, 0.75)
        out["knn_%s_ref_nnr" % tag] = m
    path = os.path.join(HERE, "golden_v2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
