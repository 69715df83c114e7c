"""Generate the committed golden vectors (run in the build container, where cv2 4.13.0 is importable).

Sources of truth:
  * ORB keypoints/descriptors and LSD keylines come from oracle/cv2_pipeline.py: the reference's glue logic
    restated in Python over the SAME OpenCV primitives the reference calls (cv2.resize, copyMakeBorder, FAST,
    GaussianBlur, fastAtan2, pyrDown, createLineSegmentDetector) -- independent of the C oracle and the CUDA code.
  * knn2 tables come from cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) (what Linematcher::matchNNR calls).
  * LBD descriptors have no compiled implementation in this container (cv2 has no line_descriptor); they are
    regression pins produced by the C oracle (oracle/orc_lbd.c), marked as such.
Images are stored too (raw uint8 inside the npz) so the GPU box needs neither cv2 nor /root/reference.

  python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cv2  # noqa: E402
from oracle import oracle as O, cv2_pipeline as P  # noqa: E402

cv2.setNumThreads(1)
LSD_OPTS = (0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024)


def main():
    out = {}
    # BASELINE config[0]: 640x480, 1000 ORB, 8 levels x1.2, FAST 20/7 (TUM mono) + TUM line settings
    for tag, (w, h, seed, nf, nl) in {"tum640": (640, 480, 0, 1000, 8), "small320": (320, 240, 3, 300, 5)}.items():
        img = O.synth_image(w, h, seed)
        k, d = P.orb_extract(img, nf, 1.2, nl, 20, 7)
        kl = P.lsd_keylines(img, 2, LSD_OPTS, 0.0)
        prm = O.line_params(600, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0)
        K, M, D = O.line_extract(prm, img)
        out[tag + "_img"] = img
        out[tag + "_orb_params"] = np.array([nf, nl], np.int32)
        out[tag + "_kps"] = k
        out[tag + "_desc"] = d
        out[tag + "_keylines_cv2"] = kl
        out[tag + "_line_kl_oraclepin"] = K
        out[tag + "_line_desc_oraclepin"] = D
        print(tag, len(k), "keypoints", len(kl), "keylines", len(K), "selected lines")
    rng = np.random.default_rng(1234)
    for tag, hi in (("uniform", 256), ("ties", 4)):
        q = rng.integers(0, hi, (200, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (777, 32), dtype=np.uint8)
        idx, dist = P.knn2(q, t)
        out["knn_%s_q" % tag] = q
        out["knn_%s_t" % tag] = t
        out["knn_%s_idx" % tag] = idx
        out["knn_%s_dist" % tag] = dist
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), os.path.getsize(os.path.join(HERE, "golden_v1.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
