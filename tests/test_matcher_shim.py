"""The drop-in matcher classes (include/plf_matcher_shim.hpp: PL_SLAM::ORBmatcher::SearchForInitialization,
PL_SLAM::Linematcher::SearchByKNN / SearchForTriangulation) against the REFERENCE'S OWN functions, cut by line range from
src/ORBmatcher.cc / src/Linematcher.cc / src/Frame.cc into oracle/_ref, on the same mock Frame / KeyFrame objects
(tests/shim/mock_slam.hpp).  CPU: the header compiles and links.  GPU: results are identical."""
import os
import struct
import subprocess
import numpy as np
import pytest

from oracle import ref as R
from oracle.oracle import KEYPOINT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "shim")


def _build(tmp):
    exe = os.path.join(tmp, "matcher_main")
    lib = os.path.join(ROOT, "spl_slam_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", SHIM,
                           os.path.join(SHIM, "matcher_main.cpp"), "-o", exe, "-L", lib, "-lplf", "-Wl,-rpath," + lib])
    return exe


def test_matcher_shim_compiles_and_links(tmp_path):
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(_build(str(tmp_path)))


def _noisy(rng, d, flips):
    out = d.copy()
    for i in range(len(out)):
        for b in rng.integers(0, 256, int(rng.integers(0, flips + 1))):
            out[i, b >> 3] ^= 1 << (b & 7)
    return out


def _cases(seed):
    rng = np.random.default_rng(seed)
    # ---- SearchForInitialization: F2 = F1 moved a little, descriptors with a few flipped bits, some duplicates ----
    n1, n2 = 700, 760
    bounds = np.array([0.0, 752.0, 0.0, 480.0], np.float32)
    k1 = np.zeros(n1, KEYPOINT_DTYPE)
    k1["x"] = rng.uniform(20, 730, n1).astype(np.float32); k1["y"] = rng.uniform(20, 460, n1).astype(np.float32)
    k1["octave"] = (rng.random(n1) < 0.25).astype(np.int32) * rng.integers(1, 4, n1)
    k1["angle"] = rng.uniform(0, 360, n1).astype(np.float32); k1["size"] = 31; k1["class_id"] = -1
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    src = rng.integers(0, n1, n2)
    k2 = k1[src].copy()
    k2["x"] += rng.uniform(-6, 6, n2).astype(np.float32); k2["y"] += rng.uniform(-6, 6, n2).astype(np.float32)
    k2["octave"] = 0
    rot = np.where(rng.random(n2) < 0.8, 12.0, rng.uniform(0, 360, n2))        # a dominant rotation + outliers for the histogram
    k2["angle"] = np.mod(k1["angle"][src] - rot, 360).astype(np.float32)
    d2 = _noisy(rng, d1[src], 40)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    init = dict(k1=k1, d1=d1, k2=k2, d2=d2, bounds=bounds, prev=prev, window=10, nnr=np.float32(0.9), ori=1)
    # ---- SearchByKNN ----
    nkf, nf = 300, 340
    dkf = rng.integers(0, 256, (nkf, 32), dtype=np.uint8)
    srcl = rng.integers(0, nkf, nf)
    df = _noisy(rng, dkf[srcl], 30)
    state = rng.choice([0, 1, 1, 1, 2], nkf).astype(np.uint8)
    knn = dict(dkf=dkf, df=df, state=state, mllen=rng.uniform(20, 200, nkf).astype(np.float32), flen=rng.uniform(20, 200, nf).astype(np.float32),
               nnr=np.float32(0.75), checklen=1, lengtherr=np.float32(0.25))
    # ---- SearchForTriangulation ----
    t1, t2 = 260, 280
    dt1 = rng.integers(0, 256, (t1, 32), dtype=np.uint8)
    srct = rng.integers(0, t1, t2)
    dt2 = _noisy(rng, dt1[srct], 30)
    m1 = np.zeros(t1, KEYPOINT_DTYPE); m2 = np.zeros(t2, KEYPOINT_DTYPE)
    m1["x"] = rng.uniform(0, 752, t1).astype(np.float32); m1["y"] = rng.uniform(0, 480, t1).astype(np.float32); m1["octave"] = rng.integers(0, 2, t1)
    m2["x"] = (m1["x"][srct] + rng.uniform(-30, 30, t2)).astype(np.float32); m2["y"] = (m1["y"][srct] + rng.uniform(-3, 3, t2)).astype(np.float32)
    m2["octave"] = rng.integers(0, 2, t2)
    cam = np.array([435.2, 435.2, 367.4, 252.2], np.float32)
    # pure x translation: F12 = K^-T [t]x K^-1 gives horizontal epipolar lines; the epipole is far away
    Kinv = np.linalg.inv(np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float64))
    tx = np.array([[0, 0, 0], [0, 0, -0.2], [0, 0.2, 0]], np.float64)
    F12 = (Kinv.T @ tx @ Kinv).astype(np.float32)
    pose = np.concatenate([np.array([0.2, 0.0, 0.0]), np.eye(3).ravel(), np.array([0.0, 0.0, 1.0])]).astype(np.float32)
    tri = dict(d1=dt1, d2=dt2, h1=(rng.random(t1) < 0.5).astype(np.uint8), h2=(rng.random(t2) < 0.5).astype(np.uint8), m1=m1, m2=m2,
               scale=np.array([1.0, 1.1], np.float32), sigma2=np.array([1.0, 1.21], np.float32), cam=cam, pose=pose, F12=F12, nnr=np.float32(0.8))
    return init, knn, tri


def _write(path, init, knn, tri):
    with open(path, "wb") as f:
        f.write(struct.pack("<iiiif", len(init["k1"]), len(init["k2"]), init["window"], init["ori"], float(init["nnr"])))
        f.write(init["bounds"].tobytes())
        f.write(init["k1"].tobytes()); f.write(init["d1"].tobytes()); f.write(init["k2"].tobytes()); f.write(init["d2"].tobytes())
        f.write(init["prev"].tobytes())
        f.write(struct.pack("<iiiff", len(knn["dkf"]), len(knn["df"]), knn["checklen"], float(knn["nnr"]), float(knn["lengtherr"])))
        f.write(knn["dkf"].tobytes()); f.write(knn["df"].tobytes()); f.write(knn["state"].tobytes()); f.write(knn["mllen"].tobytes()); f.write(knn["flen"].tobytes())
        f.write(struct.pack("<iiif", len(tri["d1"]), len(tri["d2"]), len(tri["scale"]), float(tri["nnr"])))
        f.write(tri["d1"].tobytes()); f.write(tri["d2"].tobytes()); f.write(tri["h1"].tobytes()); f.write(tri["h2"].tobytes())
        f.write(tri["m1"].tobytes()); f.write(tri["m2"].tobytes()); f.write(tri["scale"].tobytes()); f.write(tri["sigma2"].tobytes())
        f.write(tri["cam"].tobytes()); f.write(tri["pose"].tobytes()); f.write(tri["F12"].tobytes())


def test_reference_matchers_run_on_the_mock_objects():
    """CPU sanity of the oracle/_ref side: the reference's functions produce matches on the synthetic cases."""
    if not (R.available() or R.build()):
        pytest.skip("oracle/_ref is not built")
    init, knn, tri = _cases(1)
    n, m12, prev = R.orb_search_for_initialization(init["k1"], init["d1"], init["k2"], init["d2"], init["bounds"], init["prev"], init["window"],
                                                   init["nnr"], init["ori"])
    assert n == int((m12 >= 0).sum()) and n > 100
    r, ml = R.line_search_by_knn(knn["dkf"], knn["df"], knn["state"], knn["mllen"], knn["flen"], knn["nnr"], knn["checklen"], knn["lengtherr"])
    assert (ml >= 0).sum() > 30
    pairs = R.line_search_for_triangulation(tri["d1"], tri["d2"], tri["h1"], tri["h2"], tri["m1"], tri["m2"], tri["scale"], tri["sigma2"],
                                            tri["cam"], tri["pose"], tri["F12"], tri["nnr"])
    assert len(pairs) > 20


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_matcher_shim_equals_reference(tmp_path, seed):
    assert R.available(), "oracle/_ref/libref.so must travel with the snapshot"
    exe = _build(str(tmp_path))
    init, knn, tri = _cases(seed)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write(inp, init, knn, tri)
    subprocess.check_call([exe, inp, out])
    buf = open(out, "rb").read()
    off = 0
    n1, nf = len(init["k1"]), len(knn["df"])
    n = struct.unpack_from("<i", buf, off)[0]; off += 4
    m12 = np.frombuffer(buf, np.int32, n1, off); off += 4 * n1
    prev = np.frombuffer(buf, np.float32, 2 * n1, off).reshape(n1, 2); off += 8 * n1
    rn, rm12, rprev = R.orb_search_for_initialization(init["k1"], init["d1"], init["k2"], init["d2"], init["bounds"], init["prev"], init["window"],
                                                      init["nnr"], init["ori"])
    assert n == rn and np.array_equal(m12, rm12) and np.array_equal(prev.view(np.uint32), rprev.view(np.uint32))
    kn = struct.unpack_from("<i", buf, off)[0]; off += 4
    ml = np.frombuffer(buf, np.int32, nf, off); off += 4 * nf
    rr, rml = R.line_search_by_knn(knn["dkf"], knn["df"], knn["state"], knn["mllen"], knn["flen"], knn["nnr"], knn["checklen"], knn["lengtherr"])
    assert kn == rr and np.array_equal(ml, rml)
    tn = struct.unpack_from("<i", buf, off)[0]; off += 4
    pairs = np.frombuffer(buf, np.int32, 2 * tn, off).reshape(tn, 2)
    rpairs = R.line_search_for_triangulation(tri["d1"], tri["d2"], tri["h1"], tri["h2"], tri["m1"], tri["m2"], tri["scale"], tri["sigma2"],
                                             tri["cam"], tri["pose"], tri["F12"], tri["nnr"])
    assert tn == len(rpairs) and np.array_equal(pairs, rpairs)
