"""The C++ drop-in header (include/plf_slam_shim.hpp) compiles against layout-compatible mock OpenCV types (CPU)
and, on the GPU box, produces the same results as the oracle."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "shim")


def _build(tmp):
    exe = os.path.join(tmp, "shim_main")
    lib = os.path.join(ROOT, "spl_slam_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", SHIM,
                           os.path.join(SHIM, "shim_main.cpp"), "-o", exe, "-L", lib, "-lplf", "-Wl,-rpath," + lib])
    return exe


def test_shim_compiles_and_links(tmp_path):
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(_build(str(tmp_path)))


@pytest.mark.gpu
def test_shim_results_match_oracle(tmp_path, oracle):
    exe = _build(str(tmp_path))
    img = oracle.synth_image(640, 480, 21)
    raw, out = str(tmp_path / "in.raw"), str(tmp_path / "out.bin")
    img.tofile(raw)
    subprocess.check_call([exe, raw, "640", "480", out])
    buf = open(out, "rb").read()
    hdr = np.frombuffer(buf, np.int32, 8)
    nk, nl, nm, d01, k2, levels, p1w, p1h = (int(v) for v in hdr)
    off = 32
    kps = np.frombuffer(buf, oracle.KEYPOINT_DTYPE, nk, off); off += nk * 28
    desc = np.frombuffer(buf, np.uint8, nk * 32, off).reshape(nk, 32); off += nk * 32
    kl = np.frombuffer(buf, oracle.KEYLINE_DTYPE, nl, off); off += nl * 68
    ld = np.frombuffer(buf, np.uint8, nl * 32, off).reshape(nl, 32); off += nl * 32
    m12 = np.frombuffer(buf, np.int32, nk, off); off += nk * 4
    uR = np.frombuffer(buf, np.float32, nk, off); off += nk * 4
    dep = np.frombuffer(buf, np.float32, nk, off)
    ok, od = oracle.ORBextractor(500, 1.2, 6, 20, 7)(img)
    assert nk == len(ok) and np.array_equal(kps.view(np.uint8), ok.view(np.uint8)) and np.array_equal(desc, od)
    oK, oM, oD = oracle.line_extract(oracle.line_params(100, 2, 0, 1.1, 0.6, 2.2, 12.5, 1.0, 0.6, 1024, 0.0), img)
    assert nl == len(oK) and np.array_equal(kl.view(np.uint8), oK.view(np.uint8)) and np.array_equal(ld, oD)
    om, on = oracle.match_nnr(od, od, 0.9)
    assert nm == on and np.array_equal(m12, om)
    assert d01 == oracle.descriptor_distance(od[0], od[1])
    oxL = oracle.ORBextractor(500, 1.2, 6, 20, 7); oxR = oracle.ORBextractor(500, 1.2, 6, 20, 7)
    okL, odL = oxL(img); okR, odR = oxR(np.roll(img, -9, axis=1).copy())
    ou, oz = oracle.stereo_match(oxL, oxR, okL, odL, okR, odR, np.float32(0.11), np.float32(0.11) * np.float32(435.2))
    assert (ou >= 0).sum() > 20
    assert np.array_equal(uR.view(np.uint32), ou.view(np.uint32)) and np.array_equal(dep.view(np.uint32), oz.view(np.uint32))
    assert k2 == 3 and levels == 6 and (p1w, p1h) == (533, 400)    # empty image left the caller's vector untouched
