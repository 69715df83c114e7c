"""Times the host-buffer (e2e) calls: ORB-only, lines-only, for a few context counts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from concurrent.futures import ThreadPoolExecutor
import spl_slam_b200 as S
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frames = bench.make_frames(B // 2, 0)
h_img = torch.from_numpy(frames).pin_memory()
W, H = bench.W, bench.H
O, L = bench.ORB, bench.LINE


def run(NO, NL, reps=3, steps=3):
    cos = [S.Context(0, priority=-1) for _ in range(NO)]
    cls = [S.Context(0, priority=1) for _ in range(NL)]
    orbs = [S.ORBextractor(O["nfeatures"], O["scaleFactor"], O["nlevels"], O["iniThFAST"], O["minThFAST"], ctx=c) for c in cos]
    les = [S.Lineextractor(L["nfeatures"], L["nlevels"], L["refine"], L["scale"], L["sigma_scale"], L["quant"], L["ang_th"],
                           L["log_eps"], L["density_th"], L["n_bins"], L["min_line_length"], ctx=c) for c in cls]
    lib = (cos + cls)[0].lib
    capk = orbs[0].max_keypoints if NO else 1
    capl = les[0].max_keylines if NL else 1
    h_kps = torch.empty((B, capk, 28), dtype=torch.uint8).pin_memory(); h_desc = torch.empty((B, capk, 32), dtype=torch.uint8).pin_memory()
    h_kl = torch.empty((B, capl, 68), dtype=torch.uint8).pin_memory(); h_mid = torch.empty((B, capl, 28), dtype=torch.uint8).pin_memory()
    h_ld = torch.empty((B, capl, 32), dtype=torch.uint8).pin_memory()
    n_k = np.zeros(B, np.int32); n_l = np.zeros(B, np.int32)
    pool = ThreadPoolExecutor(max(NL + NO, 1))
    BO = B // max(NO, 1); BL = B // max(NL, 1)
    tt = {}

    def orb_i(i):
        for _ in range(steps):
            t0 = time.perf_counter()
            s = slice(i * BO, (i + 1) * BO)
            cos[i].check(lib.plf_orb_extract_batch(orbs[i].h, h_img[s].data_ptr(), BO, W, H, W, W * H, h_kps[s].data_ptr(), h_desc[s].data_ptr(), capk, n_k[s].ctypes.data))
            tt.setdefault("orb%d" % i, []).append((time.perf_counter() - t0) * 1e3)

    def line_i(i):
        for _ in range(steps):
            t0 = time.perf_counter()
            s = slice(i * BL, (i + 1) * BL)
            cls[i].check(lib.plf_line_extract_batch(les[i].h, h_img[s].data_ptr(), BL, W, H, W, W * H, h_kl[s].data_ptr(), h_mid[s].data_ptr(), h_ld[s].data_ptr(), capl, n_l[s].ctypes.data))
            tt.setdefault("line%d" % i, []).append((time.perf_counter() - t0) * 1e3)

    ts = []
    for rep in range(reps):
        tt.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        futs = [pool.submit(orb_i, i) for i in range(NO)] + [pool.submit(line_i, i) for i in range(NL)]
        for f in futs:
            f.result()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3 / steps)
    return min(ts[1:]), {k: [round(x, 1) for x in v] for k, v in tt.items()}


print("e2e ORB only 2 ctx: %.1f ms/step %s" % run(2, 0), flush=True)
print("e2e ORB only 1 ctx: %.1f ms/step %s" % run(1, 0), flush=True)
print("e2e lines only 4 ctx: %.1f ms/step %s" % run(0, 4), flush=True)
print("e2e lines only 1 ctx: %.1f ms/step %s" % run(0, 1), flush=True)
print("e2e both 2+4: %.1f ms/step %s" % run(2, 4), flush=True)
