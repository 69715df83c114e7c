"""Where does the step go?  Times ORB-only, lines-only and the combined step for several context counts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from concurrent.futures import ThreadPoolExecutor
import spl_slam_b200 as S
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frames = bench.make_frames(B // 2, 0)
d_img = torch.from_numpy(frames).cuda()
W, H = bench.W, bench.H
O, L = bench.ORB, bench.LINE


def setup(NO, NL):
    cos = [S.Context(0) for _ in range(NO)]
    cls = [S.Context(0) for _ in range(NL)]
    orbs = [S.ORBextractor(O["nfeatures"], O["scaleFactor"], O["nlevels"], O["iniThFAST"], O["minThFAST"], ctx=c) for c in cos]
    les = [S.Lineextractor(L["nfeatures"], L["nlevels"], L["refine"], L["scale"], L["sigma_scale"], L["quant"], L["ang_th"],
                           L["log_eps"], L["density_th"], L["n_bins"], L["min_line_length"], ctx=c) for c in cls]
    return cos, cls, orbs, les


def run(NO, NL, do_orb, do_line, sub=0, reps=4):
    cos, cls, orbs, les = setup(NO, NL)
    lib = (cos + cls)[0].lib
    capk = orbs[0].max_keypoints if NO else 1
    capl = les[0].max_keylines if NL else 1
    kps = torch.empty((B, capk, 28), dtype=torch.uint8, device="cuda"); desc = torch.empty((B, capk, 32), dtype=torch.uint8, device="cuda")
    nk = torch.empty(B, dtype=torch.int32, device="cuda")
    kl = torch.empty((B, capl, 68), dtype=torch.uint8, device="cuda"); mid = torch.empty((B, capl, 28), dtype=torch.uint8, device="cuda")
    ld = torch.empty((B, capl, 32), dtype=torch.uint8, device="cuda"); nl = torch.empty(B, dtype=torch.int32, device="cuda")
    pool = ThreadPoolExecutor(max(NL, 1))
    BO = B // max(NO, 1); BL = B // max(NL, 1)
    SUB = sub if sub else BL

    def orb_i(i):
        s = slice(i * BO, (i + 1) * BO)
        cos[i].check(lib.plf_orb_extract_batch_device(orbs[i].h, d_img[s].data_ptr(), BO, W, H, W, W * H, kps[s].data_ptr(), desc[s].data_ptr(), capk, nk[s].data_ptr()))

    def line_i(i):
        for j in range(i * BL, (i + 1) * BL, SUB):
            s = slice(j, j + SUB)
            cls[i].check(lib.plf_line_extract_batch_device(les[i].h, d_img[s].data_ptr(), SUB, W, H, W, W * H, kl[s].data_ptr(), mid[s].data_ptr(), ld[s].data_ptr(), capl, nl[s].data_ptr()))

    ts = []
    for rep in range(reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if do_orb:
            for i in range(NO):
                orb_i(i)
        if do_line:
            list(pool.map(line_i, range(NL)))
        for c in cos + cls:
            c.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts[2:])


print("ORB only, 2 ctx: %.1f ms" % run(2, 0, True, False), flush=True)
for NL in (1, 4):
    print("lines only, %d ctx: %.1f ms" % (NL, run(0, NL, False, True)), flush=True)
print("both, orb 2 line 4: %.1f ms" % run(2, 4, True, True), flush=True)
