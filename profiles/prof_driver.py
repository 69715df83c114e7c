"""Small driver for ncu captures: runs the line and/or ORB path of one bench configuration on a batch of synthetic frames.
python profiles/prof_driver.py line|orb|both NFRAMES REPS [c2|c3|c4]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import spl_slam_b200 as S
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "line"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = bench.CONFIGS[sys.argv[4] if len(sys.argv) > 4 else "c4"]
W, H = cfg["W"], cfg["H"]
frames = bench.make_frames(cfg, list(range(B)))
ctx = S.Context(0)
d_img = torch.from_numpy(frames).cuda()
if which in ("line", "both"):
    L = cfg["line"]
    le = S.Lineextractor(L["nfeatures"], L["nlevels"], L["refine"], L["scale"], L["sigma_scale"], L["quant"], L["ang_th"],
                         L["log_eps"], L["density_th"], L["n_bins"], L["min_line_length"], ctx=ctx)
    cap = le.max_keylines
    kl = torch.empty((B, cap, 68), dtype=torch.uint8, device="cuda"); mid = torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda")
    ld = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda"); nl = torch.empty(B, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        ctx.check(ctx.lib.plf_line_extract_batch_device(le.h, d_img.data_ptr(), B, W, H, W, W * H,
                                                        kl.data_ptr(), mid.data_ptr(), ld.data_ptr(), cap, nl.data_ptr()))
        ctx.synchronize()
    print("lines per frame", nl.cpu().numpy()[:4])
if which in ("orb", "both"):
    O = cfg["orb"]
    orb = S.ORBextractor(O["nfeatures"], O["scaleFactor"], O["nlevels"], O["iniThFAST"], O["minThFAST"], ctx=ctx)
    cap = orb.max_keypoints
    kps = torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda"); desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
    nk = torch.empty(B, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        ctx.check(ctx.lib.plf_orb_extract_batch_device(orb.h, d_img.data_ptr(), B, W, H, W, W * H,
                                                       kps.data_ptr(), desc.data_ptr(), cap, nk.data_ptr()))
        ctx.synchronize()
    print("keypoints per frame", nk.cpu().numpy()[:4])
