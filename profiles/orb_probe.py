"""Per-kernel times of one ORB batch (CUDA-event brackets, one stream): python profiles/orb_probe.py c2|c4 NFRAMES"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, spl_slam_b200 as S
cfg = bench.CONFIGS[sys.argv[1]]; B = int(sys.argv[2]); W, H = cfg["W"], cfg["H"]
frames = bench.make_frames(cfg, list(range(B)))
ctx = S.Context(0); O = cfg["orb"]
orb = S.ORBextractor(O["nfeatures"], O["scaleFactor"], O["nlevels"], O["iniThFAST"], O["minThFAST"], ctx=ctx)
cap = orb.max_keypoints
d = torch.from_numpy(frames).cuda()
kps = torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda"); desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda"); nk = torch.empty(B, dtype=torch.int32, device="cuda")
def run(): ctx.check(ctx.lib.plf_orb_extract_batch_device(orb.h, d.data_ptr(), B, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), cap, nk.data_ptr())); ctx.synchronize()
for _ in range(3): run()
ctx.profile_enable(True)
for _ in range(5): run()
for k, v in sorted(ctx.profile_report().items(), key=lambda kv: -kv[1][0]): print("  %-18s %8.3f ms per batch" % (k, v[0] / 5))
