"""Brute-force Hamming top-2 sweep (BASELINE config 5): 1e4 queries x T train rows on one GPU, popc roofline fraction."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spl_slam_b200 as S
ctx = S.Context(0)
peak = ctx.popc_peak()
NQ = 10000
g = torch.Generator(device="cuda"); g.manual_seed(1)
dq = torch.randint(0, 256, (NQ, 32), dtype=torch.uint8, device="cuda", generator=g)
for NT in (10**4, 10**5, 10**6, 10**7):
    dt = torch.randint(0, 256, (NT, 32), dtype=torch.uint8, device="cuda", generator=g)
    idx = torch.empty((NQ, 2), dtype=torch.int32, device="cuda"); dist = torch.empty_like(idx)
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); ctx.timer_start()
        ctx.check(ctx.lib.plf_hamming_knn2_device(ctx.h, dq.data_ptr(), NQ, dt.data_ptr(), NT, 0, idx.data_ptr(), dist.data_ptr()))
        ms = ctx.timer_stop()
        if rep: best = min(best, ms)
    print("T=%.0e: %.3f ms, %.3e pairs/s, %.0f queries/s, popc frac %.3f" % (NT, best, NQ * NT / best * 1e3, NQ / best * 1e3, 8 * NQ * NT / best * 1e3 / peak), flush=True)
