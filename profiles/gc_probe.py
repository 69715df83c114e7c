"""Diagnosis of the speculative region-growing kernel: per-kernel times of one line extraction call and (PLF_GC_DEBUG=1)
the kernel's own counters.  python profiles/gc_probe.py W H NFRAMES"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import spl_slam_b200 as S

W, H, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = bench.CONFIGS["c4" if W > 1300 else "c1"]
frames = np.stack([bench.synth_image(W, H, s) for s in range(N)])
ctx = S.Context(0)
L = cfg["line"]
le = S.Lineextractor(L["nfeatures"], L["nlevels"], L["refine"], L["scale"], L["sigma_scale"], L["quant"], L["ang_th"], L["log_eps"], L["density_th"],
                     L["n_bins"], L["min_line_length"], ctx=ctx)
for rep in range(3):
    t0 = time.perf_counter()
    out = le.extract_batch(frames)
    print("rep %d: %.2f ms for %d frames" % (rep, (time.perf_counter() - t0) * 1e3, N), file=sys.stderr)
ctx.profile_enable(True)
le.extract_batch(frames)
rep = ctx.profile_report()
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])[:12]:
    print("  %-22s %8.3f ms  %d launches" % (k, v[0], v[1]), file=sys.stderr)
