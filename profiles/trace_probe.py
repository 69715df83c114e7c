"""Host-side phase trace of one drop-in line call: PLF_TRACE=1 python profiles/trace_probe.py c1|c2|c3|c4"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, spl_slam_b200 as S
cfg = bench.CONFIGS[sys.argv[1]]; W, H, line = cfg["W"], cfg["H"], cfg["line"]
imgs = [bench.synth_image(W, H, s) for s in range(4)]
cl = S.Context(0, priority=1)
le = S.Lineextractor(line["nfeatures"], line["nlevels"], line["refine"], line["scale"], line["sigma_scale"], line["quant"], line["ang_th"],
                     line["log_eps"], line["density_th"], line["n_bins"], line["min_line_length"], ctx=cl)
for r in range(8):
    t0 = time.perf_counter(); le.ComputeLsdWithLbd(imgs[r % 4]); print("call %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
