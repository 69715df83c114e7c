"""Launch timeline of the timed configuration: python profiles/c4_timeline.py c4 1024 line|both [line_contexts] [line_sub] > timeline
(CUDA-event brackets around every launch of every context; columns: start ms, end ms, context, kernel)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, spl_slam_b200 as S
cfgname, B, what = sys.argv[1], int(sys.argv[2]), sys.argv[3]
args = types.SimpleNamespace(orb_contexts=bench.ORB_CONTEXTS, line_contexts=int(sys.argv[4]) if len(sys.argv) > 4 else bench.LINE_CONTEXTS,
                             line_sub=int(sys.argv[5]) if len(sys.argv) > 5 else 0, orb_sub=0)
ex = bench.Extraction(S, torch, bench.CONFIGS[cfgname], B, 0, 1, 0, args)
ex.run_device(2, what)
print("# plain: %.2f ms per step" % (ex.run_device(3, what) / 3))
for c in ex.contexts(): c.profile_enable(True)
t = ex.run_device(2, what) / 2
print("# with brackets: %.2f ms per step" % t)
rows = []
for ci, c in enumerate(ex.contexts()):
    for name, t0, t1 in c.profile_timeline(ex.ctx_o): rows.append((t0, t1, ci, name))
for t0, t1, ci, name in sorted(rows): print("%9.3f %9.3f %d %s" % (t0, t1, ci, name))
